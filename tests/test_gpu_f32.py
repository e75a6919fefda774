"""float32 mode (`precision="f32"`): particles, weights, gather and push in float32; density accumulation and the
mesh field stay integer / float64.  Stated tolerances: one step |dx| <= 2e-5 (x up to 50), |dv| <= 1e-5, PE rel 1e-4
against the float64 reference; cell indices bit-exact against the float32 restatement (oracle.step_f32); 500-step
energy trace rel <= 2e-2 and growth rate within 2e-4 of the float64 reference."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pic_oracle as O  # noqa: E402  (checker only)


@pytest.mark.parametrize("mode", ["resident", "streaming"])
def test_f32_one_step_vs_f32_restatement_and_f64_reference(golden, mode):
    from pic_b200 import Engine
    g = golden("bump_vb3")
    N, M, L, dt = 5000, 250, 50.0, 0.1
    p = O.PicParams(N=N, N_mesh=M, n0=1.0, L=L, dt=dt)
    eng = Engine(N, M, L, dt, precision="f32", mode=mode)
    eng.set_state(g["t0_x"][None], g["t0_v"][None])
    x0, v0 = eng.get_state()                                  # the float32-rounded initial state
    assert np.array_equal(x0[0], g["t0_x"].astype(np.float32).astype(np.float64))
    eng.step_mesh(None, 1)
    x, v = eng.get_state()
    o = O.step_f32(x0[0], v0[0], p)
    il, *_ = eng.get_cells(False, False)
    assert np.array_equal(il[0], o["indx_l"])                 # index parity in float32 arithmetic
    assert np.abs(x[0] - o["x"]).max() < 1e-5 and np.abs(v[0] - o["v"]).max() < 1e-6
    assert np.abs(x[0] - g["t1_x"]).max() < 2e-5 and np.abs(v[0] - g["t1_v"]).max() < 1e-5
    d = eng.get_diag()[0]
    assert abs(d[1] - g["PE_mesh"][1]) < 1e-4 * g["PE_mesh"][1]
    assert eng.error_flags() == 0


@pytest.mark.parametrize("name,rate", [("twostream_vb3", 0.021354), ("bump_vb5", 0.005568)])
def test_f32_500_step_trace(golden, name, rate):
    from pic_b200 import Engine
    g = golden(name)
    N, M, L, dt = int(g["N"]), int(g["N_mesh"]), float(g["L"]), float(g["dt"])
    eng = Engine(N, M, L, dt, precision="f32")
    eng.set_state(g["t0_x"][None], g["t0_v"][None])
    eng.step_mesh(None, 500)
    pe = eng.get_trace(500)[:, 0, 1]
    rel = np.abs(pe - g["PE_mesh"][1:]) / g["PE_mesh"][1:]
    assert rel[:100].max() < 1e-3 and rel.max() < 2e-2
    assert abs(O.growth_rate(pe, 50.0) - rate) < 2e-4
    assert eng.error_flags() == 0


def test_f32_large_streaming_invariants():
    from pic_b200 import Engine
    N, M, L = 20_000_001, 4096, 50.0
    eng = Engine(N, M, L, 2 / np.sqrt(N / L), precision="f32", mode="streaming")
    eng.sample_state("bump-on-tail", seed=9)
    d0 = eng.get_diag()[0]
    rho, k = eng.get_density_fixed()
    assert sum(int(r) for r in rho.ravel()) == N * (1 << k)
    eng.step_mesh(None, 10)
    tr = eng.get_trace(10)[:, 0, :]
    H0 = d0[0] + d0[1] * N / L
    assert np.max(np.abs(tr[:, 0] + tr[:, 1] * N / L - H0)) / H0 < 1e-5
    assert eng.error_flags() == 0


@pytest.mark.parametrize("M,mode", [(4096, "streaming"), (250, "resident"), (1000, "streaming")])
def test_f32_cell_edges_deposit_bit_exact(M, mode):
    """The float32 cell-index bracket (fast_cell, eps = 2^-22) at its worst: positions on every cell edge and 1..8 ulp to
    either side.  The integer density after set_state must equal, bit for bit, the one built from floor(x / dx) in
    float32 arithmetic (np.floor of the IEEE quotient) -- a particle put into the neighbouring cell would deposit a
    weight near 2^k or below zero and change the sum."""
    from pic_b200 import Engine
    f = np.float32
    L = 50.0
    dx = f(f(L) / f(M)) if False else f(L / M)                # the library rounds the float64 dx to float32
    edges = (np.arange(M, dtype=np.float64) * (L / M)).astype(f)
    xs = [edges]
    for k in (1, 2, 3, 5, 8):
        up, dn = edges.copy(), edges.copy()
        for _ in range(k):
            up = np.nextafter(up, f(np.inf)); dn = np.nextafter(dn, f(-np.inf))
        xs += [up, dn]
    rng = np.random.RandomState(3)
    xs.append(rng.uniform(0, L, 20 * M).astype(f))
    x = np.concatenate(xs)
    x = x[(x >= 0) & (x < f(L))]
    N = x.size
    eng = Engine(N, M, L, 0.01, precision="f32", mode=mode)
    eng.set_state(x.astype(np.float64)[None], np.zeros((1, N)))
    rho, k = eng.get_density_fixed()
    il = np.floor(x / dx).astype(np.int64)                    # float32 division, then floor
    assert il.min() >= 0 and il.max() < M
    wr = ((x - il.astype(f) * dx) * (f(1) / dx)).astype(f)    # the library's weights: separate float32 roundings
    W = np.rint(wr.astype(np.float64) * 2.0 ** k).astype(np.int64)
    exp = np.zeros(M, dtype=np.int64)
    np.add.at(exp, il, (1 << k) - W)
    np.add.at(exp, (il + 1) % M, W)
    assert np.array_equal(rho[0].astype(np.int64), exp)
    got, *_ = eng.get_cells(False, False)
    assert np.array_equal(got[0], il)
    assert eng.error_flags() == 0
    eng.close()
