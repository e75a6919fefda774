"""The oracle (oracle/pic_oracle.py) against vectors produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import pic_oracle as O


def params(g):
    return O.PicParams(N=int(g["N"]), N_mesh=int(g["N_mesh"]), n0=1.0, L=float(g["L"]), dt=float(g["dt"]))


@pytest.mark.parametrize("name,simcase,kw", [
    ("bump_vb3", "bump-on-tail", {}),
    ("twostream_vb3", "two-stream", {}),
    ("bump_vb5", "bump-on-tail", {"vb": 5.0}),
])
def test_sampler_and_init_bit_exact(golden, name, simcase, kw):
    """src/env/dist.py samplers + pic.py:64-68 restated: same legacy RNG stream, same particles."""
    g = golden(name)
    x, v = O.runner_initial_state(simcase, **kw)
    assert np.array_equal(x, g["t0_x"])
    assert np.array_equal(v, g["t0_v"])
    f = O.init_fields(x, params(g), faithful=True)
    assert np.array_equal(f["indx_l"], g["t0_indx_l"])
    assert np.array_equal(f["n"], g["t0_n"])
    assert np.array_equal(f["E_mesh"], g["t0_E_mesh"])


def test_dt_clip(golden):
    g = golden("clip_dt")
    assert O.clip_dt(0.1, 40000, 50.0) == float(g["dt"])
    assert O.clip_dt(0.1, 5000, 50.0) == 0.1
    p = params(g)
    x, v = g["t0_x"].copy(), g["t0_v"].copy()
    for _ in range(5):
        o = O.step(x, v, p, None, faithful=False)
        x, v = o["x"], o["v"]
    assert np.abs(x - g["t5_x"]).max() < 1e-12
    assert np.abs(v - g["t5_v"]).max() < 1e-12


@pytest.mark.parametrize("name", ["bump_vb3", "twostream_vb3"])
def test_faithful_oracle_is_bit_identical_for_10_steps(golden, name):
    g = golden(name)
    p = params(g)
    x, v = g["t0_x"].copy(), g["t0_v"].copy()
    for t in range(1, 11):
        o = O.step(x, v, p, None, faithful=True)
        x, v = o["x"], o["v"]
        if t in (1, 10):
            for k in ("x", "v", "n", "E_mesh", "E"):
                assert np.array_equal(o[k], g[f"t{t}_{k}"]), (t, k)
            assert np.array_equal(o["indx_l"], g[f"t{t}_indx_l"])
        assert O.hamiltonian(x, v, p, faithful=True) == g["H"][t]
        assert O.electric_energy(x, p, faithful=True) == g["PE"][t]


@pytest.mark.parametrize("name", ["bump_vb3", "twostream_vb3"])
def test_lean_oracle_500_steps(golden, name):
    """Prefix-sum field + no discarded work: indices bit-equal, state within the stated fp64 tolerance
    (per step 1e-12; chaos amplifies ulp noise to ~1e-9 by step 500)."""
    g = golden(name)
    p = params(g)
    x, v = g["t0_x"].copy(), g["t0_v"].copy()
    pem = []
    for t in range(1, 501):
        o = O.step(x, v, p, None, faithful=False)
        x, v = o["x"], o["v"]
        pem.append(O.pe_mesh(o["E_mesh"], p.dx))
        if t in (1, 10):
            assert np.abs(x - g[f"t{t}_x"]).max() < 1e-12
            assert np.abs(v - g[f"t{t}_v"]).max() < 1e-12
            assert np.abs(o["E_mesh"] - g[f"t{t}_E_mesh"]).max() < 1e-12 * max(1.0, np.abs(g[f"t{t}_E_mesh"]).max())
            assert np.array_equal(o["indx_l"], g[f"t{t}_indx_l"])
        if t == 500:
            assert np.abs(x - g["t500_x"]).max() < 1e-6
            assert np.array_equal(o["indx_l"], g["t500_indx_l"])
    pem = np.array(pem)
    assert np.max(np.abs(pem - g["PE_mesh"][1:]) / g["PE_mesh"][1:]) < 1e-9
    assert abs(g["sum_v"][-1] - g["sum_v"][0]) < 1e-9          # momentum conservation without control


def test_published_growth_rates(golden):
    """analysis/optimal_control_two_stream.ipynb:52 -> 0.02135; optimal_control_bump_on_tail.ipynb:51 -> 0.00557
    (the latter was produced with --vb 5.0)."""
    ts = O.growth_rate(golden("twostream_vb3")["PE_mesh"][1:], 50.0)
    b5 = O.growth_rate(golden("bump_vb5")["PE_mesh"][1:], 50.0)
    b3 = O.growth_rate(golden("bump_vb3")["PE_mesh"][1:], 50.0)
    assert round(ts, 5) == 0.02135
    assert round(b5, 5) == 0.00557
    assert abs(b3 - (-0.001295)) < 1e-6


@pytest.mark.parametrize("name,steps,m", [("bump_vb3_constctrl", 10, 3), ("bump_vb3_randctrl", 200, 3),
                                          ("sac_cfg", 40, 5)])
def test_controlled_runs(golden, name, steps, m):
    g = golden(name)
    p = params(g)
    bc, bs = O.actuator_basis(p.L, p.N_mesh, m)
    assert np.array_equal(bc, g["basis_cos"]) and np.array_equal(bs, g["basis_sin"])
    x, v = g["t0_x"].copy(), g["t0_v"].copy()
    pe_prev = O.pe_mesh(g["t0_E_mesh"], p.dx)
    for t in range(1, steps + 1):
        c = g["coeffs"][t - 1]
        E_ext = O.actuator_field(bc, bs, c[:m], c[m:])
        assert np.abs(E_ext - g["E_ext"][t - 1]).max() < 1e-14
        r = O.reward(pe_prev, c, p.L)                       # reward uses the PRE-step state
        assert abs(r - g["rewards"][t - 1]) < 1e-9
        o = O.step(x, v, p, E_ext, faithful=False)
        x, v = o["x"], o["v"]
        pe_prev = O.pe_mesh(o["E_mesh"], p.dx)
        assert abs(pe_prev - g["PE_mesh"][t]) < 1e-9 * max(1.0, g["PE_mesh"][t])
        if f"t{t}_x" in g.files:
            tol = 1e-12 if t <= 10 else 1e-8
            assert np.abs(x - g[f"t{t}_x"]).max() < tol
            assert np.abs(v - g[f"t{t}_v"]).max() < tol
            assert np.array_equal(o["indx_l"], g[f"t{t}_indx_l"])
    assert abs(O.hamiltonian(x, v, p) - g["H"][steps]) < 1e-8 * g["H"][steps]


def test_deposit_edge_cases(golden):
    g = golden("deposit_edges")
    for (L, M) in [(50.0, 250), (50.0, 500), (50.0, 4096), (10.0, 64)]:
        key = f"L{L:g}_M{M}"
        x = g[key + "_x"].copy()
        N = x.shape[0]
        assert np.array_equal(O.wrap(x, L), g[key + "_xw"])
        n, il, ir, wl, wr = O.compute_n(x, L / M, M, 1.0, L, N)
        assert np.array_equal(il, g[key + "_cic_il"]) and np.array_equal(ir, g[key + "_cic_ir"])
        assert np.array_equal(wl, g[key + "_cic_wl"]) and np.array_equal(wr, g[key + "_cic_wr"])
        assert np.array_equal(n, g[key + "_cic_n"])
        nt, tl, tm, tr, *_ = O.tsc(g[key + "_xw"], 1.0, L, N, M, L / M)
        assert np.array_equal(tm, g[key + "_tsc_im"])
        assert np.array_equal(nt, g[key + "_tsc_n"])


def test_prefix_field_equals_dense_field():
    rng = np.random.RandomState(3)
    for (L, M) in [(50.0, 250), (50.0, 500), (50.0, 4096)]:
        b = rng.normal(size=M) * 0.3
        b -= b.mean()
        n = 1.0 + b
        _, Ed = O.field_dense(n, 1.0, L, M)
        Ep = O.field_prefix(n, 1.0, L, M)
        assert np.abs(Ed - Ep).max() < 1e-10 * max(1.0, np.abs(Ed).max())


@pytest.mark.parametrize("name,steps,ctrl", [("bump_vb3_tsc", 40, False), ("twostream_tsc_ctrl", 20, True)])
def test_tsc_interpolation(golden, name, steps, ctrl):
    """interpol="TSC" (src/env/interpolate.py:22-44, run_wo_oc.py --interpol TSC)."""
    g = golden(name)
    p = params(g)
    p.interpol = "TSC"
    bc, bs = O.actuator_basis(p.L, p.N_mesh, 3)
    for faithful, tol in ((True, 0.0), (False, 1e-11)):
        x, v = g["t0_x"].copy(), g["t0_v"].copy()
        for t in range(1, steps + 1):
            e = None
            if ctrl:
                c = g["coeffs"][t - 1]
                e = O.actuator_field(bc, bs, c[:3], c[3:])
            o = O.step(x, v, p, e, faithful=faithful)
            x, v = o["x"], o["v"]
            if f"t{t}_x" in g.files:
                assert np.abs(x - g[f"t{t}_x"]).max() <= tol and np.abs(v - g[f"t{t}_v"]).max() <= tol
                assert np.array_equal(o["indx_m"], g[f"t{t}_indx_m"])
                assert np.abs(o["E_mesh"] - g[f"t{t}_E_mesh"]).max() <= max(tol, 0.0) * 10 + (0 if faithful else 1e-12)
        assert abs(O.hamiltonian(x, v, p, faithful=faithful) - g["H"][steps]) <= (0 if faithful else 1e-9 * g["H"][steps])
