"""Resident kernel with one env spread over a thread-block cluster (DSMEM histogram exchange): the integer density
makes it bit-identical to the one-CTA kernel in x, v, density and field energy for every cluster size and thread
count; the kinetic sums (float) agree to rounding.  Also: envs too large for ONE CTA's shared memory run resident in
a cluster and match the streaming kernels bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pic_oracle as O  # noqa: E402  (checker only)


def _run(N, M, B, shape, steps, coeffs, x, v, precision="f64", interpol="CIC", m=3, mode="resident", dt=0.05):
    import pic_b200
    eng = pic_b200.Engine(N, M, 50.0, dt, n_envs=B, mode=mode, deposit="split32", max_mode=m, precision=precision,
                          interpol=interpol)
    if shape is not None:
        eng.set_tuning(shape[0], shape[1], -1)
    bc, bs = O.actuator_basis(50.0, M, m)
    eng.set_actuator_basis(bc, bs)
    eng.enable_modes(m)
    eng.set_state(x, v)
    d0 = eng.get_diag()
    eng.step_coeffs(coeffs, steps)
    xs, vs = eng.get_state()
    rho, k = eng.get_density_fixed()
    out = dict(x=xs, v=vs, rho=rho, diag=eng.get_diag(), trace=eng.get_trace(steps), d0=d0, modes=eng.get_mode_trace(steps),
               fields=eng.get_fields(), flags=eng.error_flags(), info=eng.launch_info())
    eng.close()
    return out


def _same(a, b):
    for key in ("x", "v", "rho"):
        assert np.array_equal(a[key], b[key]), key
    for arr_a, arr_b in ((a["diag"], b["diag"]), (a["trace"].reshape(-1, 6), b["trace"].reshape(-1, 6)), (a["d0"], b["d0"])):
        assert np.array_equal(arr_a[:, [1, 3, 4, 5]], arr_b[:, [1, 3, 4, 5]])          # PE_mesh, sum E^2, reward, input energy
        assert np.allclose(arr_a[:, 0], arr_b[:, 0], rtol=1e-14, atol=0)             # kinetic energy: float sum, other order
        assert np.allclose(arr_a[:, 2], arr_b[:, 2], rtol=0, atol=1e-10)
    assert np.array_equal(a["modes"], b["modes"])
    assert np.array_equal(a["fields"][0], b["fields"][0]) and np.array_equal(a["fields"][1], b["fields"][1])
    assert a["flags"] == 0 and b["flags"] == 0


@pytest.mark.parametrize("shape", [(256, 2), (512, 2), (1024, 2), (256, 4), (512, 4)])
def test_cluster_env_is_bit_identical_to_one_cta_env(shape):
    B, N, M, steps = 5, 5000, 250, 6
    rng = np.random.RandomState(12)
    x = rng.uniform(0, 50.0, (B, N)); v = rng.normal(size=(B, N)) + 3.0 * (rng.uniform(size=(B, N)) < 0.17)
    x[1, :40] = rng.uniform(-300, 300, 40)                     # far outside the box: the careful path inside a cluster
    coeffs = rng.uniform(-1, 1, (steps, B, 6))
    base = _run(N, M, B, (512, 1), steps, coeffs, x, v)
    clu = _run(N, M, B, shape, steps, coeffs, x, v)
    assert clu["info"]["per_thread"] == shape[1] and clu["info"]["grid_x"] == B * shape[1]
    _same(base, clu)


@pytest.mark.parametrize("N,M", [(4999, 250), (3, 5), (257, 31)])
def test_cluster_ragged_sizes(N, M):
    B, steps = 3, 3
    m = 3 if M > 6 else 1                                      # spectral read-out needs n_modes < N_mesh // 2
    rng = np.random.RandomState(N)
    x = rng.uniform(0, 50.0, (B, N)); v = rng.normal(size=(B, N))
    coeffs = rng.uniform(-1, 1, (steps, B, 2 * m))
    _same(_run(N, M, B, (256, 1), steps, coeffs, x, v, dt=0.02, m=m), _run(N, M, B, (256, 4), steps, coeffs, x, v, dt=0.02, m=m))


def test_cluster_f32_and_tsc():
    B, N, M, steps = 3, 5000, 250, 4
    rng = np.random.RandomState(5)
    x = rng.uniform(0, 50.0, (B, N)); v = rng.normal(size=(B, N))
    coeffs = rng.uniform(-1, 1, (steps, B, 6))
    _same(_run(N, M, B, (512, 1), steps, coeffs, x, v, precision="f32"), _run(N, M, B, (256, 2), steps, coeffs, x, v, precision="f32"))
    _same(_run(N, M, B, (512, 1), steps, coeffs, x, v, interpol="TSC"), _run(N, M, B, (256, 2), steps, coeffs, x, v, interpol="TSC"))


def test_env_too_large_for_one_cta_runs_resident_in_a_cluster():
    """N = 40 000, N_mesh = 500 (640 KB of particle state): four CTAs of a cluster hold it; same bits as streaming."""
    import pic_b200
    B, N, M, steps = 2, 40_000, 500, 3
    rng = np.random.RandomState(8)
    x = rng.uniform(0, 50.0, (B, N)); v = rng.normal(size=(B, N)) + 3.0 * (rng.uniform(size=(B, N)) < 0.17)
    coeffs = rng.uniform(-1, 1, (steps, B, 10))
    with pytest.raises(pic_b200.PicError):
        _run(N, M, B, (1024, 1), steps, coeffs, x, v, m=5)                 # one CTA: does not fit
    clu = _run(N, M, B, (1024, 4), steps, coeffs, x, v, m=5)
    stream = _run(N, M, B, None, steps, coeffs, x, v, m=5, mode="streaming")
    for key in ("x", "v", "rho"):
        assert np.array_equal(clu[key], stream[key]), key
    assert np.array_equal(clu["diag"][:, 1], stream["diag"][:, 1])
    p = O.PicParams(N=N, N_mesh=M, n0=1.0, L=50.0, dt=O.clip_dt(0.05, N, 50.0))
    bc, bs = O.actuator_basis(50.0, M, 5)
    xo, vo = x[0], v[0]
    for t in range(steps):
        o = O.step(xo, vo, p, O.actuator_field(bc, bs, coeffs[t, 0, :5], coeffs[t, 0, 5:]))
        xo, vo = o["x"], o["v"]
    assert np.abs(clu["x"][0] - xo).max() < 1e-11 and np.abs(clu["v"][0] - vo).max() < 1e-11


def test_auto_mode_uses_a_cta_pair_for_large_batches_of_mid_size_envs():
    """20 000-particle envs do not fit one CTA; a batch that fills the GPU runs them resident as CTA pairs (measured
    faster than the streaming kernels), a handful of them streams."""
    import pic_b200
    many = pic_b200.Engine(20_000, 250, 50.0, 0.02, n_envs=160)
    assert many.launch_info()["mode"] == "resident" and many.launch_info()["per_thread"] == 2
    few = pic_b200.Engine(20_000, 250, 50.0, 0.02, n_envs=4)
    assert few.launch_info()["mode"] == "streaming"
    tsc = pic_b200.Engine(20_000, 250, 50.0, 0.02, n_envs=160, interpol="TSC")
    assert tsc.launch_info()["mode"] == "resident"
    rng = np.random.RandomState(2)
    x = rng.uniform(0, 50.0, (160, 20_000)); v = rng.normal(size=(160, 20_000))
    many.set_state(x, v); many.step_mesh(None, 2)
    s = pic_b200.Engine(20_000, 250, 50.0, 0.02, n_envs=160, mode="streaming")
    s.set_state(x, v); s.step_mesh(None, 2)
    assert np.array_equal(many.get_state()[0], s.get_state()[0]) and many.error_flags() == 0
