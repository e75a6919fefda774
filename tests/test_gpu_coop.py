"""The cooperative single-launch step (pic_set_coop: three passes + finalize of every step of a call inside ONE
cooperative kernel, grid barriers instead of kernel boundaries) against the kernel-per-pass path: the same device
functions on the same integer densities, so x, v, the fixed-point density and the fields must be IDENTICAL bits; only
the kinetic sums are added over a different number of CTAs when the env has more tiles than the device has SMs.
(The reference has no counterpart: both are `PIC.update_state`, src/env/pic.py:131-146.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
L = 50.0


def _engine(N, M, n_envs, precision="f64", interpol="CIC", with_basis=False):
    import pic_b200
    dt = 2 / np.sqrt(N / L) if N > 50000 else 0.05
    eng = pic_b200.Engine(N, M, L, dt, mode="streaming", n_envs=n_envs, precision=precision, interpol=interpol,
                          max_mode=3 if with_basis else 0)
    if with_basis:
        from pic_b200.actuator import E_field
        act = E_field(L, M, 3)
        eng.set_actuator_basis(act.basis_cos, act.basis_sin)
    return eng


def _state(N, n_envs, seed=3):
    rng = np.random.RandomState(seed)
    x = rng.uniform(0, L, (n_envs, N))
    v = rng.normal(0, 1, (n_envs, N)) + 3.0 * (rng.uniform(size=(n_envs, N)) < 0.2)
    return x, v


def _snapshot(eng, last_call_steps):
    xs, vs = eng.get_state()
    rho, _ = eng.get_density_fixed()
    n, E = eng.get_fields()
    return xs, vs, rho, n, E, eng.get_diag(), eng.get_trace(last_call_steps)


def _run(coop, N, M, n_envs, calls, ext=None, coeffs=None, precision="f64"):
    """calls: list of n_steps per step call (coeffs are consumed in order)."""
    eng = _engine(N, M, n_envs, precision, with_basis=coeffs is not None)
    eng.set_coop(coop)
    on, workers = eng.coop
    assert on == (coop == "on")
    x, v = _state(N, n_envs)
    eng.set_state(x, v)
    l0, s = eng.launch_count(), 0
    for k in calls:
        eng.step_coeffs(coeffs[s:s + k], k) if coeffs is not None else eng.step_mesh(ext, k)
        s += k
    launches = eng.launch_count() - l0              # (the float32 state read-back below launches conversions)
    out = _snapshot(eng, calls[-1])
    flags = eng.error_flags()
    info = eng.launch_info()
    eng.close()
    return out, launches, flags, info


def _assert_same(a, b, same_grid):
    for p, q in zip(a[:5], b[:5]):                 # x, v, fixed-point density, n, E_mesh
        assert np.array_equal(p, q)
    for p, q in zip(a[5:], b[5:]):                 # diagnostics of the state / of every step of the last call
        if same_grid:
            assert np.array_equal(p, q)
        else:                                      # KE, sum v: per-CTA partials, added over 147 instead of 148 CTAs
            assert np.allclose(p, q, rtol=1e-13, atol=1e-9)


@pytest.mark.parametrize("N,M,n_envs", [(200_003, 4096, 1), (60_000, 1000, 3), (1_000_001, 1024, 1), (20_000, 250, 1), (33, 7, 1), (1, 2, 2)])
def test_coop_step_is_bit_identical(N, M, n_envs):
    rng = np.random.RandomState(1)
    ext = 0.2 * np.sin(2 * np.pi * np.arange(M) / M)[None].repeat(n_envs, 0) + 0.05 * rng.normal(size=(n_envs, M))
    (a, la, fa, ia) = _run("off", N, M, n_envs, [1, 4, 1], ext=ext)
    (b, lb, fb, ib) = _run("on", N, M, n_envs, [1, 4, 1], ext=ext)
    assert fa == 0 and fb == 0
    assert lb == 3                                 # one launch per call
    assert la >= 4 * 6
    _assert_same(a, b, same_grid=ib["coop_workers"] == ia["grid_x"])
    # six single-step calls == the same six steps in one launch
    (c, lc, _, _) = _run("on", N, M, n_envs, [6], ext=ext)
    assert lc == 1
    for p, q in zip(b[:6], c[:6]):
        assert np.array_equal(p, q)


def test_coop_step_with_per_step_coefficients_and_trace():
    N, M, B, steps = 150_000, 512, 2, 6
    rng = np.random.RandomState(2)
    coeffs = rng.uniform(-1, 1, (steps, B, 6))
    (a, _, fa, ia) = _run("off", N, M, B, [steps], coeffs=coeffs)
    (b, lb, fb, ib) = _run("on", N, M, B, [steps], coeffs=coeffs)
    (c, _, _, _) = _run("on", N, M, B, [2, 1, 3], coeffs=coeffs)
    assert fa == 0 and fb == 0 and lb == 1
    same = ib["coop_workers"] == ia["grid_x"]
    _assert_same(a, b, same)
    assert a[6].shape == (steps, B, a[5].shape[-1])          # reward / input energy of every step of the call
    for p, q in zip(b[:6], c[:6]):
        assert np.array_equal(p, q)
    assert np.array_equal(b[6][-3:], c[6])                   # the trace of the last call = its three steps


def test_coop_step_float32():
    N, M = 300_001, 1000
    ext = 0.2 * np.cos(2 * np.pi * np.arange(M) / M)[None]
    (a, _, fa, ia) = _run("off", N, M, 1, [3, 2], ext=ext, precision="f32")
    (b, lb, fb, ib) = _run("on", N, M, 1, [3, 2], ext=ext, precision="f32")
    assert fa == 0 and fb == 0 and lb == 2
    _assert_same(a, b, ib["coop_workers"] == ia["grid_x"])


def test_coop_and_kernel_per_pass_calls_interleave_on_one_handle():
    N, M = 400_000, 2048
    ext = 0.1 * np.sin(4 * np.pi * np.arange(M) / M)[None]
    (a, _, _, ia) = _run("off", N, M, 1, [7], ext=ext)
    eng = _engine(N, M, 1)
    x, v = _state(N, 1)
    eng.set_state(x, v)
    for k, mode in ((2, "on"), (1, "off"), (3, "on"), (1, "off")):
        eng.set_coop(mode)
        eng.step_mesh(ext, k)
    b = _snapshot(eng, 1)
    assert eng.error_flags() == 0
    eng.close()
    for p, q in zip(a[:5], b[:5]):
        assert np.array_equal(p, q)


def test_coop_auto_and_unsupported_flavours():
    import pic_b200
    eng = _engine(100_000, 500, 1)
    assert eng.coop[0]                             # AUTO: multi-step calls of a mid-size env take the cooperative step
    assert eng.launch_info()["gather"] == "shared"
    eng.sample_state("bump-on-tail", seed=1)
    l0 = eng.launch_count(); eng.step_mesh(None, 3); l1 = eng.launch_count(); eng.step_mesh(None, 1)
    assert l1 - l0 == 1 and eng.launch_count() - l1 == 4      # ... one-step calls stay on the kernel-per-pass path
    eng.set_gather("texture")                      # an explicit texture route keeps the kernel-per-pass path
    assert not eng.coop[0]
    with pytest.raises(pic_b200.PicError):
        eng.set_coop("on")
    eng.close()
    eng = _engine(100_000, 500, 1, interpol="TSC")
    assert not eng.coop[0]
    with pytest.raises(pic_b200.PicError):
        eng.set_coop("on")
    eng.close()
