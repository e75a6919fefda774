"""The streaming kernels' two gather routes -- shared-memory table rebuilt in every CTA's prologue vs. one table per
sub-stage read through the texture pipe (pic_set_gather) -- must produce IDENTICAL bits: same block_field instance on
the same integer density, same per-particle arithmetic.  (The reference has no counterpart: this is the kick's
E[indx_l] / E[indx_r] gather of src/env/util.py:106.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(route, N, M, n_envs, steps, ext=None, coeffs=None, precision="f64", interpol="CIC", multi=False, seed=3):
    import pic_b200
    L = 50.0
    dt = 2 / np.sqrt(N / L) if N > 50000 else 0.05
    eng = pic_b200.Engine(N, M, L, dt, mode="streaming", n_envs=n_envs, precision=precision, interpol=interpol,
                          max_mode=3 if coeffs is not None else 0)
    eng.set_gather(route)
    assert eng.gather == route
    rng = np.random.RandomState(seed)
    x = rng.uniform(0, L, (n_envs, N))
    v = rng.normal(0, 1, (n_envs, N)) + 3.0 * (rng.uniform(size=(n_envs, N)) < 0.2)
    eng.set_state(x, v)
    if coeffs is not None:
        from pic_b200.actuator import E_field
        act = E_field(L, M, 3)
        eng.set_actuator_basis(act.basis_cos, act.basis_sin)
    if multi:                                   # one call: the finalize of step s overlaps the first pass of step s + 1
        eng.step_coeffs(coeffs, steps) if coeffs is not None else eng.step_mesh(ext, steps)
    else:
        for s in range(steps):
            eng.step_coeffs(coeffs[s:s + 1], 1) if coeffs is not None else eng.step_mesh(ext, 1)
    xs, vs = eng.get_state()
    rho, k = eng.get_density_fixed()
    n, E = eng.get_fields()
    diag = eng.get_diag()
    flags = eng.error_flags()
    launches = eng.launch_count()
    eng.close()
    return xs, vs, rho, n, E, diag, flags, launches


@pytest.mark.parametrize("N,M,n_envs", [(200_003, 4096, 1), (60_000, 1000, 3), (1_000_000, 250, 1)])
def test_texture_route_is_bit_identical(N, M, n_envs):
    rng = np.random.RandomState(1)
    ext = 0.2 * np.sin(2 * np.pi * np.arange(M) / M)[None].repeat(n_envs, 0) + 0.05 * rng.normal(size=(n_envs, M))
    a = _run("shared", N, M, n_envs, 4, ext=ext)
    b = _run("texture", N, M, n_envs, 4, ext=ext)
    c = _run("texture:3", N, M, n_envs, 4, ext=ext)     # what AUTO picks for large envs
    for p, q, r in zip(a[:6], b[:6], c[:6]):
        assert np.array_equal(p, q) and np.array_equal(p, r)
    assert a[6] == 0 and b[6] == 0 and c[6] == 0
    assert b[7] > c[7] > a[7]                    # the texture route launches one field kernel per texture pass on top


def test_texture_route_with_coefficients_and_multi_step_call():
    N, M, B, steps = 150_000, 512, 2, 6
    rng = np.random.RandomState(2)
    coeffs = rng.uniform(-1, 1, (steps, B, 6))
    a = _run("shared", N, M, B, steps, coeffs=coeffs, multi=True)
    b = _run("texture", N, M, B, steps, coeffs=coeffs, multi=True)
    c = _run("texture", N, M, B, steps, coeffs=coeffs, multi=False)
    for p, q, r in zip(a[:6], b[:6], c[:6]):
        assert np.array_equal(p, q) and np.array_equal(p, r)


@pytest.mark.parametrize("precision,interpol", [("f32", "CIC"), ("f64", "TSC")])
def test_texture_route_float32_and_tsc(precision, interpol):
    N, M = 120_001, 1000
    ext = 0.2 * np.cos(2 * np.pi * np.arange(M) / M)[None]
    a = _run("shared", N, M, 1, 3, ext=ext, precision=precision, interpol=interpol)
    b = _run("texture", N, M, 1, 3, ext=ext, precision=precision, interpol=interpol)
    for p, q in zip(a[:6], b[:6]):
        assert np.array_equal(p, q)
    assert a[6] == 0 and b[6] == 0


def test_tsc_on_a_4096_cell_mesh_runs_table_less():
    """TSC's two 20-byte-per-cell histograms + a shared gather table exceed one CTA's shared memory at 4096 cells: the
    stage-3 and init kernels then run table-less (texture route), and the result still equals the lean oracle's."""
    import pic_b200
    from oracle import pic_oracle as O
    N, M, L = 200_000, 4096, 50.0
    rng = np.random.RandomState(11)
    x = rng.uniform(0, L, N); v = rng.normal(0, 1, N)
    dt = O.clip_dt(0.1, N, L)
    eng = pic_b200.Engine(N, M, L, dt, mode="streaming", interpol="TSC")
    assert eng.gather == "texture:3"
    eng.set_gather("shared")                      # stage 3 cannot honour it: stays on the texture route
    assert eng.gather == "texture:3"
    eng.set_state(x[None], v[None])
    eng.step_mesh(None, 2)
    xg, vg = eng.get_state()
    p = O.PicParams(N=N, N_mesh=M, n0=1.0, L=L, dt=dt, interpol="TSC")
    xo, vo = x, v
    for _ in range(2):
        o = O.step(xo, vo, p, None)
        xo, vo = o["x"], o["v"]
    assert np.abs(xg[0] - xo).max() < 1e-12 and np.abs(vg[0] - vo).max() < 1e-12
    assert eng.error_flags() == 0
    eng.close()


def test_auto_route_by_size_and_shape():
    import pic_b200
    small = pic_b200.Engine(100_000, 1024, 50.0, 0.01, mode="streaming")
    assert small.gather == "shared"
    big = pic_b200.Engine(1 << 23, 4096, 50.0, 0.001, mode="streaming")
    assert big.gather == "texture:3"             # the stage-3 pass only (measured)
    big.set_tuning(512, 2, 0)                    # no texture variant at this launch shape: back to the shared table
    assert big.gather == "shared"
    with pytest.raises(pic_b200.PicError):
        big.set_gather("texture")
    big.set_tuning(1024, 2, 0)
    assert big.gather == "texture:3"
    small.close(); big.close()


def test_fused_exchange_timeout_in_the_field_table_kernel_leaves_the_state_alone():
    """Fused peer exchange + texture-route stage 3 on ONE GPU with a peer that stops publishing: rank 0 of a world of two
    whose 'peer' buffers are plain local allocations.  The fake peer has 'published' exchanges 1 and 2 (all-zero slots: it
    owns no particles), so the init deposit, stage 1 and stage 2 see complete exchanges; exchange 3 never arrives.  Its
    consumer is the one-CTA field_table_kernel of the stage-3 pass: the wait must end (bounded spin, no hang), raise the
    sticky ERR_COMM_TIMEOUT flag, and the stage-3 pass must leave x and v exactly as stage 2 left them."""
    import torch
    import pic_b200
    N, M, L = 50_000, 512, 50.0
    rng = np.random.RandomState(5)
    x = rng.uniform(0, L, (1, N)); v = rng.normal(0, 1, (1, N))
    eng = pic_b200.Engine(N, M, L, 0.05, mode="streaming")
    words = eng.comm_exchange_words(2)
    exch = [torch.zeros(words, dtype=torch.int64, device="cuda") for _ in range(2)]
    flags = [torch.zeros(16, dtype=torch.int64, device="cuda") for _ in range(2)]
    flags[0][1] = 2                                  # what rank 0 sees of rank 1: exchanges 1 and 2 are there
    torch.cuda.synchronize()
    eng.comm_init_peer(0, 2, [t.data_ptr() for t in exch], [t.data_ptr() for t in flags], words)
    eng.set_gather("texture:3")
    assert eng.gather == "texture:3"
    eng.set_state(x, v)
    assert eng.error_flags() == 0
    eng.step_mesh(None, 1)                           # stage 3 (and the finalize) time out
    eng.sync()
    assert eng.error_flags() & 4                     # ERR_COMM_TIMEOUT
    xs, vs = eng.get_state()
    eng.close()
    ref = pic_b200.Engine(N, M, L, 0.05, mode="streaming")
    ref.set_state(x, v)
    ref.set_stage_actuation(None, None)
    ref.run_stage(1); ref.run_stage(2)
    xr, vr = ref.get_state()
    ref.close()
    assert np.array_equal(xs, xr) and np.array_equal(vs, vr)
