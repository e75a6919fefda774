"""Places where a duck-type could silently diverge from the reference class (`src/env/pic.py`), each made loud or
faithful: in-place writes to `.x/.v`, `update_params(dt=...)` / `sim.dt = ...`, device error flags, writes through
the zero-copy views."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pic_oracle as O  # noqa: E402  (checker only)


def _sim(dt=0.1, **kw):
    from pic_b200 import PIC
    from pic_b200.dist import BumpOnTail
    np.random.seed(42)
    dist = BumpOnTail(a=0.2, v0=3.0, sigma=1.0, n_samples=5000, L=50.0)
    return PIC(N=5000, N_mesh=250, n0=1.0, L=50.0, dt=dt, tmin=0.0, tmax=50.0, gamma=5.0, A=0.1, n_mode=2,
               interpol="CIC", init_dist=dist, **kw)


@pytest.mark.parametrize("mode", ["resident", "streaming"])
def test_dt_change_takes_effect_like_the_reference(mode):
    """pic.py:79-82 + :133: update_params(dt=...) (or sim.dt = ...) changes the step size of the NEXT update_state.
    Here the device handle -- and the pre-deposited first sub-stage, which has c0*dt baked in -- is rebuilt."""
    sim = _sim(mode=mode)
    x0, v0 = sim.x[:, 0].copy(), sim.v[:, 0].copy()
    sim.update_state(None)
    sim.update_params(dt=0.04, gamma=None)
    assert sim.dt == 0.04
    sim.update_state(None)
    sim.dt = 0.07
    sim.update_state(None)
    xo, vo = x0, v0
    for dt in (0.1, 0.04, 0.07):
        o = O.step(xo, vo, O.PicParams(N=5000, N_mesh=250, n0=1.0, L=50.0, dt=dt), None)
        xo, vo = o["x"], o["v"]
    assert np.abs(sim.x[:, 0] - xo).max() < 1e-12 and np.abs(sim.v[:, 0] - vo).max() < 1e-12
    assert np.abs(sim.E_mesh[:, 0] - o["E_mesh"]).max() < 1e-12
    with pytest.raises(ValueError):
        sim.update_params(N_mesh=500)          # the reference would keep a stale dx/grad/laplacian: refused


def test_state_arrays_are_read_only_and_assignable():
    """In the reference `sim.x` IS the state; here it is a cached host copy, so an in-place write must fail loudly
    instead of being dropped, while assignment (`sim.x = ...`) uploads and rebuilds the fields."""
    sim = _sim()
    with pytest.raises(ValueError):
        sim.x[:] = 0.0
    with pytest.raises(ValueError):
        sim.v[3, 0] = 1.0
    xc = sim.x.copy()                           # what the runners do (run_wo_oc.py:116)
    xc[0, 0] = 1.0
    new_x = np.mod(sim.x + 0.3, sim.L)
    sim.x = new_x
    assert np.array_equal(sim.x, new_x)
    o = O.init_fields(new_x[:, 0], O.PicParams(N=5000, N_mesh=250, n0=1.0, L=50.0, dt=0.1))
    assert np.abs(sim.E_mesh[:, 0] - o["E_mesh"]).max() < 1e-12
    s = sim.get_state()
    s[0, 0] = -1.0                              # get_state returns a fresh writable copy (pic.py:165-167)
    assert sim.x[0, 0] == new_x[0, 0]


def test_device_error_flags_raise_in_the_pic_class():
    """np.bincount raises in the reference when a NaN / out-of-range index reaches the deposit; the device clamps,
    flags and continues, and the PIC class turns the flag into an exception on the next read."""
    from pic_b200 import PicDeviceError
    sim = _sim()
    x = sim.x.copy()
    x[7, 0] = np.nan
    sim.set_state(x, sim.v)
    with pytest.raises(PicDeviceError) as e:
        sim.get_energy()
    assert e.value.flags & 2
    sim.engine.clear_error_flags()
    x[7, 0] = 1.0
    sim.set_state(x, sim.v)
    sim.update_state(None)
    assert np.isfinite(sim.get_energy())


@pytest.mark.parametrize("mode", ["resident", "streaming"])
def test_write_through_views_then_refresh(mode):
    """views() are read-only by contract; a caller that writes particles through them calls refresh_fields(), after
    which the env behaves exactly as if set_state had been called with those particles."""
    import torch
    from pic_b200 import Engine
    N, M, L, dt = 5000, 250, 50.0, 0.05
    rng = np.random.RandomState(3)
    x = rng.uniform(0, L, N); v = rng.normal(size=N)
    a = Engine(N, M, L, dt, mode=mode)
    a.set_state(x[None], v[None])
    a.step_mesh(None, 1)
    x2 = rng.uniform(0, L, N); v2 = rng.normal(size=N)
    vw = a.views()
    torch.as_tensor(vw["x"], device="cuda")[0].copy_(torch.as_tensor(x2, device="cuda"))
    torch.as_tensor(vw["v"], device="cuda")[0].copy_(torch.as_tensor(v2, device="cuda"))
    torch.cuda.synchronize()
    a.refresh_fields()
    a.step_mesh(None, 2)
    b = Engine(N, M, L, dt, mode=mode)
    b.set_state(x2[None], v2[None])
    b.step_mesh(None, 2)
    xa, va = a.get_state(); xb, vb = b.get_state()
    assert np.array_equal(xa, xb) and np.array_equal(va, vb)
    assert np.array_equal(a.get_diag(), b.get_diag())


def test_handles_with_different_meshes_share_the_kernels():
    """Kernel attributes (the opt-in shared-memory limit) belong to the kernel, not to the handle: creating a small env
    after a large one must not break the large one's launches (several envs of different sizes in one process, as an
    RL script with a small training env and a large evaluation env has them)."""
    from pic_b200 import Engine
    L = 50.0
    rng = np.random.RandomState(4)

    def make(N, M, gather):
        e = Engine(N, M, L, 0.01, mode="streaming")
        e.set_gather(gather)
        x = rng.uniform(0, L, (1, N)); v = rng.normal(size=(1, N))
        e.set_state(x, v)
        return e, x, v

    for gather in ("shared", "texture"):
        big, xb, vb = make(300_000, 4096, gather)
        big.step_mesh(None, 1)
        small, xs, vs = make(40_000, 128, gather)        # same kernels, a fraction of the shared memory
        small.step_mesh(None, 2)
        big.step_mesh(None, 1)                           # must still launch with its own (larger) plan
        alone, _, _ = make(300_000, 4096, gather)
        alone.set_state(xb, vb)
        alone.step_mesh(None, 2)
        for p, q in zip(big.get_state(), alone.get_state()):
            assert np.array_equal(p, q)
        assert big.error_flags() == 0 and small.error_flags() == 0
        for e in (big, small, alone):
            e.close()
