"""Live check of the oracle against the UNMODIFIED reference (only where /root/reference exists, i.e. in the build
container; the GPU box runs the golden-vector tests instead)."""
import os
import sys

import numpy as np
import pytest

REF = os.environ.get("PIC_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "src", "env", "pic.py")),
                                reason="reference tree not present")

from oracle import pic_oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def ref():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    try:
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        import src.env.pic as pic
        import src.env.dist as dist
        import src.control.actuator as act
        import src.control.rl.reward as rew
        yield dict(pic=pic, dist=dist, act=act, rew=rew)
    finally:
        sys.path.remove(REF)


@pytest.mark.parametrize("N,M,L,dt,simcase,m", [(3000, 128, 50.0, 0.1, "two-stream", 2), (7001, 200, 50.0, 0.05, "bump-on-tail", 4),
                                                (40000, 400, 50.0, 0.1, "two-stream", 3)])
def test_oracle_tracks_live_reference(ref, N, M, L, dt, simcase, m):
    np.random.seed(123)
    if simcase == "two-stream":
        d = ref["dist"].TwoStream(v0=3.0, sigma=1.0, n_samples=N, L=L)
    else:
        d = ref["dist"].BumpOnTail(a=0.3, v0=4.0, sigma=0.5, n_samples=N, L=L)
    sim = ref["pic"].PIC(N=N, N_mesh=M, n0=1.0, L=L, dt=dt, tmin=0.0, tmax=5.0, gamma=5.0, A=0.05, n_mode=3,
                         interpol="CIC", init_dist=d)
    p = O.PicParams(N=N, N_mesh=M, n0=1.0, L=L, dt=sim.dt)
    assert sim.dt == O.clip_dt(dt, N, L)
    actuator = ref["act"].E_field(L, M, m)
    bc, bs = O.actuator_basis(L, M, m)
    rng = np.random.RandomState(0)
    x, v = sim.x[:, 0].copy(), sim.v[:, 0].copy()
    xf, vf = x.copy(), v.copy()
    for t in range(4):
        c = rng.uniform(-1, 1, 2 * m)
        actuator.update_E(c[:m], c[m:])
        E_ext = actuator.compute_E() if t % 2 == 0 else None
        sim.update_state(E_ext)
        e = None if E_ext is None else O.actuator_field(bc, bs, c[:m], c[m:])
        o = O.step(x, v, p, e, faithful=False)
        of = O.step(xf, vf, p, e, faithful=True)
        x, v, xf, vf = o["x"], o["v"], of["x"], of["v"]
        assert np.array_equal(of["x"], sim.x[:, 0]) and np.array_equal(of["v"], sim.v[:, 0])       # faithful: same bits
        assert np.array_equal(o["indx_l"], sim.indx_l[:, 0])
        assert np.abs(o["x"] - sim.x[:, 0]).max() < 1e-12 and np.abs(o["v"] - sim.v[:, 0]).max() < 1e-12
        assert np.abs(o["E_mesh"] - sim.E_mesh[:, 0]).max() < 1e-11 * max(1.0, np.abs(sim.E_mesh).max())
    assert abs(O.hamiltonian(x, v, p) - sim.get_energy()) < 1e-10 * sim.get_energy()
    r = ref["rew"].Reward(d.get_init_state(), M, L, -25.0, 25.0, 1.0)
    act6 = rng.uniform(-1, 1, 2 * m)
    assert abs(r.compute_reward(sim.get_state(), act6) - O.reward(O.pe_mesh(o["E_mesh"], p.dx), act6, L)) < 1e-10
