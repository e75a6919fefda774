"""Generate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container only (needs /root/reference, numba, scipy):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference cannot travel to the GPU box, so its outputs are committed as
fixtures.  Every array below is produced by the reference's own classes
(`src.env.pic.PIC`, `src.env.dist.*`, `src.control.actuator.E_field`,
`src.control.rl.reward.Reward`, `src.env.interpolate.CIC/TSC`); nothing from
this repository takes part.
"""
import os
import sys
import types

import numpy as np

REF = os.environ.get("PIC_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)

from src.env.pic import PIC                                  # noqa: E402  (runs np.random.seed(42))
from src.env.dist import BumpOnTail, TwoStream               # noqa: E402
from src.env.interpolate import TSC                          # noqa: E402
from src.env.util import compute_n                           # noqa: E402
from src.control.actuator import E_field                     # noqa: E402
from src.control.rl.reward import Reward                     # noqa: E402


def make_sim(simcase, N=5000, N_mesh=250, L=50.0, dt=0.1, vb=3.0, vth=1.0, a=0.2, A=0.1, n_mode=2,
             interpol="CIC"):
    np.random.seed(42)   # what importing src/env/pic.py does right before the runner builds its dist
    if simcase == "two-stream":
        dist = TwoStream(v0=vb, sigma=vth, n_samples=N, L=L)
    else:
        dist = BumpOnTail(a=a, v0=vb, sigma=vth, n_samples=N, L=L)
    sim = PIC(N=N, N_mesh=N_mesh, n0=1.0, L=L, dt=dt, tmin=0.0, tmax=50.0, gamma=5.0, A=A, n_mode=n_mode,
              interpol=interpol, init_dist=dist)
    return sim


def snap(sim, tag, out, light=False):
    out[f"{tag}_x"] = sim.x[:, 0].copy()
    out[f"{tag}_v"] = sim.v[:, 0].copy()
    if light:
        return
    out[f"{tag}_n"] = np.asarray(sim.n).copy()
    out[f"{tag}_E_mesh"] = sim.E_mesh[:, 0].copy()
    out[f"{tag}_E"] = sim.E[:, 0].copy()
    out[f"{tag}_indx_l"] = sim.indx_l[:, 0].astype(np.int32)
    if sim.interpol == "TSC":
        out[f"{tag}_indx_m"] = sim.indx_m[:, 0].astype(np.int32)
        out[f"{tag}_weight_m"] = sim.weight_m[:, 0].copy()


def run_case(name, simcase, steps, checkpoints, control=None, light=False, **kw):
    sim = make_sim(simcase, **kw)
    out = {"dt": np.float64(sim.dt), "N": np.int64(sim.N), "N_mesh": np.int64(sim.N_mesh), "L": np.float64(sim.L)}
    snap(sim, "t0", out, light)
    out["raw_x_init"] = sim.init_dist.x_init.copy()      # before the velocity perturbation
    out["raw_v_init"] = sim.init_dist.v_init.copy()
    H = [sim.get_energy()]
    PE = [sim.get_electric_energy()]
    PEm = [0.5 * float(np.sum(sim.E_mesh * sim.E_mesh)) * sim.dx]
    sv = [float(np.sum(sim.v))]
    sx = [float(np.sum(sim.x))]
    rewards = []
    reward_cls = Reward(sim.init_dist.get_init_state(), sim.N_mesh, sim.L, -25.0, 25.0, 1.0)
    coeff_log, eext_log = [], []
    actuator = None
    if control is not None:
        actuator = E_field(sim.L, sim.N_mesh, control["max_mode"])
        out["basis_cos"] = actuator.basis_cos.copy()
        out["basis_sin"] = actuator.basis_sin.copy()
    for t in range(1, steps + 1):
        if control is None:
            sim.update_state(None)
        else:
            coeffs = control["fn"](t - 1)
            m = control["max_mode"]
            state = sim.get_state()
            actuator.update_E(coeffs[:m], coeffs[m:])
            E_ext = actuator.compute_E()
            sim.update_state(E_ext)
            rewards.append(reward_cls.compute_reward(state, coeffs))   # reward on the PRE-step state (ddpg.py:455)
            coeff_log.append(coeffs.copy())
            eext_log.append(E_ext[:, 0].copy())
        H.append(sim.get_energy())
        PE.append(sim.get_electric_energy())
        PEm.append(0.5 * float(np.sum(sim.E_mesh * sim.E_mesh)) * sim.dx)
        sv.append(float(np.sum(sim.v)))
        sx.append(float(np.sum(sim.x)))
        if t in checkpoints:
            snap(sim, f"t{t}", out, light)
    out["H"] = np.array(H)
    out["PE"] = np.array(PE)
    out["PE_mesh"] = np.array(PEm)
    out["sum_v"] = np.array(sv)
    out["sum_x"] = np.array(sx)
    if control is not None:
        out["coeffs"] = np.array(coeff_log)
        out["E_ext"] = np.array(eext_log)
        out["rewards"] = np.array(rewards)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "steps", steps, "H_end", H[-1], "PE_end", PE[-1])


def deposit_edge_cases():
    """CIC / TSC / np.mod on adversarial positions, straight from the reference."""
    out = {}
    for (L, M) in [(50.0, 250), (50.0, 500), (50.0, 4096), (10.0, 64)]:
        dx = L / M
        ks = np.arange(0, M + 1, dtype=np.float64)
        base = ks * dx
        pts = [base, np.nextafter(base, -np.inf), np.nextafter(base, np.inf)]
        pts.append(np.array([0.0, -0.0, 5e-324, -5e-324, -1e-300, -1e-17, -1e-16, L, np.nextafter(L, 0.0),
                             np.nextafter(L, 2 * L), 2 * L - 1e-9, -L + 1e-9, -L, 1.5 * L, -0.5 * L,
                             3.25 * L, -2.75 * L, 7 * L + 0.3, -9 * L - 0.3]))
        rng = np.random.RandomState(7)
        pts.append(rng.uniform(-2 * L, 3 * L, size=4000))
        x_in = np.concatenate(pts)
        xw = np.mod(np.mod(x_in, L), L)
        # the reference would crash in np.bincount if floor(x/dx) == M; keep only legal points (none are dropped
        # for the shipped grids, asserted below)
        il = np.floor(xw / dx).astype(np.int64)
        assert il.max() < M and il.min() >= 0, (L, M)
        N = x_in.shape[0]
        # compute_n (util.py:48) is the entry the hot path uses: np.mod in place, then CIC (which mods again)
        n, i_l, i_r, w_l, w_r = compute_n(x_in.reshape(-1, 1).copy(), dx, M, 1.0, L, N, True, "CIC")
        key = f"L{L:g}_M{M}"
        out[key + "_x"] = x_in
        out[key + "_xw"] = xw
        out[key + "_cic_n"] = n
        out[key + "_cic_il"] = i_l[:, 0].astype(np.int32)
        out[key + "_cic_ir"] = i_r[:, 0].astype(np.int32)
        out[key + "_cic_wl"] = w_l[:, 0]
        out[key + "_cic_wr"] = w_r[:, 0]
        nt, tl, tm, tr, wl, wm, wr = compute_n(x_in.reshape(-1, 1).copy(), dx, M, 1.0, L, N, True, "TSC")
        out[key + "_tsc_n"] = nt
        out[key + "_tsc_im"] = tm[:, 0].astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "deposit_edges.npz"), **out)
    print("deposit_edges written")


if __name__ == "__main__":
    # TSC interpolation (run_wo_oc.py --interpol TSC)
    if os.environ.get("GOLDEN_ONLY", "") in ("", "tsc"):
        run_case("bump_vb3_tsc", "bump-on-tail", 40, {1, 40}, interpol="TSC")
        rs3 = np.random.RandomState(77)
        seq3 = rs3.uniform(-1.0, 1.0, size=(20, 6))
        run_case("twostream_tsc_ctrl", "two-stream", 20, {1, 20}, interpol="TSC",
                 control={"max_mode": 3, "fn": lambda t: seq3[t]})
        if os.environ.get("GOLDEN_ONLY", "") == "tsc":
            sys.exit(0)

    run_case("bump_vb3", "bump-on-tail", 500, {1, 10, 500})
    run_case("twostream_vb3", "two-stream", 500, {1, 10, 500})
    run_case("bump_vb5", "bump-on-tail", 500, {500}, vb=5.0)

    const = np.array([.5, -.25, .125, -.5, .25, -.125])
    run_case("bump_vb3_constctrl", "bump-on-tail", 10, {1, 10},
             control={"max_mode": 3, "fn": lambda t: const})

    rs = np.random.RandomState(1234)
    seq = rs.uniform(-1.25, 1.25, size=(200, 6))
    run_case("bump_vb3_randctrl", "bump-on-tail", 200, {1, 200},
             control={"max_mode": 3, "fn": lambda t: seq[t]})

    rs2 = np.random.RandomState(99)
    seq2 = rs2.uniform(-1.0, 1.0, size=(40, 10))
    run_case("sac_cfg", "bump-on-tail", 40, {1, 40}, N=10000, N_mesh=500, dt=0.05,
             control={"max_mode": 5, "fn": lambda t: seq2[t]})

    # dt clip (pic.py:71-73): N=40000, L=50 -> dt 0.0707
    run_case("clip_dt", "two-stream", 5, {5}, light=True, N=40000, N_mesh=400, dt=0.1)
    deposit_edge_cases()
