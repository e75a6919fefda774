"""Latency of the single small env (BASELINE configs 1-3: N=5000, N_mesh=250) through the reference-facing `PIC` class:
the loop body of run_wo_oc.py:108-122 (update_state + energies + x/v copies + get_state) and of run_ddpg.py:276-313
(get_state -> coefficients -> E_external -> update_state)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pic_b200  # noqa: E402
from pic_b200.dist import BumpOnTail  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    np.random.seed(42)
    dist = BumpOnTail(a=0.2, v0=3.0, sigma=1.0, n_samples=5000, L=50.0)
    sim = pic_b200.PIC(N=5000, N_mesh=250, n0=1.0, L=50.0, dt=0.1, tmin=0.0, tmax=50.0, gamma=5.0, A=0.1, n_mode=2,
                       interpol="CIC", init_dist=dist, max_mode=3)
    act = pic_b200.E_field(50.0, 250, 3)
    out = {}
    for _ in range(50):
        sim.update_state(None)
    t0 = time.perf_counter()
    for _ in range(steps):
        sim.update_state(None)
    sim.engine.sync()
    out["update_state_only_us"] = (time.perf_counter() - t0) / steps * 1e6
    t0 = time.perf_counter()
    for _ in range(steps):                       # run_wo_oc.py loop body without the host-side Reward
        sim.update_state(None)
        E = sim.get_energy(); PE = sim.get_electric_energy()
        xs = sim.x.copy(); vs = sim.v.copy()
        st = sim.get_state()
    out["run_wo_oc_body_us"] = (time.perf_counter() - t0) / steps * 1e6
    rng = np.random.RandomState(0)
    t0 = time.perf_counter()
    for _ in range(steps):                       # run_ddpg.py loop body with a stand-in for the actor
        st = sim.get_state()
        c = rng.uniform(-1, 1, 6)
        act.update_E(c[:3], c[3:])
        sim.update_state(act.compute_E())
        E = sim.get_energy(); PE = sim.get_electric_energy()
    out["run_ddpg_body_us"] = (time.perf_counter() - t0) / steps * 1e6
    t0 = time.perf_counter()
    for _ in range(steps):                       # coefficient fast path, energies only
        c = rng.uniform(-1, 1, 6)
        sim.update_state_coeffs(c[:3], c[3:], act.basis_cos if _ == 0 else None, act.basis_sin if _ == 0 else None)
        PE = sim.get_electric_energy()
    out["coeff_path_energy_only_us"] = (time.perf_counter() - t0) / steps * 1e6
    sim.engine.step_mesh(None, 500)
    sim.engine.sync()
    t0 = time.perf_counter()
    sim.engine.step_mesh(None, 5000)
    sim.engine.sync()
    out["device_only_us_per_step_5000_steps_one_launch"] = (time.perf_counter() - t0) / 5000 * 1e6
    out["launch"] = sim.engine.launch_info()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
