"""Mid-size single envs: the cooperative single-launch step against the kernel-per-pass path, multi-step calls and
one-step calls.    python tools/coop_probe.py [f32]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pic_b200
L = 50.0
prec = "f32" if "f32" in sys.argv[1:] else "f64"
for N, M in ((20000, 250), (100000, 500), (1000000, 1024), (1000000, 4096), (3000000, 4096), (10000000, 4096), (30000000, 4096), (100000000, 4096)):
    row = []
    for coop in ("off", "on"):
        eng = pic_b200.Engine(N, M, L, min(0.05, 2 / np.sqrt(N / L)), n_envs=1, mode="streaming", precision=prec)
        eng.set_coop(coop)
        eng.sample_state("bump-on-tail", seed=1)
        steps = 400 if N <= 1000000 else (100 if N <= 10000000 else 20)
        eng.step_mesh(None, steps // 4); eng.sync()
        t0 = time.perf_counter(); eng.step_mesh(None, steps); eng.sync()
        multi = (time.perf_counter() - t0) / steps * 1e6
        t0 = time.perf_counter()
        for _ in range(steps): eng.step_mesh(None, 1)
        eng.sync()
        single = (time.perf_counter() - t0) / steps * 1e6
        info = eng.launch_info()
        row.append((multi, single, info.get("coop_workers", info["grid_x"]), info["gather"]))
        assert eng.error_flags() == 0
        eng.close()
    (m0, s0, g0, r0), (m1, s1, g1, r1) = row
    print("N=%9d M=%5d  per-pass(grid %3d %-9s): %8.1f us/step in one call, %8.1f per 1-step call | coop(%3d workers): %8.1f / %8.1f  -> %6.2f G vs %6.2f G" % (
        N, M, g0, r0, m0, s0, g1, m1, s1, N / m0 / 1e3, N / m1 / 1e3), flush=True)
