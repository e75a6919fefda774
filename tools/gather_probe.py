"""Gather route (shared table vs texture pipe, per Yoshida stage) against env size:  python tools/gather_probe.py
Whole-step time of one env inside a 50-step device call, for the routes listed."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pic_b200  # noqa: E402

L = 50.0
routes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["shared", "texture:3", "texture:23", "texture"]
for N, M in ((1000000, 1024), (4000000, 4096), (10000000, 4096), (30000000, 4096), (100000000, 4096)):
    for route in routes:
        eng = pic_b200.Engine(N, M, L, min(0.05, 2 / np.sqrt(N / L)), n_envs=1, mode="streaming")
        eng.set_gather(route)
        eng.sample_state("bump-on-tail", seed=1)
        eng.step_mesh(None, 20); eng.sync()
        steps = 200 if N <= 10000000 else 40
        t0 = time.perf_counter(); eng.step_mesh(None, steps); eng.sync()
        us = (time.perf_counter() - t0) / steps * 1e6
        print("N=%10d M=%5d %-11s %10.1f us/step  %7.2f G particle-steps/s" % (N, M, eng.gather, us, N / us / 1e3), flush=True)
        eng.close()
