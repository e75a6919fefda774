"""Step time of one env across sizes, from the resident kernel's range into the streaming kernels':  python tools/midsize_latency.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pic_b200  # noqa: E402

L = 50.0
for N, M in ((5000, 250), (12000, 250), (20000, 250), (100000, 500), (1000000, 1024), (10000000, 4096)):
    rng = np.random.RandomState(0)
    x = rng.uniform(0, L, size=(1, N)); v = rng.normal(size=(1, N))
    eng = pic_b200.Engine(N, M, L, pic_b200.clip_dt(0.05, N, L) if hasattr(pic_b200, "clip_dt") else min(0.05, 2 / np.sqrt(N / L)), n_envs=1)
    eng.set_state(x, v)
    eng.step_mesh(None, 50); eng.sync()
    steps = 500 if N <= 1000000 else 100
    t0 = time.perf_counter(); eng.step_mesh(None, steps); eng.sync()
    us = (time.perf_counter() - t0) / steps * 1e6
    info = eng.launch_info()
    print("N=%9d M=%5d %-9s threads=%4d grid=%4d  %9.1f us/step  %7.2f G particle-steps/s" % (
        N, M, info["mode"], info["threads"], info["grid_x"], us, N / us / 1e3), flush=True)
    eng.close()
