"""Fixed per-step cost of the streaming path: few particles, full-size mesh.  python tools/fixed_cost_probe.py [steps]
Run it plain for us/step, and under `ncu --metrics gpu__time_duration.sum` for the per-kernel durations."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pic_b200  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
for N, M in ((20_000, 4096), (1_000_000, 1024), (1_000_000, 4096), (10_000_000, 4096)):
    eng = pic_b200.Engine(N, M, 50.0, min(0.05, 2 / np.sqrt(N / 50.0)), mode="streaming")
    eng.sample_state("bump-on-tail", seed=3)
    eng.step_mesh_device(None, 20); eng.sync()
    t0 = time.perf_counter(); eng.step_mesh_device(None, steps); eng.sync()
    us = (time.perf_counter() - t0) / steps * 1e6
    print("N=%9d M=%5d grid=%4d  %8.1f us/step  %6.2f G particle-steps/s" % (N, M, eng.launch_info()["grid_x"], us, N / us / 1e3), flush=True)
    eng.close()
