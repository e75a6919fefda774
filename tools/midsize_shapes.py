"""Mid-size single envs (1e6 .. 3e7 particles): step time against the launch shape of the streaming kernels.
    python tools/midsize_shapes.py      (measured: no shape beats 1024 x 2 by more than 3 %; the fixed cost per pass decides)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pic_b200
L = 50.0
for N, M in ((1000000, 1024), (3000000, 4096), (10000000, 4096), (30000000, 4096)):
    for shape in ((1024, 2), (1024, 1), (512, 2), (1024, 4)):
        eng = pic_b200.Engine(N, M, L, min(0.05, 2 / np.sqrt(N / L)), n_envs=1, mode="streaming")
        try:
            eng.set_tuning(shape[0], shape[1], 0)
        except Exception as e:
            print("skip", shape, e); eng.close(); continue
        eng.sample_state("bump-on-tail", seed=1)
        eng.step_mesh(None, 20); eng.sync()
        steps = 200 if N <= 10000000 else 40
        t0 = time.perf_counter(); eng.step_mesh(None, steps); eng.sync()
        us = (time.perf_counter() - t0) / steps * 1e6
        info = eng.launch_info()
        print("N=%9d M=%5d %4dx%d grid=%3d %-10s %8.1f us/step %6.2f G/s" % (N, M, shape[0], shape[1], info["grid_x"], info.get("gather"), us, N / us / 1e3), flush=True)
        eng.close()
