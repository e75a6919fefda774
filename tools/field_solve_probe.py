"""Batched resident kernel vs env size and launch shape (run on the GPU box):  python tools/field_solve_probe.py
The N = 32 row is the fixed cost of an env step (four field solves, barriers, diagnostics); the others show which
thread count wins once several envs fit per SM."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pic_b200  # noqa: E402

L, B, M = 50.0, 4096, 250
for prec in ("f64", "f32"):
    for N in (32, 1250, 2500, 5000):
        rng = np.random.RandomState(0)
        x = rng.uniform(0, L, size=(B, N)); v = rng.normal(size=(B, N))
        c = torch.as_tensor(rng.uniform(-1, 1, size=(20, B, 6)), device="cuda")
        for th in (0, 256, 512):
            eng = pic_b200.Engine(N, M, L, 0.05, n_envs=B, mode="resident", max_mode=3, precision=prec)
            act = pic_b200.E_field(L, M, 3); eng.set_actuator_basis(act.basis_cos, act.basis_sin)
            if th:
                eng.set_tuning(th, 0, -1)
            eng.set_state(x, v)
            eng.step_coeffs_device(c.data_ptr(), 20); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                eng.step_coeffs_device(c.data_ptr(), 20)
            e1.record(); torch.cuda.synchronize()
            info = eng.launch_info()
            print("%s N=%5d threads=%4s -> %4d  %.4f ms/step  smem %d" % (prec, N, th or "auto", info["threads"],
                                                                       e0.elapsed_time(e1) / 60, info["smem_bytes"]), flush=True)
            eng.close()
