"""Batches of envs too large for one CTA: cluster-resident kernel vs the streaming kernels (blockIdx.y = env).
python tools/bigenv_probe.py"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pic_b200

for (B, N, M) in ((256, 40000, 500), (64, 40000, 500), (1024, 20000, 250), (32, 100000, 1000), (512, 13000, 250)):
    for mode in ("resident", "streaming"):
        try:
            eng = pic_b200.Engine(N, M, 50.0, 0.02, n_envs=B, mode=mode, deposit="split32", max_mode=3)
        except Exception as e:
            print("B=%4d N=%6d M=%4d %-9s skipped: %s" % (B, N, M, mode, str(e)[:80])); continue
        act = pic_b200.E_field(50.0, M, 3)
        eng.set_actuator_basis(act.basis_cos, act.basis_sin)
        eng.sample_state("bump-on-tail", seed=7, n_global=N)
        T = 5
        coeffs = torch.rand(T, B, 6, dtype=torch.float64, device="cuda") * 2 - 1
        for _ in range(3):
            eng.step_coeffs_device(coeffs.data_ptr(), T)
        eng.sync()
        t0 = time.perf_counter()
        reps = 10
        for _ in range(reps):
            eng.step_coeffs_device(coeffs.data_ptr(), T)
        eng.sync()
        ms = (time.perf_counter() - t0) / (reps * T) * 1e3
        info = eng.launch_info()
        print("B=%4d N=%6d M=%4d %-9s threads=%4d per_thread/cluster=%d grid=%5d  %.4f ms/step  %.3f M env-steps/s  %.2f G particle-steps/s" % (
            B, N, M, info["mode"], info["threads"], info["per_thread"], info["grid_x"], ms, B / ms / 1e3, B * N / ms / 1e6), flush=True)
        eng.close()
