"""SAC-variant batched envs (run_sac.py:33-34,57: N=10000, N_mesh=500, max_mode=5): launch-shape sweep."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pic_b200
L, B = 50.0, 2048
for prec in ("f64", "f32"):
    for th in (0, 256, 512, 1024):
        bp = pic_b200.Engine(10000, 500, L, 0.05, n_envs=B, mode="resident", max_mode=5, precision=prec)
        act = pic_b200.E_field(L, 500, 5); bp.set_actuator_basis(act.basis_cos, act.basis_sin)
        if th:
            try:
                bp.set_tuning(th, 0, -1)
            except Exception as e:
                print(prec, th, "skip", e); continue
        bp.sample_state("bump-on-tail", seed=2, n_global=10000)
        c = torch.rand(10, B, 10, dtype=torch.float64, device="cuda") * 2 - 1
        bp.step_coeffs_device(c.data_ptr(), 10); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): bp.step_coeffs_device(c.data_ptr(), 10)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 50
        print("%s threads=%s: %.3f ms/step  %.2f M env-steps/s  %.1f G particle-steps/s %s" % (
            prec, th or "auto", ms, B / ms / 1e3, B * 10000 / ms / 1e6, bp.launch_info()), flush=True)
        bp.close()
