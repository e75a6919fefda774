"""TSC interpolation timings (streaming 5e8 particles / 4096 cells, batched 4096 x 5000 / 250)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pic_b200

L = 50.0
for interpol in ("CIC", "TSC"):
    N = 500_000_000
    eng = pic_b200.Engine(N, 4096, L, 2 / np.sqrt(N / L), mode="streaming", interpol=interpol)
    eng.sample_state("bump-on-tail", seed=1)
    eng.step_mesh_device(None, 3); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.step_mesh_device(None, 10); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("streaming %s: %.3f ms/step  %.2f G particle-steps/s  (%.0f GB/s of 96 B)" % (interpol, ms, N / ms / 1e6, 96 * N / ms / 1e6), flush=True)
    eng.close()
    B = 4096
    bp = pic_b200.Engine(5000, 250, L, 0.05, n_envs=B, mode="resident", max_mode=3, interpol=interpol)
    act = pic_b200.E_field(L, 250, 3); bp.set_actuator_basis(act.basis_cos, act.basis_sin)
    bp.sample_state("bump-on-tail", seed=2, n_global=5000)
    c = torch.rand(10, B, 6, dtype=torch.float64, device="cuda") * 2 - 1
    bp.step_coeffs_device(c.data_ptr(), 10); torch.cuda.synchronize()
    e0.record()
    for _ in range(10): bp.step_coeffs_device(c.data_ptr(), 10)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 100
    print("batched   %s: %.3f ms/step  %.2f M env-steps/s %s" % (interpol, ms, B / ms / 1e3, bp.launch_info()), flush=True)
    bp.close()
