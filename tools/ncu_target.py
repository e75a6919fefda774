"""Small fixed workload for ncu captures:  python tools/ncu_target.py [stream|resident] [N] [threads] [per_thread] [dep]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pic_b200  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "stream"
L = 50.0
if which == "stream":
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 26
    th = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    pt = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    dep = sys.argv[5] if len(sys.argv) > 5 else "split32"
    gather = sys.argv[6] if len(sys.argv) > 6 else "auto"
    torch.manual_seed(0)
    x = torch.rand(N, dtype=torch.float64, device="cuda") * L
    v = torch.randn(N, dtype=torch.float64, device="cuda") + 3.0 * (torch.rand(N, device="cuda") < 0.1667)
    eng = pic_b200.Engine(N, 4096, L, 2 / np.sqrt(N / L), mode="streaming", deposit=dep)
    eng.set_tuning(th, pt, 0)
    eng.set_gather(gather)
    eng.set_state_device(x.data_ptr(), v.data_ptr())
    eng.step_mesh(None, 3)
    eng.sync()
    print("stream ok", eng.launch_info(), eng.get_diag())
else:
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 1184
    th = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    pt = 0
    dep = sys.argv[5] if len(sys.argv) > 5 else "split32"
    rng = np.random.RandomState(0)
    x = rng.uniform(0, L, size=(B, 5000)); v = rng.normal(size=(B, 5000)) + 3.0 * (rng.uniform(size=(B, 5000)) < 0.1667)
    eng = pic_b200.Engine(5000, 250, L, 0.05, n_envs=B, mode="resident", deposit=dep, max_mode=3)
    act = pic_b200.E_field(L, 250, 3)
    eng.set_actuator_basis(act.basis_cos, act.basis_sin)
    eng.set_tuning(th, pt, -1)
    eng.set_state(x, v)
    eng.step_coeffs(rng.uniform(-1, 1, size=(4, B, 6)), 4)
    eng.step_coeffs(rng.uniform(-1, 1, size=(4, B, 6)), 4)
    eng.sync()
    print("resident ok", eng.launch_info(), eng.get_diag()[0])
