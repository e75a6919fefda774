"""fp32-mode streaming workload for ncu:  python tools/ncu_target_f32.py [N]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pic_b200
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 500_000_000
eng = pic_b200.Engine(N, 4096, 50.0, 2 / np.sqrt(N / 50.0), mode="streaming", precision="f32")
eng.sample_state("bump-on-tail", seed=42)
eng.step_mesh_device(None, 3); eng.sync()
print("ok", eng.launch_info(), eng.error_flags())
