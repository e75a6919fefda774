"""Fixed workload for an ncu capture of the cooperative step kernel:  python tools/ncu_target_coop.py [N] [M] [steps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pic_b200
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
eng = pic_b200.Engine(N, M, 50.0, min(0.05, 2 / np.sqrt(N / 50.0)), n_envs=1, mode="streaming")
eng.set_coop("on")
eng.sample_state("bump-on-tail", seed=1)
eng.step_mesh(None, steps); eng.sync()
eng.step_mesh(None, steps); eng.sync()
print("coop ok", eng.launch_info(), eng.get_diag(), eng.error_flags())
