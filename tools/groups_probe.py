"""Do two half-batches on two streams fill the tail of the wave?  python tools/groups_probe.py [envs] [groups...]"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pic_b200

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
for G in [int(g) for g in (sys.argv[2:] or ["1", "2", "4"])]:
    streams = [torch.cuda.Stream() for _ in range(G)]
    engs, coeffs = [], []
    for g in range(G):
        lo, hi = pic_b200.shard_range(B, g, G)
        e = pic_b200.Engine(5000, 250, 50.0, 0.05, n_envs=hi - lo, mode="resident", deposit="split32", max_mode=3,
                            stream=streams[g].cuda_stream)
        act = pic_b200.E_field(50.0, 250, 3)
        e.set_actuator_basis(act.basis_cos, act.basis_sin)
        e.sample_state("bump-on-tail", seed=7, n_global=5000, env_offset=lo)
        engs.append(e)
        coeffs.append(torch.rand(10, hi - lo, 6, dtype=torch.float64, device="cuda") * 2 - 1)
    torch.cuda.synchronize()
    for _ in range(5):
        for g in range(G):
            engs[g].step_coeffs_device(coeffs[g].data_ptr(), 10)
    torch.cuda.synchronize()
    reps = 60
    t0 = time.perf_counter()
    for _ in range(reps):
        for g in range(G):
            engs[g].step_coeffs_device(coeffs[g].data_ptr(), 10)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("B=%d groups=%d  %.4f ms per batched step  %.3f M env-steps/s" % (B, G, dt / (reps * 10) * 1e3, B * reps * 10 / dt / 1e6), flush=True)
    for e in engs:
        e.close()
