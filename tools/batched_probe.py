"""Batched resident kernels across launch shapes and batch sizes (run on the GPU box):

    python tools/batched_probe.py [--envs 4096,2048,1024,512,256] [--n 5000] [--mesh 250] [--shapes 512x1,256x2,...]

shape = threads x CTAs-per-env (1 = one CTA per env, 2/4 = thread-block cluster).  Prints ms per batched env step and
env-steps/s, and checks that every shape leaves bit-identical particles."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pic_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", default="4096,2048,1024,512,256")
ap.add_argument("--n", type=int, default=5000)
ap.add_argument("--mesh", type=int, default=250)
ap.add_argument("--modes", type=int, default=3)
ap.add_argument("--shapes", default="512x1,1024x1,256x2,512x2,256x4")
ap.add_argument("--precision", default="f64")
ap.add_argument("--T", type=int, default=10)
ap.add_argument("--reps", type=int, default=30)
a = ap.parse_args()
L = 50.0
out = []
for B in [int(b) for b in a.envs.split(",")]:
    ref = None
    for shape in a.shapes.split(","):
        th, cl = (int(t) for t in shape.split("x"))
        eng = pic_b200.Engine(a.n, a.mesh, L, 0.05, n_envs=B, mode="resident", deposit="split32", max_mode=a.modes,
                              precision=a.precision)
        try:
            eng.set_tuning(th, cl, -1)
        except Exception as e:
            print("B=%5d %-7s skipped: %s" % (B, shape, str(e)[:90]))
            eng.close()
            continue
        act = pic_b200.E_field(L, a.mesh, a.modes)
        eng.set_actuator_basis(act.basis_cos, act.basis_sin)
        eng.sample_state("bump-on-tail", seed=7, n_global=a.n)
        torch.manual_seed(1)
        coeffs = torch.rand(a.T, B, 2 * a.modes, dtype=torch.float64, device="cuda") * 2 - 1
        for _ in range(3):
            eng.step_coeffs_device(coeffs.data_ptr(), a.T)
        torch.cuda.synchronize()
        x, v = eng.get_state()
        if ref is None:
            ref = (x, v)
        same = bool(np.array_equal(x, ref[0]) and np.array_equal(v, ref[1]))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            eng.step_coeffs_device(coeffs.data_ptr(), a.T)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (a.reps * a.T)
        info = eng.launch_info()
        rec = dict(envs=B, shape=shape, ms=ms, env_steps_per_s=B / (ms * 1e-3), smem=info["smem_bytes"], same=same,
                   flags=eng.error_flags())
        out.append(rec)
        print("B=%5d %-7s smem %6d  %.4f ms/step  %7.3f M env-steps/s  identical=%s flags=%d" % (
            B, shape, info["smem_bytes"], ms, B / ms / 1e3, same, rec["flags"]), flush=True)
        eng.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/batched_probe.json", "w"), indent=1)
