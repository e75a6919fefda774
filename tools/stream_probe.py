"""Per-pass timing of the streaming step (run on the GPU box):

    python tools/stream_probe.py [--n 1e9] [--mesh 4096] [--deps split32,cas64] [--reps 10] [--precision f64]

CUDA events around every sub-stage kernel (KICK0 32 B, KICK 24 B, FINAL 32 B per particle at float64, finalize), plus
the device-timed whole step.  Prints achieved algorithmic GB/s per pass and writes gpurun_out/stream_probe.json.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pic_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=float, default=1e9)
ap.add_argument("--mesh", type=int, default=4096)
ap.add_argument("--deps", default="split32")
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--precision", default="f64")
ap.add_argument("--shapes", default="1024x2")
ap.add_argument("--tag", default="")
ap.add_argument("--gather", default="auto")
ap.add_argument("--interpol", default="CIC")
a = ap.parse_args()
N, M, L = int(a.n), a.mesh, 50.0
esz = 4 if a.precision == "f32" else 8
out = []
for dep in a.deps.split(","):
    for shape in a.shapes.split(","):
        th, un = (int(t) for t in shape.split("x"))
        kw = {}
        if "@" in dep:                                   # e.g. split32@32: fixed_bits = 32
            kw["fixed_bits"] = int(dep.split("@")[1])
        eng = pic_b200.Engine(N, M, L, 2 / np.sqrt(N / L), mode="streaming", deposit=dep.split("@")[0],
                              precision=a.precision, interpol=a.interpol, **kw)
        try:
            eng.set_tuning(th, un, 0)
        except Exception as e:
            print("skip", dep, shape, e)
            continue
        eng.set_gather(a.gather)
        eng.sample_state("bump-on-tail", seed=42)
        for _ in range(3):
            eng.step_mesh_device(None, 1)
        torch.cuda.synchronize()
        stages = (1, 2, 3, 4)
        ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in stages]
              for _ in range(a.reps)]
        eng.set_stage_actuation(None, None)
        for r in range(a.reps):
            for k, st in enumerate(stages):
                ev[r][k][0].record(); eng.run_stage(st); ev[r][k][1].record()
        torch.cuda.synchronize()
        t = np.array([[p.elapsed_time(q) for p, q in row] for row in ev]).mean(axis=0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        clk = []
        try:
            import pynvml
            pynvml.nvmlInit()
            hdev = pynvml.nvmlDeviceGetHandleByIndex(0)
        except Exception:
            hdev = None
        e0.record()
        for i in range(a.steps):
            eng.step_mesh_device(None, 1)
            if hdev is not None and i == a.steps - 1:
                torch.cuda.current_stream().synchronize() if False else None
                clk.append((pynvml.nvmlDeviceGetClockInfo(hdev, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(hdev) / 1000.0))
        e1.record(); torch.cuda.synchronize()
        step_ms = e0.elapsed_time(e1) / a.steps
        if clk:
            print("   clocks MHz %s  power W %s" % ([c[0] for c in clk], [int(c[1]) for c in clk]))
        bytes_ = np.array([4, 3, 4]) * esz * N
        gbs = bytes_ / (t[:3] * 1e-3) / 1e9
        info = eng.launch_info()
        rec = dict(dep=dep, shape=shape, gather=eng.gather, precision=a.precision, n=N, mesh=M, stage_ms=t.tolist(), gbs=gbs.tolist(),
                   step_ms=step_ms, gps=N / (step_ms * 1e-3) / 1e9, flags=eng.error_flags(), info=info,
                   diag=eng.get_diag()[0].tolist())
        out.append(rec)
        print("%-10s %-7s %s %s k=%d | ms %s | GB/s %s | step %.3f ms = %.2f G/s = %.3f of %d B roofline | flags %d" % (
            dep, shape, a.precision, eng.gather, info["fixed_bits"], " ".join("%.3f" % m for m in t), " ".join("%.0f" % g for g in gbs),
            step_ms, rec["gps"], 11 * esz * N / (step_ms * 1e-3) / 6536.7e9, 11 * esz, rec["flags"]), flush=True)
        eng.close()
        torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/stream_probe%s.json" % a.tag, "w") as f:
    json.dump(out, f, indent=1)
