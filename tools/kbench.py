"""Kernel tuning sweep (run on the GPU box):  python tools/kbench.py [--n 134217728] [--mesh 4096]

Times every streaming sub-stage kernel (CUDA events on the launching stream) for a grid of launch shapes and
deposit flavours, and the resident batched kernel for its shapes.  Prints achieved algorithmic GB/s
(32 B/particle per pass; stage 3 also deposits stage 0 of the next step) and writes gpurun_out/kbench.json.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pic_b200  # noqa: E402


def time_stages(eng, reps):
    ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(4)]
          for _ in range(reps)]
    for st in (1, 2, 3, 4):
        eng.run_stage(st)           # warm-up step
    torch.cuda.synchronize()
    for r in range(reps):
        for k, st in enumerate((1, 2, 3, 4)):
            ev[r][k][0].record()
            eng.run_stage(st)
            ev[r][k][1].record()
    torch.cuda.synchronize()
    t = np.array([[a.elapsed_time(b) for a, b in row] for row in ev])   # ms
    return t.mean(axis=0), t.min(axis=0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1 << 27)
    ap.add_argument("--mesh", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--skip-stream", action="store_true")
    ap.add_argument("--skip-resident", action="store_true")
    ap.add_argument("--precisions", default="f64")
    ap.add_argument("--sorted", action="store_true", help="also time cell-sorted particles")
    a = ap.parse_args()
    out = {"gpu": torch.cuda.get_device_name(0), "n": a.n, "mesh": a.mesh, "stream": [], "resident": []}
    L = 50.0
    torch.manual_seed(0)

    # copy bandwidth of this box for reference
    src = torch.empty(1 << 28, dtype=torch.float64, device="cuda"); dst = torch.empty_like(src)
    for _ in range(3):
        dst.copy_(src)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        dst.copy_(src)
    e1.record(); torch.cuda.synchronize()
    out["copy_gbs"] = 2 * src.numel() * 8 * 5 / (e0.elapsed_time(e1) * 1e-3) / 1e9
    print("copy bandwidth %.1f GB/s" % out["copy_gbs"], flush=True)
    del src, dst

    if not a.skip_stream:
        N = a.n
        dt = 2 / np.sqrt(N / L)
        x0 = torch.rand(N, dtype=torch.float64, device="cuda") * L
        v0 = torch.randn(N, dtype=torch.float64, device="cuda")
        v0 += 3.0 * (torch.rand(N, device="cuda") < 0.1667)
        orders = [("random", x0, v0)]
        if a.sorted:
            idx = torch.argsort(torch.floor(x0 / (L / a.mesh)))
            orders.append(("sorted", x0[idx].contiguous(), v0[idx].contiguous()))
            del idx
        shapes = [(512, 2, 0), (512, 4, 0), (768, 2, 0), (1024, 1, 0), (1024, 2, 0), (1024, 4, 0)]
        if a.quick:
            shapes = [(512, 2, 0), (1024, 1, 0), (1024, 2, 0)]
        for prec in a.precisions.split(","):
            for oname, xs, vs in orders:
                for dep in ("cas64", "split32"):
                    eng = pic_b200.Engine(N, a.mesh, L, dt, mode="streaming", deposit=dep, precision=prec)
                    if prec == "f32":
                        xin, vin = xs.float(), vs.float()
                    else:
                        xin, vin = xs, vs
                    for (th, un, occ) in shapes:
                        if prec == "f32" and (th, un) not in [(512, 2), (1024, 1), (1024, 2)]:
                            continue
                        try:
                            eng.set_tuning(th, un, occ)
                        except Exception as e:
                            print("skip", th, un, occ, e)
                            continue
                        eng.set_state_device(xin.data_ptr(), vin.data_ptr())
                        mean, best = time_stages(eng, a.reps)
                        esz = 4 if prec == "f32" else 8
                        bytes_ = np.array([4, 4, 4]) * esz * N          # kick, kick, final (+ next stage-0 deposit)
                        gbs = bytes_ / (mean[:3] * 1e-3) / 1e9
                        info = eng.launch_info()
                        rec = dict(prec=prec, order=oname, dep=dep, threads=th, unroll=un, occ_req=occ, grid=info["grid_x"],
                                   ms=mean.tolist(), ms_min=best.tolist(), gbs=gbs.tolist(),
                                   step_ms=float(mean.sum()), part_steps_per_s=N / (mean.sum() * 1e-3))
                        out["stream"].append(rec)
                        print("%s %-6s %-7s th=%4d un=%d occ=%d grid=%4d | ms %s | GB/s %s | step %.3f ms  %.2f Gp-steps/s" % (
                            prec, oname, dep, th, un, occ, info["grid_x"], " ".join("%.3f" % m for m in mean),
                            " ".join("%.0f" % g for g in gbs), mean.sum(), N / (mean.sum() * 1e-3) / 1e9), flush=True)
                    eng.close()
        del x0, v0

    if not a.skip_resident:
        B, N, M = a.envs, 5000, 250
        rng = np.random.RandomState(0)
        x = rng.uniform(0, L, size=(B, N)); v = rng.normal(size=(B, N)) + 3.0 * (rng.uniform(size=(B, N)) < 0.1667)
        coeffs = rng.uniform(-1, 1, size=(20, B, 6))
        act = pic_b200.E_field(L, M, 3)
        for prec in a.precisions.split(","):
            for dep in ("cas64", "split32"):
                eng = pic_b200.Engine(N, M, L, 0.05, n_envs=B, mode="resident", deposit=dep, max_mode=3, precision=prec)
                eng.set_actuator_basis(act.basis_cos, act.basis_sin)
                cdev = torch.as_tensor(coeffs, device="cuda")
                shapes = [(256, 0), (512, 0), (1024, 0)]
                for (th, ppt) in shapes:
                    try:
                        eng.set_tuning(th, ppt, -1)
                    except Exception as e:
                        print("skip", th, ppt, e)
                        continue
                    eng.set_state(x, v)
                    eng.step_coeffs_device(cdev.data_ptr(), 20)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(3):
                        eng.step_coeffs_device(cdev.data_ptr(), 20)
                    e1.record(); torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / 60
                    rec = dict(prec=prec, dep=dep, threads=th, ppt=ppt, ms_per_step=ms, env_steps_per_s=B / (ms * 1e-3),
                               part_steps_per_s=B * N / (ms * 1e-3))
                    out["resident"].append(rec)
                    print("resident %s %-7s th=%4d ppt=%2d | %.3f ms/step  %.2f M env-steps/s  %.2f Gp-steps/s" % (
                        prec, dep, th, ppt, ms, B / ms / 1e3, B * N / ms / 1e6), flush=True)
                eng.close()
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/kbench.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
