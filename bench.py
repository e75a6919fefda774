"""bench.py -- headline benchmark of the B200-native PIC step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[4], the configuration the metric's roofline target is quoted on): ONE env,
N = 1e9 particles, N_mesh = 4096, L = 50, bump-on-tail (a = 0.2, vb = 3), dt clipped to 2/sqrt(N/L) as the reference
does (src/env/pic.py:71-72), float64.  A "step" is one `PIC.update_state` = one Yoshida-4 env step = 3 fused
push/gather/deposit passes (the drift-only first sub-stage rides along with the previous pass) + field solves.  At N GPUs the 1e9 particles are sharded over the ranks (strong scaling)
with one NCCL all-reduce of the 4096-cell fixed-point density per sub-stage.  Inputs are synthetic: the device-side
sampler draws the reference's bump-on-tail distribution.

One JSON line on stdout (rank 0).  `value` = particle-steps/s with the state resident in HBM; `e2e` = the same
through the reference-facing call with HOST buffers (E_external from pinned host memory in, energies out, every
step); `roofline` = the dominant kernel (kick+drift+deposit pass) against the measured HBM peak; `cpu_baseline` = the
oracle port of the reference's numpy path on this box's host cores.  `--impl reference` times that CPU path alone.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "particle-steps/s (push+gather+deposit+Poisson)"
UNIT = "particle-steps/s"
N_FULL = 1_000_000_000
N_MESH = 4096
L_BOX = 50.0
SETTLE_S = 0.25          # minimum wall time of the untimed warm-up (W steps + extra untimed steps), see run_gpu_arm
FALLBACK_HBM_GBS = 6650.0


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------- CPU baseline
_WARM = False


def _oracle_worker(args):
    """One independent sample env advanced with the faithful oracle; returns (particle_steps, seconds)."""
    global _WARM
    n, mesh, steps, seed = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import pic_oracle as O      # the one place bench.py executes oracle/: the CPU baseline legs
    rng = np.random.RandomState(seed)
    x = rng.uniform(0, L_BOX, n)
    v = rng.normal(size=n) + 3.0 * (rng.uniform(size=n) < 1.0 / 6.0)
    p = O.PicParams(N=n, N_mesh=mesh, n0=1.0, L=L_BOX, dt=O.clip_dt(0.1, n, L_BOX))
    if not _WARM:                           # JIT warm-up outside the timed region, once per process
        O.step(x[:2000].copy(), v[:2000].copy(), O.PicParams(N=2000, N_mesh=mesh, n0=1.0, L=L_BOX, dt=0.01), None,
               faithful=True)
        _WARM = True
    t0 = time.perf_counter()
    for _ in range(steps):
        o = O.step(x, v, p, None, faithful=True)
        x, v = o["x"], o["v"]
    return n * steps, time.perf_counter() - t0


def cpu_baseline(n_sample, steps, procs, pool=None):
    """Throughput of the oracle port (faithful mode: the reference's 8 deposits + 8 periodic solves per step) on
    `procs` host processes, each advancing its own sample env of n_sample particles."""
    jobs = [(n_sample, N_MESH, steps, 100 + i) for i in range(procs)]
    t0 = time.perf_counter()
    if procs == 1:
        res = [_oracle_worker(jobs[0])]
    elif pool is not None:
        res = pool.map(_oracle_worker, jobs, chunksize=1)
    else:
        import multiprocessing as mp
        with mp.get_context("spawn").Pool(procs) as p:
            res = p.map(_oracle_worker, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    total = sum(r[0] for r in res)
    slowest = max(r[1] for r in res)
    return total / slowest, slowest, wall


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    procs = max(1, min(os.cpu_count() or 1, 64))
    n_sample = 1_000_000
    per_step, vals = [], []
    pool = mp.get_context("spawn").Pool(procs) if procs > 1 else None
    try:
        for _ in range(max(1, args.warmup)):
            cpu_baseline(n_sample, 1, procs, pool)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            v, slow, _ = cpu_baseline(n_sample, 1, procs, pool)
            vals.append(v); per_step.append(slow)
        wall = time.perf_counter() - t0
    finally:
        if pool is not None:
            pool.close(); pool.join()
    value = float(np.mean(vals))
    sample = ("oracle port (faithful: 8 deposits + 8 Thomas/Sherman-Morrison solves per step) of the reference's "
              "numpy path, %d processes x one independent env of %d particles, N_mesh=%d, 1 step per timed step; the "
              "reference itself is single-threaded Python and cannot run 1e9 particles" % (procs, n_sample, N_MESH))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(per_step)), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "large-N single env: 1e9 particles, 4096 cells (bounded CPU sample)", "n_particles": N_FULL,
                   "n_mesh": N_MESH, "L": L_BOX, "sample_particles_per_process": n_sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING a timed region (NVML every 10 ms; nvidia-smi as fallback)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None
        self.max_mhz = None
        self._nvml = None
        self.mode = os.environ.get("PIC_BENCH_CLOCKS", "on")          # experiment switch: on | off
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else index
            self._dev = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._dev, pynvml.NVML_CLOCK_SM))
            self._nvml = pynvml
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        sm = float(n.nvmlDeviceGetClockInfo(self._dev, n.NVML_CLOCK_SM))
        pw = n.nvmlDeviceGetPowerUsage(self._dev) / 1000.0
        r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._dev)
        flags = [bool(r & n.nvmlClocksThrottleReasonHwSlowdown), bool(r & n.nvmlClocksThrottleReasonHwThermalSlowdown),
                 bool(r & n.nvmlClocksThrottleReasonSwThermalSlowdown), bool(r & n.nvmlClocksThrottleReasonSwPowerCap)]
        self.rows.append((sm, pw, flags))

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout
        p = [t.strip() for t in out.strip().split(",")]
        if len(p) >= 7:
            self.max_mhz = float(p[1])
            self.rows.append((float(p[0]), float(p[2]), [t.lower().startswith("active") for t in p[3:7]]))

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.01 if self._nvml else 0.1)

    def __enter__(self):
        if self.mode == "off":
            self._t = None
            return self
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t is not None:
            self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(r[2][i] for r in self.rows)]
        return {"sm_mhz": float(np.median([r[0] for r in self.rows])), "sm_max_mhz": self.max_mhz,
                "power_w_max": max(r[1] for r in self.rows), "samples": len(self.rows), "reasons": reasons,
                "source": "nvml" if self._nvml else "nvidia-smi"}


# ---------------------------------------------------------------------------------------------- GPU arm
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    import pic_b200
    from pic_b200 import _lib as PL

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    N = int(args.particles)
    hbm_peak, peak_src = measured_peaks()

    sim = pic_b200.ShardedPIC(N, N_MESH, 1.0, L_BOX, 0.1, rank=rank, world_size=world, device=local,
                              collective=args.collective, deposit=args.deposit)
    eng = sim.engine
    if args.threads:
        eng.set_tuning(args.threads, args.unroll, args.ctas)
    sim.sample_state("bump-on-tail", a=0.2, v0=3.0, sigma=1.0, A=0.1, n_mode=2, seed=42)
    info = eng.launch_info()
    N_local = sim.N_local
    sim_dt, sim_collective = sim.dt, sim.collective

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # ---- value: K steps, state resident in HBM, device-timed, max over ranks
    for _ in range(args.warmup):
        eng.step_mesh_device(None, 1)
    # With many GPUs a step is ~2 ms and W steps are over before clocks and NCCL channels have settled: keep stepping,
    # untimed, until the warm-up has lasted ~SETTLE_S.  The count is derived from one timed step (max over ranks) so
    # that every rank runs the same number of collectives.
    torch.cuda.synchronize()
    t_w = time.perf_counter()
    eng.step_mesh_device(None, 1)
    torch.cuda.synchronize()
    one_step_s = max_over_ranks((time.perf_counter() - t_w) * 1e3) * 1e-3
    settle_steps = 1 + int(min(100, max(0, np.ceil(SETTLE_S / max(one_step_s, 1e-4)) - args.warmup - 1)))
    for _ in range(settle_steps - 1):
        eng.step_mesh_device(None, 1)
    barrier()
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        for _ in range(args.steps):
            eng.step_mesh_device(None, 1)
        e1.record()
        barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = eng.launch_count() - l0
    ms_per_step = ms_total / args.steps
    value = N * args.steps / (ms_total * 1e-3)
    clocks = clk.summary()

    # ---- e2e: the reference-facing call with host buffers, every step:
    #      E_external (N_mesh float64, pinned host) -> update_state -> energies back to the host (sync)
    ext_host = torch.zeros(N_MESH, dtype=torch.float64).pin_memory()
    ext_host += 0.05 * torch.sin(2 * np.pi * torch.arange(N_MESH, dtype=torch.float64) / N_MESH)
    diag_bytes = PL.DIAG_N * 8
    for _ in range(max(1, args.warmup // 2)):
        eng.step_mesh_ptr(ext_host.data_ptr(), 1)
        eng.get_diag()
    barrier()
    t0 = time.perf_counter()
    energies = []
    for _ in range(args.steps):
        eng.step_mesh_ptr(ext_host.data_ptr(), 1)        # H2D copy of this step's input inside the call
        d = eng.get_diag()[0]                           # D2H read of this step's result (synchronises)
        energies.append(float(d[PL.DIAG_KE] + d[PL.DIAG_PE_MESH] * N / L_BOX))
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    e2e_value = N * args.steps / e2e_s

    # ---- for context: what a step costs when the caller insists on round-tripping the whole particle state through
    #      host memory every step (pinned buffers, PCIe).  This is why the env state is device-resident.
    roundtrip = None
    if not args.no_roundtrip and world == 1:
        Nr = 50_000_000
        er = pic_b200.Engine(Nr, N_MESH, L_BOX, 2 / np.sqrt(Nr / L_BOX), mode="streaming", device=local)
        er.sample_state("bump-on-tail", seed=1)
        xh = torch.empty(Nr, dtype=torch.float64).pin_memory()
        vh = torch.empty(Nr, dtype=torch.float64).pin_memory()
        er.get_state_into(xh.data_ptr(), vh.data_ptr())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(2):
            er.set_state_ptr(xh.data_ptr(), vh.data_ptr())                          # H2D of x, v (+ field build)
            er.step_mesh_ptr(ext_host.data_ptr(), 1)
            er.get_state_into(xh.data_ptr(), vh.data_ptr())                         # D2H of x, v (synchronises)
        dt_rt = (time.perf_counter() - t0) / 2
        roundtrip = {"value": Nr / dt_rt, "unit": UNIT, "n_particles": Nr, "h2d_bytes_per_step": 16 * Nr + N_MESH * 8,
                     "d2h_bytes_per_step": 16 * Nr, "s_per_step": dt_rt,
                     "api": "pic_set_state(host x, v) + pic_step_mesh(host E_external) + pic_get_state(host x, v) every step"}
        er.close()
        del xh, vh

    # ---- roofline: the dominant kernel (kick + drift + deposit pass, stages 1-3) timed alone with CUDA events
    stages = (1, 2, 3, 4)                               # kick, kick, final (+ stage 0 of the next step), finalize
    stage_ms = np.zeros(len(stages))
    reps = max(2, min(args.steps, 10))
    evs = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in stages] for _ in range(reps)]
    eng.set_stage_actuation(None, None)
    barrier()
    for r in range(reps):
        for k, st in enumerate(stages):
            evs[r][k][0].record()
            eng.run_stage(st)
            evs[r][k][1].record()
    barrier()
    for r in range(reps):
        for k in range(len(stages)):
            stage_ms[k] += evs[r][k][0].elapsed_time(evs[r][k][1]) / reps
    kick_ms = float(np.mean(stage_ms[0:2]))
    alg_bytes = 32.0 * N_local                          # read x,v + write x,v, float64 (DESIGN.md "Roofline")
    achieved = alg_bytes / (kick_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:                                                # DRAM bytes of the same kernel from the committed ncu capture
        with open(os.path.join(ROOT, "profiles", "traffic_r01.json")) as f:
            tj = json.load(f)
        traffic, traffic_src = tj["traffic_bytes_per_particle"] * N_local, tj["source"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "push_stream_kernel<MODE_KICK> (2 of the 3 passes of a step)", "achieved": achieved,
                "peak": hbm_peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kick_ms,
                "stage_ms": [float(s) for s in stage_ms],
                "stage_names": ["kick (32 B)", "kick (32 B)", "final + next stage-0 deposit (32 B)", "field finalize"],
                "bytes_per_particle_step": 96,
                "step_frac_of_hbm": (96.0 * N_local / (float(stage_ms.sum()) * 1e-3) / 1e9) / hbm_peak}
    flags = eng.error_flags()

    # ---- CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v1, slow, _ = cpu_baseline(2_000_000, 16, 1)
        cpu = {"value": v1, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "oracle port (faithful restatement of the reference's numpy path: 8 deposits + 8 periodic "
                         "solves per step), one env of 2e6 particles, N_mesh=4096, 16 steps, %.1f s; the reference is "
                         "single-threaded and cannot hold 1e9 particles" % slow}

    # ---- batched-env companion number (BASELINE configs[3]); env-sharded, no communication
    batched = None
    if not args.no_batched:
        B = 4096
        lo, hi = pic_b200.shard_range(B, rank, world)
        bp = pic_b200.Engine(5000, 250, L_BOX, 0.05, n_envs=hi - lo, mode="resident", deposit="split32", max_mode=3,
                             device=local)
        act = pic_b200.E_field(L_BOX, 250, 3)
        bp.set_actuator_basis(act.basis_cos, act.basis_sin)
        bp.sample_state("bump-on-tail", seed=7, n_global=5000, env_offset=lo)
        T = 10
        coeffs = torch.rand(T, hi - lo, 6, dtype=torch.float64, device=dev) * 2 - 1
        for _ in range(10):
            bp.step_coeffs_device(coeffs.data_ptr(), T)
        barrier()
        reps = 60                                           # 600 env steps per env: long enough for sustained clocks
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as bclk:
            b0.record()
            for _ in range(reps):
                bp.step_coeffs_device(coeffs.data_ptr(), T)
            b1.record()
            barrier()
        bms = max_over_ranks(b0.elapsed_time(b1)) / (reps * T)
        binfo = bp.launch_info()
        batched = {"workload": "4096 envs x (N=5000, N_mesh=250, dt=0.05, 6 actuator coefficients per env per step)",
                   "env_steps_per_s": B / (bms * 1e-3), "particle_steps_per_s": B * 5000 / (bms * 1e-3),
                   "ms_per_batched_step": bms, "hbm_frac": (32.0 * 5000 * (hi - lo) / (bms * 1e-3) / 1e9) / hbm_peak,
                   "launch": binfo, "clocks": bclk.summary(),
                   "note": "one CTA per env, particle state in shared memory; bound by instruction issue / shared "
                           "atomics (and the SM clock under the power cap), not HBM"}
        bp.close()

    # ---- the separate float32 mode (own tolerance: tests/test_gpu_f32.py), same workload, device-timed
    fp32 = None
    if not args.no_fp32:
        del sim, eng
        torch.cuda.empty_cache()
        s32 = pic_b200.ShardedPIC(N, N_MESH, 1.0, L_BOX, 0.1, rank=rank, world_size=world, device=local,
                                  collective=args.collective, precision="f32")
        s32.sample_state("bump-on-tail", a=0.2, v0=3.0, sigma=1.0, A=0.1, n_mode=2, seed=42)
        for _ in range(3):
            s32.engine.step_mesh_device(None, 1)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            s32.engine.step_mesh_device(None, 1)
        f1.record()
        barrier()
        fms = max_over_ranks(f0.elapsed_time(f1)) / args.steps
        fp32 = {"value": N / (fms * 1e-3), "unit": UNIT, "ms_per_step": fms, "bytes_per_particle_step": 48,
                "step_frac_of_hbm": (48.0 * s32.N_local / (fms * 1e-3) / 1e9) / hbm_peak,
                "tolerance": "per step |dx| <= 2e-5, |dv| <= 1e-5, PE rel 1e-4 vs the float64 reference; indices "
                             "bit-exact vs the float32 restatement (tests/test_gpu_f32.py)",
                "error_flags": int(s32.engine.error_flags())}
        s32.engine.close()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "large-N single env: %.3g particles, %d cells, bump-on-tail, particle-sharded over %d GPU(s)"
                                   % (N, N_MESH, world),
                       "n_particles": N, "n_mesh": N_MESH, "L": L_BOX, "dt": sim_dt, "parallelism": "particle-shard x%d" % world,
                       "collective": sim_collective,
                       "l2_policy": "inputs (16 B x %.3g particles per rank) exceed the 126 MB L2" % N_local,
                       "launch": info, "extra_untimed_warmup_steps": settle_steps},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": N_MESH * 8, "d2h_bytes_per_step": diag_bytes,
                    "api": "pic_step_mesh(host E_external) + pic_get_diag per step (what PIC.update_state + "
                           "PIC.get_energy do); particle state stays resident on the device"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "batched": batched,
            "e2e_state_roundtrip": roundtrip,
            "fp32_mode": fp32,
            "error_flags": int(flags),
            "energy_last": energies[-1] if energies else None,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--particles", type=float, default=N_FULL)
    ap.add_argument("--deposit", default="split32")
    ap.add_argument("--threads", type=int, default=1024)
    ap.add_argument("--unroll", type=int, default=2)
    ap.add_argument("--ctas", type=int, default=0)
    ap.add_argument("--collective", default="nccl", choices=["fused", "nccl"],
                    help="density exchange of the particle-sharded mode: fused peer-memory exchange or ncclAllReduce")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-batched", action="store_true")
    ap.add_argument("--no-roundtrip", action="store_true")
    ap.add_argument("--no-fp32", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3                                   # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
