"""Worker for tests/test_gpu_multi.py (launched with torch.distributed.run, one rank per GPU)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pic_b200  # noqa: E402


def main():
    collective = sys.argv[1]
    out_path = sys.argv[2]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, M, L = 400_003, 1024, 50.0
    rng = np.random.RandomState(11)
    x = rng.uniform(0, L, N)
    v = rng.normal(size=N) + 3.0 * (rng.uniform(size=N) < 0.2)
    ext = 0.2 * np.cos(2 * np.pi * np.arange(M) / M)
    sim = pic_b200.ShardedPIC(N, M, 1.0, L, 0.1, rank=rank, world_size=world, device=local, collective=collective)
    # every pass gathers through the texture pipe on the sharded side (fused exchange: the one-CTA field-table kernel is
    # then the consumer of the peers' slots); the one-GPU run below keeps the shared-memory table: the routes must agree
    # to the bit as well
    sim.engine.set_gather("texture" if collective != "fused" else "texture:13")   # fused: stage 2 through the prologue
    sim.set_state_global(x, v)
    sim.step(ext, 3)
    sim.step(None, 2)
    xl, vl = sim.get_state_local()
    rho, k = sim.engine.get_density_fixed()
    d = sim.diag()
    res = {"ok": True}
    if rank == 0:
        one = pic_b200.Engine(N, M, L, sim.dt, mode="streaming", device=local)
        one.set_state(x[None], v[None])
        one.step_mesh(ext[None], 3)
        one.step_mesh(None, 2)
        x1, v1 = one.get_state()
        rho1, k1 = one.get_density_fixed()
        d1 = one.get_diag()[0]
        res["rho_equal"] = bool(np.array_equal(rho, rho1)) and k == k1
        res["x_equal"] = bool(np.array_equal(xl, x1[0, sim.lo:sim.hi]))
        res["v_equal"] = bool(np.array_equal(vl, v1[0, sim.lo:sim.hi]))
        res["pe_equal"] = bool(d[1] == d1[1])
        res["ke_rel"] = float(abs(d[0] - d1[0]) / d1[0])
        res["sumv_abs"] = float(abs(d[2] - d1[2]))
        res["flags"] = int(sim.engine.error_flags())
    # the device sampler draws the same population whatever the sharding
    s = pic_b200.ShardedPIC(200_000, 512, 1.0, L, 0.05, rank=rank, world_size=world, device=local,
                            collective=collective if collective != "torch" else "nccl")
    s.sample_state("bump-on-tail", seed=5)
    xs, vs = s.get_state_local()
    if rank == 0:
        whole = pic_b200.Engine(200_000, 512, L, s.dt, mode="streaming", device=local)
        whole.sample_state("bump-on-tail", seed=5)
        xw, vw = whole.get_state()
        res["sampler_shard_equal"] = bool(np.array_equal(xs, xw[0, s.lo:s.hi]) and np.array_equal(vs, vw[0, s.lo:s.hi]))
        rs, _ = s.engine.get_density_fixed()
        rw, _ = whole.get_density_fixed()
        res["sampler_rho_equal"] = bool(np.array_equal(rs, rw))
        with open(out_path, "w") as f:
            json.dump(res, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
