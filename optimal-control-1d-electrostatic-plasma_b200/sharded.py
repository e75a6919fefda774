"""`ShardedPIC` -- one large env whose particles are sharded over GPUs (BASELINE config 5).

One process per GPU.  Rank r owns the contiguous particle slice shard_range(N, r, W); the mesh (N_mesh values) is
replicated.  After every Yoshida sub-stage each rank's fused push/deposit kernel leaves its partial density in fixed
point; the partial densities are summed over ranks with ONE NCCL all-reduce of N_mesh uint64 values (4 per env step),
and every rank rebuilds the same field from the same integers in the next kernel's prologue.  Because the sum is an
integer sum, x, v and the fields are bit-identical for every GPU count (tests/test_gpu_multi.py).

Ways to run the collective:
  collective="fused"  no collective library in the step loop: the last CTA of every push kernel writes the rank's
                      partial density into every peer's exchange buffer over NVLink (torch symmetric memory provides
                      the peer mappings) and raises a flag; the next kernel's prologue waits for the flags and sums
                      the slots in rank order.  <= 8 ranks of one node.
                      With an NVLS multicast mapping of the buffers (torch symmetric memory's multicast_ptr) the
                      publish is ONE store per word that the NVSwitch replicates to every rank.
  collective="auto"   "fused" when every rank has the multicast mapping (measured faster than NCCL there), else "nccl".
  collective="nccl"   the C library calls ncclAllReduce itself on the engine's stream (communicator built from a
                      unique id that is broadcast through torch.distributed);
  collective="torch"  sub-stages are driven one by one and `torch.distributed.all_reduce` runs on an int64 view of
                      the density buffer (works with any backend that can reduce CUDA int64 tensors).
"""
from typing import Optional

import os

import numpy as np

from .batched import shard_range
from .engine import Engine, DeviceArray
from . import _lib as L


def broadcast_bytes(payload: Optional[bytes], src: int = 0, group=None) -> bytes:
    """Broadcast a small bytes object from `src` to every rank through torch.distributed (any backend)."""
    import torch.distributed as dist
    box = [payload if dist.get_rank(group) == src else None]
    dist.broadcast_object_list(box, src=src, group=group)
    return box[0]


def allreduce_fixed_density(t, group=None):
    """Sum fixed-point densities (int64 tensor, any device) over ranks in place.  Integer sum => exact."""
    import torch.distributed as dist
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class ShardedPIC:
    def __init__(self, N: int, N_mesh: int = 4096, n0: float = 1.0, L: float = 50.0, dt: float = 0.1, *,
                 rank: int = 0, world_size: int = 1, device: int = 0, precision: str = "f64",
                 deposit: str = "auto", collective: str = "nccl", group=None, max_mode: int = 0, interpol: str = "CIC"):
        self.N, self.N_mesh, self.n0, self.L = int(N), int(N_mesh), n0, L
        self.rank, self.world = int(rank), int(world_size)
        self.lo, self.hi = shard_range(self.N, self.rank, self.world)
        self.N_local = self.hi - self.lo
        self.dt = dt
        if self.dt > 2 / np.sqrt(self.N / self.L):                      # src/env/pic.py:71-72 with the GLOBAL N
            self.dt = 2 / np.sqrt(self.N / self.L)
        self.group = group
        self.collective = collective if self.world > 1 else "none"
        self.multicast = False

        def make_engine():
            return Engine(self.N_local, self.N_mesh, self.L, self.dt, n0=n0, n_particles_total=self.N,
                          precision=precision, mode="streaming", deposit=deposit, device=device, max_mode=max_mode,
                          interpol=interpol)
        self.engine = make_engine()
        if self.collective == "auto":
            # the fused exchange where it was measured faster than ncclAllReduce -- with an NVLS multicast mapping of the
            # exchange buffers (8 GPUs: 500.7 vs 497.1 G particle-steps/s; per-rank peer stores: 494.4) -- else NCCL.
            # Every rank must take the same branch: the outcome is agreed on with a MIN all-reduce.
            self.collective = "fused" if self._try_fused(device, group) else "nccl"
            if self.collective == "nccl":
                self.engine.close()
                self.engine = make_engine()
        if self.collective == "nccl":
            uid = broadcast_bytes(Engine.nccl_unique_id() if self.rank == 0 else None, 0, group)
            self.engine.comm_init_rank(uid, self.rank, self.world)
        elif self.collective == "fused" and not getattr(self, "_fused_ready", False):
            self._init_fused(device, group)
        self._ext_dev = None

    def _try_fused(self, device, group):
        import torch
        import torch.distributed as dist
        ok = 0
        if self.world <= 8:
            try:
                self._init_fused(device, group)
                ok = 1 if self.multicast else 0
            except Exception:                       # no symmetric memory in this torch build, no peer access, ...
                ok = 0
        g = group if group is not None else dist.group.WORLD
        backend = dist.get_backend(g)
        t = torch.tensor([ok], dtype=torch.int32, device=torch.device("cuda", device) if backend == "nccl" else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=g)
        self._fused_ready = bool(int(t.item()))
        return self._fused_ready

    def _init_fused(self, device, group):
        """Peer-mapped exchange buffer + flag array of every rank (torch symmetric memory does the mapping)."""
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        dev = torch.device("cuda", device)
        g = group if group is not None else dist.group.WORLD
        words = self.engine.comm_exchange_words(self.world)
        self._exch = symm_mem.empty(words, dtype=torch.int64, device=dev)
        self._flags = symm_mem.empty(max(self.world, 16), dtype=torch.int64, device=dev)
        self._exch.zero_()
        self._flags.zero_()
        torch.cuda.synchronize(dev)
        he = symm_mem.rendezvous(self._exch, group=g)
        hf = symm_mem.rendezvous(self._flags, group=g)
        self._symm_handles = (he, hf)
        dist.barrier(group=g)                       # every rank has zeroed its flags before anyone can raise one
        self.engine.comm_init_peer(self.rank, self.world, list(he.buffer_ptrs), list(hf.buffer_ptrs), words)
        # NVLS: with a multicast mapping of the exchange buffers the publishing CTA stores every word once and the
        # switch replicates it (PIC_FUSED_MULTICAST=0 keeps the per-rank peer stores)
        mc = int(getattr(he, "multicast_ptr", 0) or 0)
        self.multicast = bool(mc) and os.environ.get("PIC_FUSED_MULTICAST", "1") != "0"
        if self.multicast:
            self.engine.comm_set_multicast(mc)

    # ---- state
    def sample_state(self, kind="bump-on-tail", **kw):
        if self.collective == "torch":
            raise NotImplementedError("use set_state_local with collective='torch'")
        self.engine.sample_state(kind, global_offset=self.lo, n_global=self.N, **kw)

    def set_state_global(self, x, v):
        """Every rank passes the full arrays and keeps its slice."""
        self.set_state_local(np.asarray(x)[self.lo:self.hi], np.asarray(v)[self.lo:self.hi])

    def set_state_local(self, x, v):
        if self.collective == "torch":
            self._set_state_staged(x, v)
        else:
            self.engine.set_state(np.asarray(x).reshape(1, -1), np.asarray(v).reshape(1, -1))

    def get_state_local(self):
        self.engine.check_errors()
        x, v = self.engine.get_state()
        return x[0], v[0]

    # ---- stepping
    def step(self, E_ext=None, n_steps: int = 1):
        if self.collective == "torch":
            for _ in range(n_steps):
                self._step_staged(E_ext)
        else:
            self.engine.step_mesh(None if E_ext is None else np.asarray(E_ext).reshape(1, -1), n_steps)
        # stepping stays asynchronous; errors surface in diag() / energy() / get_state_local() (check_errors())

    def check_errors(self):
        self.engine.check_errors()

    def diag(self):
        # the sticky device flags ride along: an exchange timeout (a stalled peer), an out-of-range or non-finite
        # particle, or a density overflow raises here instead of returning plausible-looking numbers
        d = self.engine.get_diag(check=True)[0]
        if self.collective == "torch":                 # kinetic sums are rank-local in this mode
            import torch
            import torch.distributed as dist
            t = torch.tensor([d[L.DIAG_KE], d[L.DIAG_SUM_V]], dtype=torch.float64,
                             device="cuda:%d" % self.engine.device)
            dist.all_reduce(t, group=self.group)
            d = d.copy()
            d[L.DIAG_KE], d[L.DIAG_SUM_V] = float(t[0]), float(t[1])
        return d

    def energy(self):
        d = self.diag()
        return float(d[L.DIAG_KE] + d[L.DIAG_PE_MESH] * self.N / self.L)

    # ---- collective="torch": sub-stages driven from here
    def _rho_tensor(self, stage):
        """int64 view of the density a sub-stage left behind; stages 3 and -1 leave two (state + next stage 0)."""
        import torch
        ptr = self.engine.stage_density_ptr(stage)
        n = self.N_mesh * (2 if stage in (3, -1) else 1)
        return torch.as_tensor(DeviceArray(ptr, (n,), "<i8", None, self.engine), device="cuda:%d" % self.engine.device)

    def _set_state_staged(self, x, v):
        import torch
        dev = "cuda:%d" % self.engine.device
        xd = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64), device=dev)
        vd = torch.as_tensor(np.ascontiguousarray(v, dtype=np.float64), device=dev)
        if self.engine.precision == "f32":
            xd, vd = xd.float(), vd.float()
        views = self.engine.views()
        torch.as_tensor(views["x"], device=dev)[0].copy_(xd)
        torch.as_tensor(views["v"], device=dev)[0].copy_(vd)
        torch.cuda.synchronize()
        self.engine.run_stage(-1)
        allreduce_fixed_density(self._rho_tensor(-1), self.group)
        self.engine.run_stage(4)

    def _step_staged(self, E_ext):
        import torch
        dev = "cuda:%d" % self.engine.device
        if E_ext is not None:
            self._ext_dev = torch.as_tensor(np.ascontiguousarray(E_ext, dtype=np.float64).reshape(-1), device=dev)
            self.engine.set_stage_actuation(self._ext_dev.data_ptr(), None)
        else:
            self.engine.set_stage_actuation(None, None)
        for st in (1, 2, 3):                 # stage 0 was deposited by the previous stage 3 / init; stage 1 redoes its drift
            self.engine.run_stage(st)
            allreduce_fixed_density(self._rho_tensor(st), self.group)
        self.engine.run_stage(4)
