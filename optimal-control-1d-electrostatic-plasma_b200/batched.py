"""`BatchedPIC` -- thousands of independent plasma envs advanced together (BASELINE config 4).

The reference has no batching: its RL loops step ONE `PIC` per process (src/control/rl/ppo.py:279, sac.py:354).
This class keeps the per-env semantics of `PIC.update_state` (every env is bit-identical to the same env run alone
through `PIC`) and adds the env axis: one CTA per env, particle state in shared memory for the whole launch, actuator
coefficients in, (KE, PE_mesh) out, all on the device.

Env sharding over GPUs: envs are independent, so rank r of W simply owns envs [lo, hi) = shard_range(n_envs, r, W);
no communication is involved (DESIGN.md "Multi-GPU").
"""
from typing import Optional

import numpy as np

from . import _lib as L
from .engine import Engine


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous, balanced [lo, hi) slice of n_total items for `rank` of `world` (first n_total % world ranks get
    one extra)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class BatchedPIC:
    def __init__(self, n_envs: int, N: int = 5000, N_mesh: int = 250, n0: float = 1.0, L: float = 50.0,
                 dt: float = 0.05, max_mode: int = 3, *, rank: int = 0, world_size: int = 1, device: int = 0,
                 precision: str = "f64", mode: str = "auto", deposit: str = "auto", alpha: float = 1.0,
                 beta: float = 1.0, n_actions: int = 10, interpol: str = "CIC", groups: Optional[int] = None):
        """groups: the local envs are split into this many contiguous groups, each with its own device handle and CUDA
        stream.  One CTA per env, two CTAs per SM: a batch that is not a multiple of 2 x 148 envs leaves the last wave
        of a launch partly empty (512 envs = 1.73 waves cost 2).  With two groups the launches of one group fill the
        slots the other leaves free -- consecutive calls overlap across groups, every env's own sequence of steps is
        unchanged -- and the throughput is that of full waves (measured: 10.6 -> 12.6 M env-steps/s at 512 envs per
        GPU).  Default: 2 when there are at least 64 local envs, else 1."""
        if interpol == "TSC":
            # src/control/objective.py:24 hard-codes interpol="CIC" when the reference's Reward re-deposits the state, even
            # for a TSC env; the device reward uses the env's own (TSC) field energy.  Say so instead of passing it off as
            # the reference's number; the host-side reference Reward gives the CIC value.
            import warnings
            warnings.warn("BatchedPIC(interpol='TSC'): the on-device reward uses the TSC field energy, the reference's "
                          "Reward always re-deposits with CIC (src/control/objective.py:24)", stacklevel=2)
        self.n_envs_total = int(n_envs)
        self.env_lo, self.env_hi = shard_range(n_envs, rank, world_size)
        self.n_envs = self.env_hi - self.env_lo
        if self.n_envs < 1:
            raise ValueError("rank %d of %d owns no env" % (rank, world_size))
        self.N, self.N_mesh, self.n0, self.L, self.max_mode = int(N), int(N_mesh), n0, L, int(max_mode)
        self.dx = L / N_mesh
        self.dt = dt
        if self.dt > 2 / np.sqrt(self.N / self.L):                  # src/env/pic.py:71-72
            self.dt = 2 / np.sqrt(self.N / self.L)
        if groups is None:
            groups = 2 if self.n_envs >= 64 else 1
        self.groups = max(1, min(int(groups), self.n_envs))
        self._bounds = [shard_range(self.n_envs, g, self.groups) for g in range(self.groups)]      # local env ranges
        self.engines = [Engine(self.N, self.N_mesh, self.L, self.dt, n0=n0, n_envs=hi - lo, precision=precision, mode=mode,
                               deposit=deposit, device=device, max_mode=self.max_mode, interpol=interpol,
                               stream="own" if self.groups > 1 else None) for lo, hi in self._bounds]
        # reward constants, src/control/rl/reward.py:32-33,71-76 -- the reward itself is computed on the device
        self.alpha, self.beta = alpha, beta
        self.r_pe_n = 1.0
        self.r_ie_n = float(n_actions) * L * 0.25
        act = None
        if self.max_mode > 0:
            from .actuator import E_field
            act = E_field(L, N_mesh, max_mode)
        for e in self.engines:
            if act is not None:
                e.set_actuator_basis(act.basis_cos, act.basis_sin)
            e.set_reward(alpha, beta, self.r_pe_n, self.r_ie_n)

    @property
    def engine(self) -> Engine:
        """The device handle (single-group batches).  With several groups use `.engines`."""
        if self.groups != 1:
            raise AttributeError("this batch runs as %d groups: use .engines (one handle per group)" % self.groups)
        return self.engines[0]

    def _split(self, a, axis=0):
        return [np.take(a, range(lo, hi), axis=axis) for lo, hi in self._bounds]

    # ---- state
    def set_state(self, x, v):
        """x, v: (n_envs_local, N) float64."""
        x, v = np.asarray(x, dtype=np.float64), np.asarray(v, dtype=np.float64)
        for e, (lo, hi) in zip(self.engines, self._bounds):
            e.set_state(x[lo:hi], v[lo:hi])

    def reset_from_sampler(self, make_dist, A: float = 0.1, n_mode: int = 2, seed: Optional[int] = 42):
        """Fills every local env from a host sampler: env e (global index) is seeded with seed + e, sampled by
        `make_dist()` (a `BumpOnTail` / `TwoStream` factory) and perturbed as src/env/pic.py:68."""
        xs = np.empty((self.n_envs, self.N)); vs = np.empty((self.n_envs, self.N))
        for i in range(self.n_envs):
            if seed is not None:
                np.random.seed(seed + self.env_lo + i)
            d = make_dist()
            x, v = d.get_sample()
            vs[i] = v * (1 + A * np.sin(2 * np.pi * n_mode * x / self.L))
            xs[i] = x
        self.set_state(xs, vs)
        return xs, vs

    def reset(self, kind: str = "bump-on-tail", a: float = 0.2, v0: float = 3.0, sigma: float = 1.0, A: float = 0.1,
              n_mode: int = 2, seed: int = 42):
        """`PIC.reinit()` for every env at once, entirely on the device: env e (global index) draws its own Philox
        stream of the reference's distribution (src/env/dist.py) followed by the perturbation of pic.py:68."""
        for e, (lo, hi) in zip(self.engines, self._bounds):
            e.sample_state(kind, a=a, v0=v0, sigma=sigma, A=A, n_mode=n_mode, seed=seed, n_global=self.N,
                           env_offset=self.env_lo + lo)
        return self.observe()

    def observe(self):
        """Zero-copy observation for a device-side policy: dict of torch CUDA tensors aliasing the env state --
        x, v of shape (n_envs, N) in `get_state()` order, plus the per-env diagnostics (n_envs, 6).  With several
        groups: a list of such dicts, one per group (consecutive env ranges, `group_bounds()`).
        Read-only by contract (Engine.views); after writing particles through them call `refresh()`."""
        import torch
        out = []
        for e in self.engines:
            e.sync()
            vw = e.views()
            dev = "cuda:%d" % e.device
            out.append({k: torch.as_tensor(vw[k], device=dev) for k in ("x", "v", "diag", "E_mesh", "n")})
        return out[0] if self.groups == 1 else out

    def group_bounds(self):
        """[(lo, hi)] local env ranges of the groups."""
        return list(self._bounds)

    def refresh(self):
        """Rebuild every env's density / field / diagnostics from the particle arrays as they are now on the device
        (needed after in-place writes through `observe()` / `views()`)."""
        for e in self.engines:
            e.refresh_fields()

    def get_state(self):
        """(n_envs_local, 2N): row e is `PIC.get_state()` of env e flattened (x then v)."""
        parts = []
        for e in self.engines:
            x, v = e.get_state()
            parts.append(np.concatenate([x, v], axis=1))
        return np.concatenate(parts, axis=0)

    # ---- stepping
    def step(self, actions=None, n_steps: int = 1):
        """Advance every env by n_steps.  actions: None (no control), (n_envs, 2m) held for all n_steps, or
        (n_steps, n_envs, 2m).  Returns dict with per-step `pe_mesh`, `ke` (n_steps, n_envs) and the reference's
        reward for each transition (computed on the PRE-step state, ddpg.py:455).  All groups are enqueued before any
        result is read back."""
        a = None
        if actions is not None:
            a = np.asarray(actions, dtype=np.float64)
            if a.ndim == 2:
                a = np.broadcast_to(a, (n_steps,) + a.shape)
        for e, (lo, hi) in zip(self.engines, self._bounds):
            if a is None:
                e.step_mesh(None, n_steps)
            else:
                e.step_coeffs(np.ascontiguousarray(a[:, lo:hi, :]), n_steps)
        trs, modes = [], []
        for e in self.engines:
            trs.append(e.get_trace(n_steps))
            e.check_errors()    # a flagged step (out-of-range / non-finite particle) raises, as np.bincount would
            if e.n_modes > 0:
                modes.append(e.get_mode_trace(n_steps))
        tr = np.concatenate(trs, axis=1)
        out = {"pe_mesh": tr[:, :, L.DIAG_PE_MESH], "ke": tr[:, :, L.DIAG_KE], "sum_v": tr[:, :, L.DIAG_SUM_V],
               "reward": tr[:, :, L.DIAG_REWARD], "input_energy": tr[:, :, L.DIAG_INPUT_E]}
        if modes:
            out["modes"] = np.concatenate(modes, axis=1)
        return out

    def step_device(self, coeff_ptrs, n_steps: int = 1):
        """Asynchronous step for device-resident loops: coeff_ptrs[g] is the device address of group g's float64
        coefficients (n_steps, group envs, 2m).  Nothing synchronises; the caller orders its producer streams."""
        for e, p in zip(self.engines, coeff_ptrs):
            e.step_coeffs_device(int(p), n_steps)

    def sync(self):
        for e in self.engines:
            e.sync()

    def enable_modes(self, n_modes: Optional[int] = None):
        """Emit the first n_modes Fourier modes of E_mesh every step (default: max_mode), e.g. for the feedback law
        a = (-Re E_k, +Im E_k) of run_feedback.py / the behaviour-cloning target of ddpg.py:429-431."""
        for e in self.engines:
            e.enable_modes(self.max_mode if n_modes is None else n_modes)

    def views(self):
        """Zero-copy device views (Engine.views) of the single group, or a list with one entry per group."""
        vs = [e.views() for e in self.engines]
        return vs[0] if self.groups == 1 else vs

    def close(self):
        for e in self.engines:
            e.close()
