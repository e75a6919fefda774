"""CPU oracle for the 1D electrostatic PIC step  --  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the algorithm that the reference
(ZINZINBIN/Optimal-Control-1D-Electrostatic-Plasma) runs inside
``PIC.update_state``.  It is the *checker* for the CUDA path: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  Nothing under
``optimal-control-1d-electrostatic-plasma_b200/`` imports this module and the
product has no CPU fallback.

Parity pin
----------
The reference ships no tests, golden vectors or known-answer fixtures for this
path, so the oracle is pinned against outputs of the
reference itself: ``tests/golden/make_golden.py`` imports the unmodified
reference from ``/root/reference`` in the build container, runs it and stores
the trajectories in ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
checks this file against those vectors (and, when ``/root/reference`` is
present, ``tests/test_oracle_vs_reference.py`` checks it against the live
reference).  The two notebook growth rates 0.02135 / 0.00557
(``analysis/*.ipynb:51-52``) are the only numbers the reference publishes for
the path; they are reproduced in the golden tests too.

Two flavours of every stage are kept:

* ``faithful=True``  follows the reference call for call (8 deposits and
  8 dense periodic solves per step, Thomas + Sherman-Morrison on the singular
  periodic Laplacian).  This is what ``bench.py`` times as the CPU baseline.
* ``faithful=False`` (lean) drops the calls whose results the reference
  throws away and replaces the dense solve by the algebraically identical
  prefix-sum form of the same 3-point discretisation.  Same particles, same
  indices; fields agree to ~1e-14.  This is what the big parity tests use.

Reference citations are ``file:line`` relative to the reference repo root.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

try:  # the reference JIT-compiles its solver with numba; use it too when present
    from numba import njit as _njit

    _HAVE_NUMBA = True
except Exception:  # pragma: no cover - numba is part of the image
    _HAVE_NUMBA = False

    def _njit(*a, **k):
        def deco(f):
            return f

        return deco


# --------------------------------------------------------------------------
# Yoshida coefficients                         src/env/integration.py:60-69
# --------------------------------------------------------------------------
def yoshida_coefficients() -> Tuple[Tuple[float, ...], Tuple[float, ...]]:
    """(c1..c4), (d0..d3) exactly as the reference evaluates them."""
    phi = 2 ** (1 / 3)
    w0 = (-1) * phi / (2 - phi)
    w1 = 1 / (2 - phi)
    c1 = c4 = 0.5 * w1
    c2 = c3 = 0.5 * (w0 + w1)
    d1 = d3 = w1
    d2 = w0
    return (c1, c2, c3, c4), (0.0, d1, d2, d3)


def clip_dt(dt: float, N: int, L: float) -> float:
    """CFL clip of the time step                        src/env/pic.py:71-72"""
    lim = 2 / np.sqrt(N / L)
    return float(lim) if dt > lim else float(dt)


def deposit_scale(n0: float, L: float, N: int, dx: float) -> float:
    """n0 * L / N / dx evaluated left to right      src/env/interpolate.py:18"""
    return n0 * L / N / dx


# --------------------------------------------------------------------------
# Deposit (CIC / TSC)                            src/env/interpolate.py:4-44
# --------------------------------------------------------------------------
def wrap(x: np.ndarray, L: float) -> np.ndarray:
    """Positions as the deposit sees them: np.mod applied by compute_n
    (src/env/util.py:51) and again by CIC (src/env/interpolate.py:6)."""
    return np.mod(np.mod(x, L), L)


def cic(x: np.ndarray, n0: float, L: float, N: int, N_mesh: int, dx: float):
    """src/env/interpolate.py:4-20 on a flat array (the reference uses (N,1))."""
    x = np.mod(x, L)
    indx_l = np.floor(x / dx).astype(np.int64)
    indx_r = indx_l + 1
    weight_l = (indx_r * dx - x) / dx
    weight_r = (x - indx_l * dx) / dx
    indx_r = np.mod(indx_r, N_mesh)
    n = np.bincount(indx_l, weights=weight_l, minlength=N_mesh)
    n += np.bincount(indx_r, weights=weight_r, minlength=N_mesh)
    n *= n0 * L / N / dx
    return n, indx_l, indx_r, weight_l, weight_r


def tsc(x: np.ndarray, n0: float, L: float, N: int, N_mesh: int, dx: float):
    """src/env/interpolate.py:22-44 (formulas followed literally)."""
    x = np.mod(x, L)
    indx_m = np.floor(x / dx).astype(np.int64)
    dist = (x - indx_m * dx) / dx
    weight_l = 0.5 * (1.5 - dist) ** 2
    weight_m = 0.75 - (dist - 1) ** 2
    weight_r = 0.5 * (dist - 0.5) ** 2
    indx_l = np.mod(indx_m - 1, N_mesh)
    indx_m = np.mod(indx_m, N_mesh)
    indx_r = np.mod(indx_m + 1, N_mesh)
    n = np.bincount(indx_m, weights=weight_m, minlength=N_mesh)
    n += np.bincount(indx_l, weights=weight_l, minlength=N_mesh)
    n += np.bincount(indx_r, weights=weight_r, minlength=N_mesh)
    n *= n0 * L / N / dx
    return n, indx_l, indx_m, indx_r, weight_l, weight_m, weight_r


def compute_n(x: np.ndarray, dx, N_mesh, n0, L, N, interpol: str = "CIC"):
    """src/env/util.py:48-70 -- wraps ``x`` IN PLACE, then CIC (5-tuple) or TSC (7-tuple)."""
    x[:] = np.mod(x, L)
    if interpol == "TSC":
        return tsc(x, n0, L, N, N_mesh, dx)
    return cic(x, n0, L, N, N_mesh, dx)


# --------------------------------------------------------------------------
# Periodic field solve
# --------------------------------------------------------------------------
@_njit(cache=False)
def _thomas_tridiag(lo, di, up, rhs):
    """src/env/solve.py:5-25 restated on the three diagonals (the reference
    walks a dense matrix but only ever touches these entries).
    lo[j] = A[j, j-1], di[j] = A[j, j], up[j] = A[j, j+1]."""
    n = di.shape[0]
    d = di.copy()
    b = rhs.copy()
    xs = np.zeros_like(b)
    for j in range(1, n):
        d[j] = d[j] - lo[j] * up[j - 1] / d[j - 1]
        b[j] = b[j] - lo[j] * b[j - 1] / d[j - 1]
    xs[n - 1] = b[n - 1] / d[n - 1]
    for i in range(n - 2, -1, -1):
        xs[i] = (b[i] - up[i] * xs[i + 1]) / d[i]
    return xs


@_njit(cache=False)
def _periodic_solve(N_mesh, dx, rhs, gamma):
    """src/env/solve.py:27-53 with the Laplacian of src/env/util.py:28-46."""
    inv = 1.0 / dx ** 2          # laplacian /= dx**2 (entries 1.0 and -2.0)
    off = 1.0 / dx ** 2
    dia = -2.0 / dx ** 2
    lo = np.full(N_mesh, off)
    up = np.full(N_mesh, off)
    di = np.full(N_mesh, dia)
    a0n = off                    # A[0,-1]
    an0 = off                    # A[-1,0]
    di[0] -= gamma
    di[N_mesh - 1] -= a0n * an0 / gamma
    u = np.zeros(N_mesh)
    u[0] = gamma
    u[N_mesh - 1] = an0
    v = np.zeros(N_mesh)
    v[0] = 1.0
    v[N_mesh - 1] = a0n / gamma
    x = _thomas_tridiag(lo, di, up, rhs)
    q = _thomas_tridiag(lo, di, up, u)
    x -= q * np.dot(v, x) / (1 + np.dot(v, q))
    return x


def field_dense(n: np.ndarray, n0: float, L: float, N_mesh: int, gamma: float = 5.0):
    """phi and E on the mesh the way the reference computes them:
    src/env/util.py:99-100 (solve, then -grad@phi with the centred periodic
    difference of src/env/util.py:7-26)."""
    dx = L / N_mesh
    phi = _periodic_solve(N_mesh, dx, np.ascontiguousarray(n - n0, dtype=np.float64), gamma)
    # grad row i: (-phi[i-1] + phi[i+1]) / (2 dx); the reference divides the
    # matrix entries (+-1.0) by 2*dx first and then does a dense mat-vec.
    g = 1.0 / (2 * dx)
    E = (-1) * (np.roll(phi, -1) * g + np.roll(phi, 1) * (-g))
    return phi, E


def field_prefix(n: np.ndarray, n0: float, L: float, N_mesh: int):
    """Same discretisation without the singular solve.

    With D_j = phi_{j+1} - phi_j the 3-point equation reads
    D_j - D_{j-1} = dx^2 b_j, so D_j = D_0 + dx^2 * sum_{i=1..j} b_i with the
    constant fixed by sum_j D_j = 0, and E_j = -(D_j + D_{j-1}) / (2 dx).
    Returns E only (phi's additive constant is numerical noise in the
    reference and is never compared)."""
    dx = L / N_mesh
    b = n - n0
    S = np.cumsum(b)
    S = S - b[0]                      # S_j = sum_{i=1..j} b_i, S_0 = 0
    D = dx * dx * (S - S.mean())
    E = -(D + np.roll(D, 1)) / (2 * dx)
    return E


def gather(E_mesh, indx_l, indx_r, weight_l, weight_r):
    """src/env/util.py:106"""
    return weight_l * E_mesh[indx_l] + weight_r * E_mesh[indx_r]


# --------------------------------------------------------------------------
# Actuator                                      src/control/actuator.py:4-63
# --------------------------------------------------------------------------
def actuator_basis(L: float, N_mesh: int, max_mode: int):
    """basis_cos, basis_sin (N_mesh, max_mode) on np.linspace(0, L, N_mesh)
    -- endpoint included, src/control/actuator.py:13,20,23-24."""
    xm = np.linspace(0, L, N_mesh)
    k = np.array([2 * np.pi / L * n for n in range(1, max_mode + 1)])
    basis_cos = np.concatenate([np.cos(kk * xm).reshape(-1, 1) for kk in k], axis=1)
    basis_sin = np.concatenate([np.sin(kk * xm).reshape(-1, 1) for kk in k], axis=1)
    return basis_cos, basis_sin


def actuator_field(basis_cos, basis_sin, coeff_cos, coeff_sin):
    """src/control/actuator.py:62 -> flat (N_mesh,) vector."""
    E = basis_cos @ np.asarray(coeff_cos, dtype=np.float64).reshape(-1, 1) \
        + basis_sin @ np.asarray(coeff_sin, dtype=np.float64).reshape(-1, 1)
    return E[:, 0]


# --------------------------------------------------------------------------
# One env step                                        src/env/pic.py:131-146
# --------------------------------------------------------------------------
@dataclass
class PicParams:
    N: int
    N_mesh: int
    n0: float
    L: float
    dt: float                 # already CFL-clipped (clip_dt)
    gamma: float = 5.0
    interpol: str = "CIC"     # "CIC" | "TSC" (src/env/interpolate.py:4 / :22)

    @property
    def dx(self) -> float:
        return self.L / self.N_mesh


def _field(n, p: PicParams, faithful: bool, gamma: float):
    if faithful:
        return field_dense(n, p.n0, p.L, p.N_mesh, gamma)[1]
    return field_prefix(n, p.n0, p.L, p.N_mesh)


def accel(x, p: PicParams, E_ext: Optional[np.ndarray], faithful: bool):
    """-E at the particles for positions x: src/env/pic.py:125-127 via
    src/env/util.py:73-116 (gamma is hard-coded 5.0 there, util.py:99).
    Does not modify ``x`` (the reference wraps a scratch copy)."""
    xs = x.copy()
    if p.interpol == "TSC":
        n, il, im, ir, wl, wm, wr = compute_n(xs, p.dx, p.N_mesh, p.n0, p.L, p.N, "TSC")
        E_mesh = _field(n, p, faithful, 5.0)
        if E_ext is not None:
            E_mesh = E_mesh + E_ext
        return (-1) * (wl * E_mesh[il] + wm * E_mesh[im] + wr * E_mesh[ir]), im          # util.py:110
    n, il, ir, wl, wr = compute_n(xs, p.dx, p.N_mesh, p.n0, p.L, p.N)
    E_mesh = _field(n, p, faithful, 5.0)
    if E_ext is not None:
        E_mesh = E_mesh + E_ext
    return (-1) * gather(E_mesh, il, ir, wl, wr), il


def step(x: np.ndarray, v: np.ndarray, p: PicParams, E_ext: Optional[np.ndarray] = None,
         faithful: bool = False, return_stage_indices: bool = False):
    """One ``PIC.update_state(E_external)``.

    x, v: flat float64 arrays (N,). Returns a dict with the post-step env
    state: x, v, n, E_mesh (self-consistent only, pic.py:146), E (at
    particles), indx_l, indx_r, weight_l, weight_r.

    Yoshida stages follow src/env/integration.py:22-47: kick with d (skipped
    when d == 0), then drift with c.  The positions carried between stages are
    NOT wrapped (``q`` is copied before ``grad_func`` wraps its argument)."""
    cs, ds = yoshida_coefficients()
    dt = p.dt
    q = x.astype(np.float64).copy()
    pm = v.astype(np.float64).copy()
    if E_ext is not None:
        E_ext = np.asarray(E_ext, dtype=np.float64).reshape(-1)
    stage_idx = []
    for c, d in zip(cs, ds):
        if d != 0.0:
            a, il = accel(q, p, E_ext, faithful)          # integration.py:32
            stage_idx.append(il)
            pm = pm + d * a * dt
        if faithful:
            # integration.py:42 evaluates grad_func(eta_m) only to read back p;
            # the deposit + solve it triggers is discarded work.
            accel(q, p, E_ext, True)
        q = q + c * pm * dt                                # integration.py:42
    xf = np.mod(q, p.L)                                    # pic.py:139
    if p.interpol == "TSC":
        n, il, im, ir, wl, wm, wr = compute_n(xf, p.dx, p.N_mesh, p.n0, p.L, p.N, "TSC")
        E_mesh = _field(n, p, faithful, p.gamma)
        E = wl * E_mesh[il] + wm * E_mesh[im] + wr * E_mesh[ir]          # pic.py:123
        out = dict(x=xf, v=pm, n=n, E_mesh=E_mesh, E=E, indx_l=il, indx_m=im, indx_r=ir, weight_l=wl, weight_m=wm,
                   weight_r=wr)
        if return_stage_indices:
            out["stage_indx_l"] = stage_idx
        return out
    n, il, ir, wl, wr = compute_n(xf, p.dx, p.N_mesh, p.n0, p.L, p.N)   # pic.py:145 (wraps xf again in place)
    E_mesh = _field(n, p, faithful, p.gamma)               # pic.py:116-117
    E = gather(E_mesh, il, ir, wl, wr)                     # pic.py:120
    out = dict(x=xf, v=pm, n=n, E_mesh=E_mesh, E=E, indx_l=il, indx_r=ir,
               weight_l=wl, weight_r=wr)
    if return_stage_indices:
        out["stage_indx_l"] = stage_idx
    return out


def init_fields(x: np.ndarray, p: PicParams, faithful: bool = False):
    """update_density + update_E_field at construction, src/env/pic.py:76-77."""
    xf = x.astype(np.float64).copy()
    n, il, ir, wl, wr = compute_n(xf, p.dx, p.N_mesh, p.n0, p.L, p.N)
    E_mesh = _field(n, p, faithful, p.gamma)
    return dict(x=xf, n=n, E_mesh=E_mesh, E=gather(E_mesh, il, ir, wl, wr), indx_l=il, indx_r=ir,
                weight_l=wl, weight_r=wr)


def perturb_velocity(x, v, A: float, n_mode: int, L: float):
    """src/env/pic.py:68"""
    return v * (1 + A * np.sin(2 * np.pi * n_mode * x / L))


# --------------------------------------------------------------------------
# Energies / reward terms
# --------------------------------------------------------------------------
def mesh_field_of_state(x, p: PicParams, faithful: bool = False):
    """E_mesh of the self-consistent field for positions x (what
    compute_electric_energy / estimate_electric_energy recompute:
    src/env/util.py:128, src/control/objective.py:24)."""
    xs = np.asarray(x, dtype=np.float64).reshape(-1).copy()
    n, *_ = compute_n(xs, p.dx, p.N_mesh, p.n0, p.L, p.N, p.interpol)
    return _field(n, p, faithful, 5.0)


def pe_mesh(E_mesh, dx: float) -> float:
    """0.5 * sum(E^2) * dx                    src/control/objective.py:31"""
    return float(0.5 * np.sum(E_mesh * E_mesh) * dx)


def electric_energy(x, p: PicParams, faithful: bool = False) -> float:
    """PIC.get_electric_energy: src/env/util.py:119-131 (PE_mesh * N / L)."""
    PE = pe_mesh(mesh_field_of_state(x, p, faithful), p.dx)
    PE *= p.N / p.L
    return PE


def kinetic_energy(v) -> float:
    """src/env/util.py:144"""
    return float(0.5 * np.sum(v * v))


def hamiltonian(x, v, p: PicParams, faithful: bool = False) -> float:
    """PIC.get_energy: src/env/util.py:133-147"""
    return kinetic_energy(v) + electric_energy(x, p, faithful)


def input_energy(actions, L: float) -> float:
    """src/control/rl/reward.py:52-54"""
    return float(np.sum(np.asarray(actions, dtype=np.float64) ** 2) * L * 0.25)


def reward(pe_mesh_value: float, actions, L: float, alpha: float = 1.0, beta: float = 1.0,
           n_actions: int = 10) -> float:
    """Reward.compute_reward, src/control/rl/reward.py:71-76, with r_pe_n = 1
    (:32) and r_ie_n = input energy of n_actions ones (:33)."""
    r_ie_n = input_energy(np.ones(n_actions), L)
    r_pe = max(1.0 - pe_mesh_value / 1.0, 0)
    r_ie = max(1.0 - input_energy(actions, L) / r_ie_n, 0)
    return r_pe * alpha + r_ie * beta


def spectrum_modes(E_mesh, max_mode: int):
    """First max_mode Fourier modes of E_mesh the way the feedback / behaviour
    cloning target uses them: fft / N_mesh * 2, modes 1..max_mode
    (src/interpret/spectrum.py:17, src/control/rl/ddpg.py:429-431)."""
    Ek = np.fft.fft(E_mesh) / E_mesh.shape[0] * 2.0
    return Ek[1:max_mode + 1]


def growth_rate(pe_mesh_trace, tmax: float):
    """0.5 * slope of log(sum E^2 dx) against time, ordinary least squares on
    np.linspace(0, tmax, Nt) (src/interpret/landau.py:44-60 uses sklearn's
    LinearRegression for the same fit).  pe_mesh_trace = 0.5*sum(E^2)*dx."""
    y = np.log(2.0 * np.asarray(pe_mesh_trace, dtype=np.float64))
    t = np.linspace(0, tmax, y.shape[0])
    A = np.stack([t, np.ones_like(t)], axis=1)
    slope = np.linalg.lstsq(A, y, rcond=None)[0][0]
    return 0.5 * float(slope)


# --------------------------------------------------------------------------
# Initial samplers (host RNG, global legacy numpy stream)   src/env/dist.py
# --------------------------------------------------------------------------
def _gauss(v, vb, sigma):
    """src/env/dist.py:66-68 / :147-149"""
    return 1 / np.sqrt(2 * np.pi) / sigma * np.exp(-0.5 * (v - vb) ** 2 / sigma ** 2)


def _accept_until(n_target, L, vb, sigma, pos, vel, strict_le=False, batch=1000):
    """The accept loop of src/env/dist.py:75-81 / :161-168: three uniform
    draws of ``batch`` per round from the GLOBAL numpy RNG, keep u < target."""
    def more():
        return (len(pos) <= n_target) if strict_le else (len(pos) < n_target)

    while more():
        x = np.random.uniform(0, L, size=batch)
        v = np.random.uniform(-10, 10, size=batch)
        u = np.random.uniform(0, 1.0, size=batch)
        keep = u < _gauss(v, vb, sigma)
        pos += x[keep].tolist()
        vel += v[keep].tolist()


def sample_two_stream(v0: float, sigma: float, n_samples: int, L: float):
    """src/env/dist.py:70-102 (note the ``<=`` in the first loop, :75)."""
    pos, vel = [], []
    _accept_until(n_samples // 2, L, v0, sigma, pos, vel, strict_le=True)
    pos = pos[:n_samples // 2]
    vel = vel[:n_samples // 2]
    _accept_until(n_samples, L, -v0, sigma, pos, vel)
    return np.array(pos[:n_samples]), np.array(vel[:n_samples])


def sample_bump_on_tail(a: float, v0: float, sigma: float, n_samples: int, L: float):
    """src/env/dist.py:151-189"""
    pos, vel = [], []
    N1 = int(n_samples * (1 / (1 + a)))
    _accept_until(N1, L, 0.0, 1.0, pos, vel)
    pos = pos[:N1]
    vel = vel[:N1]
    _accept_until(n_samples, L, v0, sigma, pos, vel)
    return np.array(pos[:n_samples]), np.array(vel[:n_samples])


def runner_initial_state(simcase: str, N=5000, L=50.0, vb=3.0, vth=1.0, a=0.2, A=0.1, n_mode=2,
                         seed=42):
    """(x, v) exactly as ``run_wo_oc.py`` / ``run_ddpg.py`` reach their first
    step: np.random.seed(42) at import of src/env/pic.py:12, one sample drawn
    by the dist constructor and discarded, a second drawn by PIC.initialize
    (pic.py:64), then the velocity perturbation (pic.py:68)."""
    np.random.seed(seed)
    for _ in range(2):
        if simcase == "two-stream":
            x, v = sample_two_stream(vb, vth, N, L)
        elif simcase == "bump-on-tail":
            x, v = sample_bump_on_tail(a, vb, vth, N, L)
        else:
            raise ValueError(simcase)
    v = perturb_velocity(x, v, A, n_mode, L)
    return x, v


# --------------------------------------------------------------------------
# float32 mode (no reference counterpart: the reference is float64 only)
# --------------------------------------------------------------------------
def step_f32(x, v, p: PicParams, E_ext: Optional[np.ndarray] = None):
    """Restatement of the library's PIC_F32 mode for index parity: particle state, wrap, cell index, weights,
    gather, kick and drift in float32 (same operation order as `step`), density accumulation and the mesh field in
    float64, the gather table rounded to float32.  Returns dict(x, v, indx_l, E_mesh, n) with x, v float32."""
    f = np.float32
    cs, ds = yoshida_coefficients()
    L, dx, dt = f(p.L), f(p.dx), f(p.dt)
    inv_dx = f(1) / dx
    q = np.asarray(x, dtype=f).copy()
    pm = np.asarray(v, dtype=f).copy()
    ext = None if E_ext is None else np.asarray(E_ext, dtype=np.float64).reshape(-1)

    def cells(xq):
        xw = np.mod(np.mod(xq, L), L).astype(f)
        il = np.floor(xw / dx).astype(np.int64)
        fl = il.astype(f)
        wl = (((fl + f(1)) * dx - xw) * inv_dx).astype(f)
        wr = ((xw - fl * dx) * inv_dx).astype(f)
        return xw, il, wl, wr

    def density(il, wr):
        w = wr.astype(np.float64)
        n = np.bincount(il, weights=1.0 - w, minlength=p.N_mesh) + np.bincount((il + 1) % p.N_mesh, weights=w,
                                                                                minlength=p.N_mesh)
        return n * (p.n0 * p.L / p.N / p.dx)

    for c, d in zip(cs, ds):
        if d != 0.0:
            xw, il, wl, wr = cells(q)
            E_mesh = field_prefix(density(il, wr), p.n0, p.L, p.N_mesh)
            Et = (E_mesh + ext if ext is not None else E_mesh).astype(f)
            Ep = (wl * Et[il] + wr * Et[(il + 1) % p.N_mesh]).astype(f)
            pm = (pm + (f(d) * (-Ep)) * dt).astype(f)
        q = (q + (f(c) * pm) * dt).astype(f)
    xw, il, wl, wr = cells(q)
    n = density(il, wr)
    return dict(x=xw, v=pm, indx_l=il, n=n, E_mesh=field_prefix(n, p.n0, p.L, p.N_mesh))
