"""`PIC` -- the reference's env class (src/env/pic.py:11-173) with the step on a B200.

Same constructor keywords, methods and attributes as the reference class, so `run_wo_oc.py` / `run_ddpg.py` and the
DDPG / PPO / SAC trainers use it unchanged (see run.py for the launcher that injects it as `src.env.pic`).  What
differs is where the work happens: `update_state` enqueues the Yoshida-4 step on the device through the C ABI
(include/pic_b200.h); `.x`, `.v`, `.n`, `.E_mesh`, ... are fetched lazily and cached until the next step.
There is no CPU implementation of the step in this package.
"""
from typing import Optional

import numpy as np

from . import _lib as L
from .engine import Engine


class PIC:
    def __init__(self, N: int = 40000, N_mesh: int = 400, n0: float = 1.0, L: float = 50.0, dt: float = 1.0,
                 tmin: float = 0.0, tmax: float = 50.0, gamma: float = 5.0, A: float = 0.1, n_mode: int = 4,
                 interpol: str = "CIC", init_dist=None, *, device: int = 0, precision: str = "f64",
                 mode: str = "auto", deposit: str = "auto", exact_weights: bool = False, max_mode: int = 0):
        if interpol not in ("CIC", "TSC"):
            raise ValueError("interpol must be 'CIC' or 'TSC' (src/env/interpolate.py), got %r" % (interpol,))
        self._eng: Optional[Engine] = None
        self._cache = {}
        self.N = N
        self.N_mesh = N_mesh
        self.n0 = n0
        self.L = L
        self._dt = dt
        self.tmin = tmin
        self.tmax = tmax
        self.dx = L / N_mesh
        self.gamma = gamma           # the reference's Sherman-Morrison parameter; E does not depend on it
        self.A = A
        self.n_mode = n_mode
        self.init_dist = init_dist
        self._interpol = interpol
        self._opts = dict(device=device, precision=precision, mode=mode, deposit=deposit,
                          exact_weights=exact_weights, max_mode=max_mode)
        self._basis = None
        if init_dist is not None:
            self.initialize()

    # ------------------------------------------------------------------ engine
    def _engine(self) -> Engine:
        if self._eng is None:
            o = self._opts
            self._eng = Engine(self.N, self.N_mesh, self.L, self.dt, n0=self.n0, n_envs=1, precision=o["precision"],
                               mode=o["mode"], deposit=o["deposit"], exact_weights=o["exact_weights"],
                               device=o["device"], max_mode=o["max_mode"], interpol=self.interpol)
        return self._eng

    @property
    def engine(self) -> Engine:
        return self._engine()

    def _rebuild_engine(self):
        """A parameter the device handle was created with has changed (dt, interpol): move the state into a new
        handle.  `pic_set_state` re-wraps (a no-op on a wrapped state), re-deposits and also re-deposits the next
        step's drift-only sub-stage, which had the old c0*dt baked in."""
        if self._eng is None:
            return
        x, v = self._eng.get_state()
        self._eng.close()
        self._eng = None
        self._cache = {}
        eng = self._engine()
        if self._basis is not None:
            eng.set_actuator_basis(*self._basis)
        eng.set_state(x, v)

    # dt and interpol are plain attributes in the reference (pic.py:35,47) that `update_state` reads on every call
    # (pic.py:133, util.py:94), so assigning them -- directly or through update_params (pic.py:79-82) -- must take
    # effect on the next step here as well.
    @property
    def dt(self):
        return self._dt

    @dt.setter
    def dt(self, value):
        value = float(value)
        if not value > 0:
            raise ValueError("dt must be positive")
        changed = value != self._dt
        self._dt = value
        if changed:
            self._rebuild_engine()

    @property
    def interpol(self):
        return self._interpol

    @interpol.setter
    def interpol(self, value):
        if value not in ("CIC", "TSC"):
            raise ValueError("interpol must be 'CIC' or 'TSC' (src/env/interpolate.py), got %r" % (value,))
        changed = value != self._interpol
        self._interpol = value
        if changed:
            self._rebuild_engine()

    # ------------------------------------------------------- src/env/pic.py:63-91
    def initialize(self):
        self.init_dist.reinit()
        x, v = self.init_dist.get_sample()
        x = np.asarray(x, dtype=np.float64).reshape(-1, 1)
        v = np.asarray(v, dtype=np.float64).reshape(-1, 1)
        v *= (1 + self.A * np.sin(2 * np.pi * self.n_mode * x / self.L))     # pic.py:68
        if self._dt > 2 / np.sqrt(self.N / self.L):                          # pic.py:71-73
            self._dt = 2 / np.sqrt(self.N / self.L)
            print("CFL condtion invalid: change dt = {:.4f}".format(self._dt))
            if self._eng is not None:
                self._eng.close()
                self._eng = None
        self.set_state(x, v)

    def reinit(self):
        self.initialize()

    def set_state(self, x, v):
        """Env checkpoint restore: load particles and rebuild density / field on the device (pic.py:76-77)."""
        self._engine().set_state(np.asarray(x, dtype=np.float64).reshape(1, -1),
                                 np.asarray(v, dtype=np.float64).reshape(1, -1))
        self._cache = {}

    # ----------------------------------------------------- src/env/pic.py:131-146
    def update_state(self, E_external: Optional[np.ndarray] = None):
        eng = self._engine()
        if E_external is None:
            eng.step_mesh(None, 1)
        else:
            eng.step_mesh(np.asarray(E_external, dtype=np.float64).reshape(1, self.N_mesh), 1)
        self._cache = {}

    def update_state_coeffs(self, coeff_cos, coeff_sin, basis_cos=None, basis_sin=None):
        """update_state(E_field.compute_E()) without building the mesh vector on the host: the device evaluates
        basis_cos @ a + basis_sin @ b (src/control/actuator.py:62).  Needs max_mode at construction."""
        eng = self._engine()
        if basis_cos is not None:
            self._basis = (np.array(basis_cos, dtype=np.float64), np.array(basis_sin, dtype=np.float64))
            eng.set_actuator_basis(*self._basis)
        if self._basis is None:
            raise RuntimeError("pass basis_cos / basis_sin (E_field.basis_cos / .basis_sin) on the first call")
        c = np.concatenate([np.asarray(coeff_cos, dtype=np.float64).ravel(),
                            np.asarray(coeff_sin, dtype=np.float64).ravel()])
        eng.step_coeffs(c.reshape(1, 1, -1), 1)
        self._cache = {}

    # ------------------------------------------------------------------ getters
    def _state(self):
        if "xv" not in self._cache:
            self._diag()                                   # raises if the device flagged the step that made this state
            x, v = self._engine().get_state()
            x, v = x.reshape(-1, 1), v.reshape(-1, 1)
            # The arrays are host COPIES of the device state, cached until the next step.  In the reference `sim.x`
            # IS the state, so `sim.x[:] = ...` changes the env; here such a write would be silently lost -- the
            # arrays are therefore read-only and the write fails loudly.  Assign instead (`sim.x = new_x`, or
            # `set_state(x, v)`): the setters upload and rebuild density / field on the device.
            x.setflags(write=False)
            v.setflags(write=False)
            self._cache["xv"] = (x, v)
        return self._cache["xv"]

    @property
    def x(self):
        """(N, 1) float64 positions, read-only host copy (see _state); `.copy()` it or assign a new array."""
        return self._state()[0]

    @x.setter
    def x(self, value):
        self.set_state(value, self._state()[1])

    @property
    def v(self):
        return self._state()[1]

    @v.setter
    def v(self, value):
        self.set_state(self._state()[0], value)

    def get_state(self):                                               # pic.py:165-167
        x, v = self._state()
        return np.concatenate([x.copy().reshape(-1, 1), v.copy().reshape(-1, 1)], axis=0)

    def _diag(self):
        if "diag" not in self._cache:
            # one D2H round trip brings the diagnostics record and the sticky device error flags: where the reference
            # raises (np.bincount on an out-of-range or NaN index, interpolate.py:16) the device path clamps, flags
            # and keeps going -- the flag becomes an exception here, before any number of that state is handed out
            self._cache["diag"] = self._engine().get_diag(check=True)[0]
        return self._cache["diag"]

    def get_electric_energy(self):                                     # util.py:119-131
        PE = float(self._diag()[L.DIAG_PE_MESH])
        PE *= self.N / self.L
        return PE

    def get_energy(self):                                              # util.py:133-147
        return float(self._diag()[L.DIAG_KE]) + self.get_electric_energy()

    def get_kinetic_energy(self):
        return float(self._diag()[L.DIAG_KE])

    def get_mesh_energy(self):
        """0.5 * sum(E_mesh^2) * dx of the self-consistent field: the reward's energy term
        (src/control/objective.py:31) for the current state."""
        return float(self._diag()[L.DIAG_PE_MESH])

    def _fields(self):
        if "fields" not in self._cache:
            n, E = self._engine().get_fields()
            self._cache["fields"] = (n[0], E[0].reshape(-1, 1))
        return self._cache["fields"]

    @property
    def n(self):
        return self._fields()[0]

    @property
    def E_mesh(self):
        return self._fields()[1]

    @property
    def phi_mesh(self):
        """Potential on the mesh, (N_mesh, 1), of the same 3-point discretisation (laplacian @ phi = n - n0) in the
        zero-mean gauge.  A derived read-out for API completeness (no runner uses it), rebuilt on the host from the
        device's density; the reference's own additive constant is numerical noise of its singular solve (its
        Sherman-Morrison denominator is ~1e-15), so only differences of phi -- i.e. E -- are comparable."""
        b = self.n - self.n0
        S = np.cumsum(b)
        D = self.dx * self.dx * (S - S.mean())           # D_j = phi_{j+1} - phi_j
        phi = np.concatenate([[0.0], np.cumsum(D[:-1])])
        return (phi - phi.mean()).reshape(-1, 1)

    @property
    def grad(self):
        """Dense periodic centred-difference matrix of src/env/util.py:7-26 (attribute parity only)."""
        M, g = self.N_mesh, np.zeros((self.N_mesh, self.N_mesh))
        i = np.arange(M)
        g[i, (i + 1) % M] = 1.0
        g[i, (i - 1) % M] = -1.0
        return g / (2 * self.dx)

    @property
    def laplacian(self):
        """Dense periodic 3-point Laplacian of src/env/util.py:28-46 (attribute parity only)."""
        M, a = self.N_mesh, np.zeros((self.N_mesh, self.N_mesh))
        i = np.arange(M)
        a[i, (i + 1) % M] = 1.0
        a[i, (i - 1) % M] = 1.0
        a[i, i] = -2.0
        return a / self.dx ** 2

    def _cells(self):
        if "cells" not in self._cache:
            il, wl, wr, E, wm = self._engine().get_cells(want_wm=True)
            self._cache["cells"] = (il[0].astype(np.int64).reshape(-1, 1), wl[0].reshape(-1, 1), wr[0].reshape(-1, 1),
                                    E[0].reshape(-1, 1), wm[0].reshape(-1, 1))
        return self._cache["cells"]

    @property
    def indx_l(self):
        if self.interpol == "TSC":                       # interpolate.py:34
            return np.mod(self._cells()[0] - 1, self.N_mesh)
        return self._cells()[0]

    @property
    def indx_m(self):
        return self._cells()[0] if self.interpol == "TSC" else None

    @property
    def weight_m(self):
        return self._cells()[4] if self.interpol == "TSC" else None

    @property
    def indx_r(self):
        return np.mod(self._cells()[0] + 1, self.N_mesh)

    @property
    def weight_l(self):
        return self._cells()[1]

    @property
    def weight_r(self):
        return self._cells()[2]

    @property
    def E(self):
        return self._cells()[3]

    _STRUCTURAL = ("N", "N_mesh", "L", "n0")

    def update_params(self, **kwargs):                                 # pic.py:79-82
        """Same contract as the reference: every known, non-None keyword is assigned.  `dt` and `interpol` take effect
        on the next step (the device handle is rebuilt around the current state).  N, N_mesh, L, n0 size the device
        buffers and the mesh; the reference would keep its old dx / grad / laplacian after such a change
        (pic.py:36,52-53 are only evaluated in __init__) and compute garbage, so they are refused here."""
        for key in kwargs.keys():
            if hasattr(self, key) is True and kwargs[key] is not None:
                if key in self._STRUCTURAL and kwargs[key] != getattr(self, key) and self._eng is not None:
                    raise ValueError("update_params(%s=...) after initialisation is not supported: build a new PIC "
                                     "(the reference keeps a stale dx/grad/laplacian in this case)" % key)
                setattr(self, key, kwargs[key])

    def simulate(self, E_external_traj=None, n_steps: Optional[int] = None):
        """Trajectory driver in the spirit of pic.py:175 (unused by the runners): returns (snapshot (2N, Nt+1),
        H (Nt+1,), PE (Nt+1,))."""
        Nt = int(np.ceil((self.tmax - self.tmin) / self.dt)) if n_steps is None else int(n_steps)
        snap = [self.get_state()]
        H = [self.get_energy()]
        PE = [self.get_electric_energy()]
        for i in range(Nt):
            self.update_state(None if E_external_traj is None else E_external_traj[i])
            snap.append(self.get_state()); H.append(self.get_energy()); PE.append(self.get_electric_energy())
        return np.concatenate(snap, axis=1), np.array(H), np.array(PE)
