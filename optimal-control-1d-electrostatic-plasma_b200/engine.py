"""Thin object wrapper over one pic_handle (include/pic_b200.h).  All array arguments are numpy float64 unless a
``*_device`` method is used; nothing here computes physics on the host."""
import ctypes as C

import numpy as np

from . import _lib as L

_PREC = {"f64": L.PIC_F64, "fp64": L.PIC_F64, "float64": L.PIC_F64, "f32": L.PIC_F32, "fp32": L.PIC_F32,
         "float32": L.PIC_F32}
_MODE = {"auto": L.PIC_MODE_AUTO, "resident": L.PIC_MODE_RESIDENT, "streaming": L.PIC_MODE_STREAMING}
_DEP = {"auto": L.PIC_DEPOSIT_AUTO, "cas64": L.PIC_DEPOSIT_CAS64, "split32": L.PIC_DEPOSIT_SPLIT32}


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a, shape):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.size != int(np.prod(shape)):
        raise ValueError("expected %s elements, got array of shape %s" % (shape, a.shape))
    return a.reshape(shape)


class DeviceArray:
    """Zero-copy view of a device buffer owned by an Engine (``__cuda_array_interface__`` v3), so that
    ``torch.as_tensor(view, device='cuda')`` aliases the particle / mesh arrays without a copy."""

    def __init__(self, ptr, shape, typestr, strides=None, owner=None):
        self._owner = owner
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": strides}


class Engine:
    def __init__(self, n_particles, n_mesh, L_box, dt, n0=1.0, n_envs=1, n_particles_total=0, precision="f64",
                 mode="auto", deposit="auto", fixed_bits=0, exact_weights=False, device=0, max_mode=0, stream=None,
                 interpol="CIC"):
        self._lib = L.load()
        self._h = C.c_void_p()
        cfg = L.PicConfig(int(n_particles), int(n_particles_total), int(n_mesh), int(n_envs), float(n0), float(L_box),
                          float(dt), _PREC[precision], _MODE[mode], _DEP[deposit], int(fixed_bits),
                          int(bool(exact_weights)), int(device), int(max_mode),
                          C.c_void_p(-1 if stream == "own" else (stream or 0)),     # "own": PIC_STREAM_OWN
                          {"CIC": L.PIC_INTERP_CIC, "TSC": L.PIC_INTERP_TSC}[interpol])
        self.interpol = interpol
        rc = self._lib.pic_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            msg = self._lib.pic_last_error(None)
            self._h = None
            raise L.PicError(rc, msg.decode() if msg else "")
        self.N, self.M, self.n_envs, self.m = int(n_particles), int(n_mesh), int(n_envs), int(max_mode)
        self.precision = "f32" if _PREC[precision] == L.PIC_F32 else "f64"
        self.device = int(device)
        self.L = float(L_box)
        self.n_modes = 0

    # ---- lifetime
    def close(self):
        if getattr(self, "_h", None):
            self._lib.pic_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            msg = self._lib.pic_last_error(self._h)
            raise L.PicError(rc, msg.decode() if msg else "")

    # ---- state
    def set_state(self, x, v):
        x = _f64(x, (self.n_envs, self.N)); v = _f64(v, (self.n_envs, self.N))
        self._ck(self._lib.pic_set_state(self._h, _ptr(x), _ptr(v)))

    def set_state_device(self, x_ptr, v_ptr):
        self._ck(self._lib.pic_set_state_device(self._h, C.c_void_p(x_ptr), C.c_void_p(v_ptr)))

    def sample_state(self, kind="bump-on-tail", a=0.2, v0=3.0, sigma=1.0, A=0.1, n_mode=2, seed=42, global_offset=0,
                     n_global=0, env_offset=0):
        """Device-side sampler + perturbation + field build (no host particles)."""
        k = {"bump-on-tail": 0, "two-stream": 1}[kind]
        self._ck(self._lib.pic_sample_state(self._h, k, float(a), float(v0), float(sigma), float(A), int(n_mode),
                                            int(seed), int(global_offset), int(n_global), int(env_offset)))

    def get_state(self, want_x=True, want_v=True):
        x = np.empty((self.n_envs, self.N)) if want_x else None
        v = np.empty((self.n_envs, self.N)) if want_v else None
        self._ck(self._lib.pic_get_state(self._h, _ptr(x), _ptr(v)))
        return x, v

    def set_state_ptr(self, x, v):
        """x, v: addresses of host float64 buffers (e.g. pinned) of n_envs * N elements."""
        self._ck(self._lib.pic_set_state(self._h, C.c_void_p(x), C.c_void_p(v)))

    def get_state_into(self, x, v):
        """Copies into caller-provided (e.g. pinned) float64 buffers."""
        self._ck(self._lib.pic_get_state(self._h, C.c_void_p(x), C.c_void_p(v)))

    def get_fields(self):
        n = np.empty((self.n_envs, self.M)); E = np.empty((self.n_envs, self.M))
        self._ck(self._lib.pic_get_fields(self._h, _ptr(n), _ptr(E)))
        return n, E

    def get_density_fixed(self):
        rho = np.empty((self.n_envs, self.M), dtype=np.uint64)
        k = C.c_int32()
        self._ck(self._lib.pic_get_density_fixed(self._h, _ptr(rho), C.byref(k)))
        return rho, k.value

    def get_diag(self, check=False):
        """(n_envs, DIAG_N) diagnostics of the current state.  check=True: the sticky device error flags come back
        in the same round trip and raise PicDeviceError when set."""
        d = np.empty((self.n_envs, L.DIAG_N))
        if not check:
            self._ck(self._lib.pic_get_diag(self._h, _ptr(d)))
            return d
        f = C.c_uint32()
        self._ck(self._lib.pic_get_diag_flags(self._h, _ptr(d), C.byref(f)))
        if f.value:
            raise L.PicDeviceError(f.value)
        return d

    def check_errors(self):
        """Raise PicDeviceError if the device has flagged anything since the flags were last cleared."""
        f = self.error_flags()
        if f:
            raise L.PicDeviceError(f)

    def get_trace(self, n_steps):
        t = np.empty((n_steps, self.n_envs, L.DIAG_N))
        self._ck(self._lib.pic_get_trace(self._h, _ptr(t), n_steps))
        return t

    def get_cells(self, want_weights=True, want_E=True, want_wm=False):
        """(indx, weight_l, weight_r, E[, weight_m]); indx = indx_l for CIC, indx_m for TSC."""
        il = np.empty((self.n_envs, self.N), dtype=np.int32)
        wl = np.empty((self.n_envs, self.N)) if want_weights else None
        wr = np.empty((self.n_envs, self.N)) if want_weights else None
        E = np.empty((self.n_envs, self.N)) if want_E else None
        wm = np.empty((self.n_envs, self.N)) if want_wm else None
        self._ck(self._lib.pic_get_cells(self._h, _ptr(il), _ptr(wl), _ptr(wr), _ptr(E), _ptr(wm)))
        return (il, wl, wr, E, wm) if want_wm else (il, wl, wr, E)

    # ---- hot path
    def step_mesh(self, E_ext=None, n_steps=1):
        e = None if E_ext is None else _f64(E_ext, (self.n_envs, self.M))
        self._ck(self._lib.pic_step_mesh(self._h, _ptr(e), int(n_steps)))

    def step_mesh_ptr(self, host_ptr, n_steps=1):
        self._ck(self._lib.pic_step_mesh(self._h, C.c_void_p(host_ptr), int(n_steps)))

    def set_actuator_basis(self, basis_cos, basis_sin):
        bc = _f64(basis_cos, (self.M, self.m)); bs = _f64(basis_sin, (self.M, self.m))
        self._ck(self._lib.pic_set_actuator_basis(self._h, _ptr(bc), _ptr(bs), self.m))

    def step_coeffs(self, coeffs, n_steps=1):
        c = _f64(coeffs, (n_steps, self.n_envs, 2 * self.m))
        self._ck(self._lib.pic_step_coeffs(self._h, _ptr(c), int(n_steps)))

    def step_coeffs_ptr(self, host_ptr, n_steps=1):
        self._ck(self._lib.pic_step_coeffs(self._h, C.c_void_p(host_ptr), int(n_steps)))

    def step_mesh_device(self, dev_ptr, n_steps=1):
        self._ck(self._lib.pic_step_mesh_device(self._h, C.c_void_p(dev_ptr or 0), int(n_steps)))

    def step_coeffs_device(self, dev_ptr, n_steps=1):
        self._ck(self._lib.pic_step_coeffs_device(self._h, C.c_void_p(dev_ptr), int(n_steps)))

    def set_reward(self, alpha=1.0, beta=1.0, r_pe_n=1.0, r_ie_n=None, n_actions=10):
        """Reward constants of src/control/rl/reward.py (r_ie_n defaults to n_actions * L / 4, reward.py:33)."""
        if r_ie_n is None:
            r_ie_n = float(n_actions) * self.L * 0.25
        self._ck(self._lib.pic_set_reward(self._h, float(alpha), float(beta), float(r_pe_n), float(r_ie_n)))

    def enable_modes(self, n_modes):
        self._ck(self._lib.pic_enable_modes(self._h, int(n_modes)))
        self.n_modes = int(n_modes)

    def get_modes(self):
        """(n_envs, 2m): Re_1..Re_m, Im_1..Im_m of fft(E_mesh)/N_mesh*2 (spectrum.py:17) for the current state."""
        out = np.empty((self.n_envs, 2 * self.n_modes))
        self._ck(self._lib.pic_get_modes(self._h, _ptr(out)))
        return out

    def get_mode_trace(self, n_steps):
        out = np.empty((n_steps, self.n_envs, 2 * self.n_modes))
        self._ck(self._lib.pic_get_mode_trace(self._h, _ptr(out), int(n_steps)))
        return out

    def phase_hist_config(self, vmin=-25.0, vmax=25.0, nbins=None):
        """Configure the phase-space histogram (objective.py:8-14); nbins defaults to N_mesh as in Reward."""
        self._nb = int(self.M if nbins is None else nbins)
        self._ph = (float(vmin), float(vmax))
        self._ck(self._lib.pic_phase_hist_config(self._h, float(vmin), float(vmax), self._nb))

    def phase_hist(self):
        """Raw counts (n_envs, nbins, nbins) uint32 == np.histogram2d(x, v, bins, range=[[0, L], [vmin, vmax]])[0]."""
        c = np.empty((self.n_envs, self._nb, self._nb), dtype=np.uint32)
        self._ck(self._lib.pic_phase_hist(self._h, _ptr(c)))
        return c

    def estimate_f(self, n0=1.0, n_total=None):
        """estimate_f of objective.py:8-14 for the current state."""
        dx, dv = self.L / self._nb, (self._ph[1] - self._ph[0]) / self._nb
        return self.phase_hist().astype(np.float64) * (n0 / dx / dv / (self.N if n_total is None else n_total))

    def set_feq(self, feq):
        f = _f64(feq, (self._nb, self._nb))
        self._ck(self._lib.pic_set_feq(self._h, _ptr(f)))

    def kl_divergence(self):
        """Reward.compute_kl_divergence (reward.py:43-46) of the current state of every env against f_eq."""
        kl = np.empty(self.n_envs)
        self._ck(self._lib.pic_kl_divergence(self._h, _ptr(kl)))
        return kl

    def refresh_fields(self):
        """Rebuild density / field / diagnostics / next-step pre-deposit from the particle arrays as they are now on
        the device: mandatory after writing x or v through `views()` (which are read-only by contract otherwise)."""
        self._ck(self._lib.pic_refresh_fields(self._h))

    def sync(self):
        self._ck(self._lib.pic_sync(self._h))

    def stream_ptr(self):
        """cudaStream_t (as int) the handle enqueues on; wrap with torch.cuda.ExternalStream to record events on it."""
        p = C.c_void_p()
        self._ck(self._lib.pic_get_stream(self._h, C.byref(p)))
        return int(p.value or 0)

    def set_stream(self, stream_ptr):
        self._ck(self._lib.pic_set_stream(self._h, C.c_void_p(stream_ptr or 0)))

    def error_flags(self):
        f = C.c_uint32()
        self._ck(self._lib.pic_get_error_flags(self._h, C.byref(f)))
        return f.value

    def clear_error_flags(self):
        self._ck(self._lib.pic_clear_error_flags(self._h))

    # ---- staged driving / sharding
    def run_stage(self, stage):
        self._ck(self._lib.pic_run_stage(self._h, int(stage)))

    def stage_density_ptr(self, stage):
        p = C.c_void_p()
        self._ck(self._lib.pic_stage_density(self._h, int(stage), C.byref(p)))
        return p.value

    def set_stage_actuation(self, ext_dev=None, coeffs_dev=None):
        self._ck(self._lib.pic_set_stage_actuation(self._h, C.c_void_p(ext_dev or 0), C.c_void_p(coeffs_dev or 0)))

    def comm_init_rank(self, uid: bytes, rank, world):
        self._ck(self._lib.pic_comm_init_rank(self._h, uid, int(rank), int(world)))

    def comm_exchange_words(self, world):
        return int(self._lib.pic_comm_exchange_words(self._h, int(world)))

    def comm_init_peer(self, rank, world, exch_ptrs, flag_ptrs, exch_words):
        """Fused exchange over peer memory: exch_ptrs / flag_ptrs are the device addresses of every rank's exchange
        buffer / flag array as mapped into THIS process."""
        ea = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in exch_ptrs])
        fa = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in flag_ptrs])
        self._ck(self._lib.pic_comm_init_peer(self._h, int(rank), int(world), ea, fa, int(exch_words)))

    def comm_set_multicast(self, exch_multicast_ptr):
        """NVLS multicast mapping of the exchange buffers (0 / None: per-rank peer stores)."""
        self._ck(self._lib.pic_comm_set_multicast(self._h, C.c_void_p(int(exch_multicast_ptr or 0))))

    @staticmethod
    def nccl_unique_id():
        buf = C.create_string_buffer(128)
        rc = L.load().pic_nccl_unique_id(buf)
        if rc != 0:
            raise L.PicError(rc, (L.load().pic_last_error(None) or b"").decode())
        return buf.raw

    # ---- introspection
    def launch_info(self):
        v = [C.c_int32() for _ in range(7)]
        self._ck(self._lib.pic_get_launch_info(self._h, *[C.byref(a) for a in v]))
        keys = ("mode", "threads", "per_thread", "grid_x", "smem_bytes", "fixed_bits", "deposit")
        d = dict(zip(keys, [a.value for a in v]))
        d["mode"] = "resident" if d["mode"] == L.PIC_MODE_RESIDENT else "streaming"
        d["deposit"] = "split32" if d["deposit"] == L.PIC_DEPOSIT_SPLIT32 else "cas64"
        if d["mode"] == "streaming":
            on, workers = self.coop
            if on:                       # whole steps in one cooperative launch: shared-memory gather in every pass
                d["coop_workers"] = workers
                d["gather"] = "shared"
            else:
                d["gather"] = self.gather
        return d

    def set_tuning(self, threads=0, per_thread=0, ctas_per_sm=-1):
        self._ck(self._lib.pic_set_tuning(self._h, int(threads), int(per_thread), int(ctas_per_sm)))

    def set_gather(self, route="auto"):
        """Streaming mode: where the kick's field gather reads the mesh field -- "shared" (table rebuilt in every CTA's
        prologue), "texture" (one table per sub-stage in global memory, read through the texture pipe), "texture:23"
        (texture for the listed Yoshida stages only) or "auto"."""
        if route.startswith("texture:"):
            code = 0x10 | sum(1 << (int(c) - 1) for c in route[8:])
        else:
            code = {"auto": 0, "shared": 1, "texture": 2}[route]
        self._ck(self._lib.pic_set_gather(self._h, code))

    def set_coop(self, mode="auto"):
        """Streaming mode, one GPU: whole env steps as ONE cooperative launch (grid barriers between the passes, finalize
        on a CTA of its own) -- "on" (raises when the handle's flavour has no such kernel), "off" or "auto" (mid-size
        envs).  Same bits as the kernel-per-pass path."""
        self._ck(self._lib.pic_set_coop(self._h, {"auto": -1, "off": 0, "on": 1}[mode]))

    @property
    def coop(self):
        """(in effect for the next step call, pass CTAs per env)"""
        on, w = C.c_int32(), C.c_int32()
        self._ck(self._lib.pic_get_coop(self._h, C.byref(on), C.byref(w)))
        return bool(on.value), int(w.value)

    @property
    def gather(self):
        r = C.c_int32()
        self._ck(self._lib.pic_get_gather(self._h, C.byref(r)))
        if r.value & 0x10:
            return "texture:" + "".join(str(i + 1) for i in range(3) if r.value >> i & 1)
        return {1: "shared", 2: "texture"}[r.value]

    def launch_count(self):
        return int(self._lib.pic_kernel_launch_count(self._h))

    def views(self):
        """Zero-copy device views: x, v as (n_envs, N) with env stride ld; n, E_mesh (n_envs, M); diag (n_envs, 6).

        READ-ONLY by contract.  (`__cuda_array_interface__` has a read-only flag, but torch.as_tensor rejects
        arrays that set it, so it cannot be used to enforce this.)  Density, field, diagnostics and the
        pre-deposited first sub-stage of the next step were built from these x, v: after writing through a view call
        `refresh_fields()`, or load new particles with set_state / set_state_device, which do it themselves."""
        dv = L.PicDeviceViews()
        self._ck(self._lib.pic_get_device_views(self._h, C.byref(dv)))
        es = dv.elem_size
        ts = "<f8" if es == 8 else "<f4"
        return {
            "x": DeviceArray(dv.x, (self.n_envs, self.N), ts, (dv.ld * es, es), self),
            "v": DeviceArray(dv.v, (self.n_envs, self.N), ts, (dv.ld * es, es), self),
            "n": DeviceArray(dv.n, (self.n_envs, self.M), "<f8", None, self),
            "E_mesh": DeviceArray(dv.E_mesh, (self.n_envs, self.M), "<f8", None, self),
            "diag": DeviceArray(dv.diag, (self.n_envs, L.DIAG_N), "<f8", None, self),
            "ld": int(dv.ld),
        }
