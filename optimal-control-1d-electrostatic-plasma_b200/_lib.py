"""ctypes binding of include/pic_b200.h (lib/libpic_b200.so).

The library is the product; this file only declares its entry points.  There is no CPU fallback: if the shared
library is missing it is built with nvcc (build.py), and if that is impossible the import raises.
"""
import ctypes as C
import os

from . import build as _build

PIC_OK = 0
PIC_ENUMERIC = -6
PIC_F64, PIC_F32 = 0, 1
PIC_MODE_AUTO, PIC_MODE_RESIDENT, PIC_MODE_STREAMING = 0, 1, 2
PIC_DEPOSIT_AUTO, PIC_DEPOSIT_CAS64, PIC_DEPOSIT_SPLIT32 = -1, 0, 1
PIC_INTERP_CIC, PIC_INTERP_TSC = 0, 1
DIAG_KE, DIAG_PE_MESH, DIAG_SUM_V, DIAG_SUM_E2, DIAG_REWARD, DIAG_INPUT_E, DIAG_N = 0, 1, 2, 3, 4, 5, 6
ERR_INDEX_RANGE, ERR_NONFINITE, ERR_COMM_TIMEOUT, ERR_DENSITY_RANGE = 1, 2, 4, 8
ERR_TEXT = {
    ERR_INDEX_RANGE: "a cell index floor(x/dx) fell outside [0, N_mesh) (the reference raises in np.bincount, "
                     "src/env/interpolate.py:16)",
    ERR_NONFINITE: "a non-finite particle position reached the deposit",
    ERR_COMM_TIMEOUT: "a peer rank of the fused density exchange did not arrive in time; the step was abandoned",
    ERR_DENSITY_RANGE: "a cell's fixed-point density overflowed its headroom (density far above the mean): lower "
                       "fixed_bits or use the split32 deposit",
}


class PicConfig(C.Structure):
    _fields_ = [
        ("n_particles", C.c_int64), ("n_particles_total", C.c_int64), ("n_mesh", C.c_int32), ("n_envs", C.c_int32),
        ("n0", C.c_double), ("L", C.c_double), ("dt", C.c_double),
        ("precision", C.c_int32), ("mode", C.c_int32), ("deposit", C.c_int32), ("fixed_bits", C.c_int32),
        ("exact_weights", C.c_int32), ("device", C.c_int32), ("max_mode", C.c_int32), ("stream", C.c_void_p),
        ("interpolation", C.c_int32),
    ]


class PicDeviceViews(C.Structure):
    _fields_ = [("x", C.c_void_p), ("v", C.c_void_p), ("ld", C.c_int64), ("n", C.c_void_p), ("E_mesh", C.c_void_p),
                ("diag", C.c_void_p), ("elem_size", C.c_int32)]


# every symbol include/pic_b200.h declares: name -> (restype, argtypes)
_H = C.c_void_p
_D = C.POINTER(C.c_double)
SIGNATURES = {
    "pic_abi_version": (C.c_int, []),
    "pic_build_info": (C.c_char_p, []),
    "pic_clip_dt": (C.c_double, [C.c_double, C.c_int64, C.c_double]),
    "pic_create": (C.c_int, [C.POINTER(PicConfig), C.POINTER(_H)]),
    "pic_destroy": (C.c_int, [_H]),
    "pic_last_error": (C.c_char_p, [_H]),
    "pic_set_stream": (C.c_int, [_H, C.c_void_p]),
    "pic_get_stream": (C.c_int, [_H, C.POINTER(C.c_void_p)]),
    "pic_set_state": (C.c_int, [_H, C.c_void_p, C.c_void_p]),
    "pic_set_state_device": (C.c_int, [_H, C.c_void_p, C.c_void_p]),
    "pic_sample_state": (C.c_int, [_H, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int32,
                                   C.c_uint64, C.c_int64, C.c_int64, C.c_int64]),
    "pic_get_state": (C.c_int, [_H, C.c_void_p, C.c_void_p]),
    "pic_get_fields": (C.c_int, [_H, C.c_void_p, C.c_void_p]),
    "pic_get_density_fixed": (C.c_int, [_H, C.c_void_p, C.POINTER(C.c_int32)]),
    "pic_get_diag": (C.c_int, [_H, C.c_void_p]),
    "pic_get_diag_flags": (C.c_int, [_H, C.c_void_p, C.POINTER(C.c_uint32)]),
    "pic_get_trace": (C.c_int, [_H, C.c_void_p, C.c_int32]),
    "pic_get_cells": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pic_step_mesh": (C.c_int, [_H, C.c_void_p, C.c_int32]),
    "pic_set_actuator_basis": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int32]),
    "pic_step_coeffs": (C.c_int, [_H, C.c_void_p, C.c_int32]),
    "pic_step_mesh_device": (C.c_int, [_H, C.c_void_p, C.c_int32]),
    "pic_step_coeffs_device": (C.c_int, [_H, C.c_void_p, C.c_int32]),
    "pic_set_reward": (C.c_int, [_H, C.c_double, C.c_double, C.c_double, C.c_double]),
    "pic_enable_modes": (C.c_int, [_H, C.c_int32]),
    "pic_get_modes": (C.c_int, [_H, C.c_void_p]),
    "pic_get_mode_trace": (C.c_int, [_H, C.c_void_p, C.c_int32]),
    "pic_phase_hist_config": (C.c_int, [_H, C.c_double, C.c_double, C.c_int32]),
    "pic_phase_hist": (C.c_int, [_H, C.c_void_p]),
    "pic_set_feq": (C.c_int, [_H, C.c_void_p]),
    "pic_kl_divergence": (C.c_int, [_H, C.c_void_p]),
    "pic_refresh_fields": (C.c_int, [_H]),
    "pic_sync": (C.c_int, [_H]),
    "pic_get_error_flags": (C.c_int, [_H, C.POINTER(C.c_uint32)]),
    "pic_clear_error_flags": (C.c_int, [_H]),
    "pic_get_device_views": (C.c_int, [_H, C.POINTER(PicDeviceViews)]),
    "pic_comm_init": (C.c_int, [_H, C.c_void_p, C.c_int32, C.c_int32]),
    "pic_nccl_unique_id": (C.c_int, [C.c_char_p]),
    "pic_comm_init_rank": (C.c_int, [_H, C.c_char_p, C.c_int32, C.c_int32]),
    "pic_comm_exchange_words": (C.c_int64, [_H, C.c_int32]),
    "pic_comm_init_peer": (C.c_int, [_H, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int64]),
    "pic_comm_set_multicast": (C.c_int, [_H, C.c_void_p]),
    "pic_run_stage": (C.c_int, [_H, C.c_int32]),
    "pic_stage_density": (C.c_int, [_H, C.c_int32, C.POINTER(C.c_void_p)]),
    "pic_set_stage_actuation": (C.c_int, [_H, C.c_void_p, C.c_void_p]),
    "pic_get_launch_info": (C.c_int, [_H] + [C.POINTER(C.c_int32)] * 7),
    "pic_set_tuning": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_int32]),
    "pic_set_gather": (C.c_int, [_H, C.c_int32]),
    "pic_get_gather": (C.c_int, [_H, C.POINTER(C.c_int32)]),
    "pic_set_coop": (C.c_int, [_H, C.c_int32]),
    "pic_get_coop": (C.c_int, [_H, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "pic_kernel_launch_count": (C.c_int64, [_H]),
}

_lib = None


def library_path():
    return _build.LIB


def load(build_if_missing=True):
    """dlopen the C-ABI library (building it first when the sources are newer) and type every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if os.environ.get("PIC_LIB_PATH"):                 # experiments: a differently built library, used as is
        path, build_if_missing = os.environ["PIC_LIB_PATH"], False
    if build_if_missing and (_build.needs_build() if os.path.isdir(_build.CSRC) else not os.path.exists(path)):
        try:
            _build.build_library(verbose=False)
        except Exception as e:  # nvcc missing on a box that got a prebuilt .so is fine; a missing .so is not
            if not os.path.exists(path):
                raise ImportError("libpic_b200.so is missing and could not be built: %s" % e) from e
    if not os.path.exists(path):
        raise ImportError("libpic_b200.so not found at %s (run build.py); there is no CPU fallback" % path)
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)            # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class PicError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("pic_b200 error %d: %s" % (code, msg))
        self.code = code


class PicDeviceError(PicError):
    """Sticky device-side error flags were found set (pic_get_error_flags)."""

    def __init__(self, flags):
        texts = [t for bit, t in ERR_TEXT.items() if flags & bit]
        RuntimeError.__init__(self, "pic_b200 device error flags 0x%x: %s" % (flags, "; ".join(texts) or "unknown"))
        self.code = PIC_ENUMERIC
        self.flags = flags


def check(rc, handle=None):
    if rc != PIC_OK:
        msg = load().pic_last_error(handle)
        raise PicError(rc, msg.decode() if msg else "")
