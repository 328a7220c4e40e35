"""ctypes binding of the C-ABI in ``include/ssfm_b200.h`` (the only way Python reaches the kernels).

There is NO CPU fallback: if the shared library is missing or no CUDA device is present the
accelerated devices raise immediately.
"""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SSFM_B200_LIB") or os.path.join(HERE, "_ssfm_b200.so")

SSFM_OK, SSFM_ERR_INVALID, SSFM_ERR_CUDA, SSFM_ERR_UNSUPPORTED, SSFM_ERR_NOMEM = 0, -1, -2, -3, -4
SSFM_C64, SSFM_C128 = 0, 1
ABI_VERSION = 1


class FiberParams(ctypes.Structure):
    """struct ssfm_fiber_params"""
    _fields_ = [
        ("dt_s", ctypes.c_double), ("length_km", ctypes.c_double), ("alpha_db_km", ctypes.c_double),
        ("beta2_ps2_km", ctypes.c_double), ("beta3_ps3_km", ctypes.c_double), ("gamma_w_km", ctypes.c_double),
        ("phi_max_rad", ctypes.c_double), ("h_km", ctypes.c_double),
    ]


# name -> (restype, argtypes); mirrors include/ssfm_b200.h one to one
PROTOTYPES = {
    "ssfm_abi_version": (ctypes.c_int, []),
    "ssfm_last_error": (ctypes.c_char_p, []),
    "ssfm_plan_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int64, ctypes.c_int32,
                                        ctypes.c_int64, ctypes.c_int32, ctypes.c_int32]),
    "ssfm_plan_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "ssfm_plan_set_option": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64]),
    "ssfm_plan_get_option": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int64)]),
    "ssfm_peek_state": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_double),
                                       ctypes.POINTER(ctypes.c_int32)]),
    "ssfm_propagate": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(FiberParams),
                                      ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p]),
    "ssfm_get_state": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_void_p]),
    "ssfm_copy_state_async": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "ssfm_propagate_streamed": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                               ctypes.c_int64, ctypes.c_void_p]),
    "ssfm_stream_write_u32": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32]),
    "ssfm_stream_wait_geq_u32": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32]),
    "ssfm_get_step_log": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]),
    "ssfm_get_last_timing": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32),
                                            ctypes.POINTER(ctypes.c_float)]),
    "ssfm_launch_count": (ctypes.c_int64, []),
    "ssfm_time_step_kernels": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(FiberParams),
                                              ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]),
    "ssfm_apply_transfer": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "ssfm_fiber_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.POINTER(FiberParams), ctypes.c_void_p]),
    "ssfm_long_plan_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                             ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]),
    "ssfm_long_begin": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(FiberParams), ctypes.c_void_p]),
    "ssfm_long_pmax": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.c_int32, ctypes.c_void_p]),
    "ssfm_long_ctrl": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]),
    "ssfm_long_outer": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]),
    "ssfm_long_inner": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "ssfm_long_p2p_export": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "ssfm_long_p2p_import": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "ssfm_long_p2p_copy": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]),
    "ssfm_long_xbar": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "ssfm_filtfilt_sos": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                         ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]),
    "ssfm_pd_lpf": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                   ctypes.c_int64, ctypes.c_int32, ctypes.c_int64, ctypes.c_double, ctypes.c_double,
                                   ctypes.c_double, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_int64,
                                   ctypes.c_int32, ctypes.c_void_p]),
    "ssfm_gaussian_noise": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_double, ctypes.c_uint64,
                                           ctypes.c_uint32, ctypes.c_int32, ctypes.c_void_p]),
    "ssfm_welch_psd": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                      ctypes.c_void_p]),
    "ssfm_edfa": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                 ctypes.c_int64, ctypes.c_double, ctypes.c_double, ctypes.c_uint64, ctypes.c_int64, ctypes.c_int32,
                                 ctypes.c_void_p]),
}

_lib = None


class ExtensionMissing(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load ``_ssfm_b200.so`` (built by ``python -m opticomlib_b200.build``); raise if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ExtensionMissing(
            "opticomlib_b200: CUDA extension %s is not built (run `python -m opticomlib_b200.build`); "
            "there is no CPU fallback for FIBER/DBP/LPF/BPF." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    if lib.ssfm_abi_version() != ABI_VERSION:
        raise ExtensionMissing("opticomlib_b200: ABI version mismatch, rebuild the extension")
    _lib = lib
    return lib


def check(rc: int) -> None:
    """Map a C return code to the exception type the reference raises for the same condition."""
    if rc == SSFM_OK:
        return
    msg = (load().ssfm_last_error() or b"").decode("utf-8", "replace")
    if rc in (SSFM_ERR_INVALID, SSFM_ERR_UNSUPPORTED):
        raise ValueError(msg)
    if rc == SSFM_ERR_NOMEM:
        raise MemoryError(msg)
    raise RuntimeError("CUDA error in opticomlib_b200: " + msg)
