"""ctypes binding of the C ABI declared in include/varycoef_b200.h.

The product path has no CPU fallback: if the CUDA library is missing (not built) or no
CUDA device is usable, every entry point raises -- it never routes to numpy.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VC_LIB_PATH") or os.path.join(HERE, "libvarycoef_b200.so")

VC_OK, VC_ERR_NOT_PD, VC_ERR_INVALID, VC_ERR_CUDA, VC_ERR_OOM = 0, 1, 2, 3, 4
COV_IDS = {"exp": 0, "mat32": 1, "mat52": 2, "sph": 3, "wend1": 4, "wend2": 5}
DIST_IDS = {"euclidean": 0, "maximum": 1, "manhattan": 2, "minkowski": 3}
TAPER_IDS = {None: 0, "wend1": 1, "wend2": 2}
VC_MU_GLS, VC_MU_FIXED = 0, 1
VC_FLAG_NO_LOOKAHEAD = 1
VC_FLAG_LEGACY_SYRK = 2
VC_FLAG_LEGACY_POTRF = 4


class VcConfig(C.Structure):
    _fields_ = [
        ("cov_id", C.c_int32), ("dist_method", C.c_int32), ("minkowski_p", C.c_double),
        ("taper_id", C.c_int32), ("device", C.c_int32), ("taper_range", C.c_double),
        ("nb_outer", C.c_int32), ("flags", C.c_int32),
    ]


class VcTiming(C.Structure):
    _fields_ = [("build_ms", C.c_float), ("chol_ms", C.c_float), ("finalize_ms", C.c_float),
                ("total_ms", C.c_float)]


class VcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("varycoef_b200 error %d: %s" % (code, msg))
        self.code = code


class NotPositiveDefinite(VcError, np.linalg.LinAlgError):
    """base::chol's "the leading minor of order k is not positive"."""


_dp = C.POINTER(C.c_double)
_lib = None

# name -> (restype, argtypes); must list every symbol include/varycoef_b200.h declares
SIGNATURES = {
    "vc_config_default": (C.c_int, [C.POINTER(VcConfig)]),
    "vc_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(VcConfig), C.c_int64, C.c_int32, C.c_int32,
                            C.c_int32, _dp, _dp, _dp, _dp]),
    "vc_create_multi": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(VcConfig), C.c_int64, C.c_int32, C.c_int32,
                                  C.c_int32, _dp, _dp, _dp, _dp, C.POINTER(C.c_int32), C.c_int32]),
    "vc_device_count": (C.c_int32, [C.c_void_p]),
    "vc_set_data": (C.c_int, [C.c_void_p, _dp, _dp, _dp, _dp]),
    "vc_destroy": (None, [C.c_void_p]),
    "vc_n2ll": (C.c_int, [C.c_void_p, _dp, C.c_int32, _dp, _dp, _dp, _dp, _dp]),
    "vc_n2ll_batch": (C.c_int, [C.c_void_p, C.c_int32, _dp, C.c_int32, _dp, _dp, _dp, C.POINTER(C.c_int32)]),
    "vc_sigma": (C.c_int, [C.c_void_p, _dp, _dp]),
    "vc_sigma_device": (C.c_int, [C.c_void_p, _dp, C.c_void_p]),
    "vc_get_chol": (C.c_int, [C.c_void_p, _dp]),
    "vc_get_whitened": (C.c_int, [C.c_void_p, _dp]),
    "vc_gls_chol": (C.c_int, [C.c_int32, C.c_int64, C.c_int32, _dp, _dp, _dp, _dp]),
    "vc_median_dist": (C.c_int, [C.c_int32, C.c_int64, C.c_int32, _dp, C.c_int32, C.c_double, _dp, C.POINTER(C.c_int32)]),
    "vc_post": (C.c_int, [C.c_void_p, _dp, _dp, _dp, C.POINTER(C.c_double)]),
    "vc_predict": (C.c_int, [C.c_void_p, _dp, _dp, C.c_int64, _dp, _dp, _dp, _dp, _dp, _dp]),
    "vc_timings": (C.c_int, [C.c_void_p, C.POINTER(VcTiming)]),
    "vc_last_info": (C.c_int, [C.c_void_p]),
    "vc_last_error": (C.c_char_p, [C.c_void_p]),
    "vc_padded_n": (C.c_int64, [C.c_void_p]),
    "vc_launch_count": (C.c_int64, [C.c_void_p]),
    "vc_abi_version": (C.c_int, []),
    "vc_probe_dmma_peak": (C.c_int, [C.c_int32, C.c_int32, _dp]),
    "vc_probe_hbm_write": (C.c_int, [C.c_int32, C.c_int64, C.c_int32, _dp]),
    "vc_nccl_unique_id": (C.c_int, [C.c_char_p]),
    "vc_create_dist": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(VcConfig), C.c_int64, C.c_int32, C.c_int32,
                                 C.c_int32, _dp, _dp, _dp, _dp, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                 C.c_char_p]),
    "vc_dist_traffic": (C.c_int, [C.c_void_p, _dp, _dp]),
    "vc_layout_describe": (C.c_int, [C.c_int64, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                     C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32]),
    "vc_plan_batch": (C.c_int, [C.c_int32, C.c_int32, _dp, C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32),
                                C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "vc_bc_owner": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "vc_bc_local": (C.c_int, [C.c_int32, C.c_int32]),
    "vc_bc_count": (C.c_int, [C.c_int32, C.c_int32, C.c_int32]),
    "vc_bc_panel_layout": (C.c_int, [C.c_int32] * 5 + [C.POINTER(C.c_int32)] * 3),
    "vc_bc_trailing_tiles": (C.c_int, [C.c_int32] * 7 + [C.POINTER(C.c_int32), C.c_int32]),
}


def lib():
    """Load libvarycoef_b200.so (raises if it has not been built: no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s is missing: build it with `python -m varycoef_b200._build` "
                "(the product has no CPU path)" % LIB_PATH)
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def as_f64(a, ndim=None):
    a = np.asarray(a, dtype=np.float64)
    if ndim == 2 and a.ndim == 1:
        a = a[:, None]
    return np.asfortranarray(a)


def ptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


def check(code, handle=None):
    if code == VC_OK:
        return
    msg = lib().vc_last_error(handle)
    msg = msg.decode() if msg else ""
    if code == VC_ERR_NOT_PD:
        raise NotPositiveDefinite(code, msg)
    if code == VC_ERR_OOM:
        raise MemoryError("varycoef_b200: " + msg)
    if code == VC_ERR_INVALID:
        raise ValueError("varycoef_b200: " + (msg or "invalid argument"))
    raise VcError(code, msg)
