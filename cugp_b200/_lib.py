"""ctypes binding of libcugp.so (include/cugp.h).  The CUDA library is the product: if it is missing or
no CUDA device is visible, calls fail loudly -- there is no CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcugp.so")

OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_NODEVICE = 0, 1, 2, 3, 4
dp = C.POINTER(C.c_double)
EVAL_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, dp, dp, dp)


class ShardStreamStats(C.Structure):
    """struct cugp_shardstream_stats (include/cugp.h)."""
    _fields_ = [("passes", C.c_long), ("groups", C.c_long), ("shards_parsed", C.c_long), ("cache_hits", C.c_long),
                ("parse_ms", C.c_double), ("reader_wait_ms", C.c_double), ("h2d_bytes", C.c_double),
                ("cache_bytes", C.c_double)]


class CugpError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"cugp error {code}: {msg}")
        self.code = code


_lib = None

# name -> (restype, argtypes); every symbol include/cugp.h declares
SIGNATURES = {
    "cugp_version": (C.c_char_p, []),
    "cugp_last_error": (C.c_char_p, []),
    "cugp_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "cugp_set_device": (C.c_int, [C.c_int]),
    "cugp_launch_count": (C.c_long, []),
    "cugp_launch_count_reset": (None, []),
    "cugp_set_tuning": (C.c_int, [C.c_char_p, C.c_long]),
    "cugp_covsum_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "cugp_covsum_destroy": (C.c_int, [C.c_void_p]),
    "cugp_covsum_set_loghyper": (C.c_int, [C.c_void_p, dp]),
    "cugp_covsum_get_loghyper": (C.c_int, [C.c_void_p, dp]),
    "cugp_covsum_K_train": (C.c_int, [C.c_void_p, dp, dp]),
    "cugp_covsum_k_test": (C.c_int, [C.c_void_p, dp, dp, dp]),
    "cugp_covsum_loglik": (C.c_int, [C.c_void_p, dp, dp, dp]),
    "cugp_covsum_grad": (C.c_int, [C.c_void_p, dp, dp, dp]),
    "cugp_covsum_predict": (C.c_int, [C.c_void_p, dp, dp, dp, C.c_int, dp, dp]),
    "cugp_nlpp": (C.c_int, [dp, dp, dp, C.c_int, dp]),
    "cugp_covsum_cg_solve": (C.c_int, [C.c_void_p, dp, dp, dp, C.c_int, C.POINTER(C.c_int)]),
    "cugp_covsum_rprop_solve": (C.c_int, [C.c_void_p, dp, dp]),
    "cugp_cg_minimize": (C.c_int, [EVAL_FN, C.c_void_p, dp, dp, C.c_int, C.POINTER(C.c_int)]),
    "cugp_rprop_minimize": (C.c_int, [EVAL_FN, C.c_void_p, dp, C.POINTER(C.c_int)]),
    "cugp_covsum_set_data": (C.c_int, [C.c_void_p, dp, dp]),
    "cugp_covsum_loglik_resident": (C.c_int, [C.c_void_p, dp]),
    "cugp_covsum_grad_resident": (C.c_int, [C.c_void_p, dp]),
    "cugp_covsum_scalars_resident": (C.c_int, [C.c_void_p, dp]),
    "cugp_covsum_alpha_resident": (C.c_int, [C.c_void_p, dp]),
    "cugp_covsum_residual_resident": (C.c_int, [C.c_void_p, dp]),
    "cugp_covsum_factorize_resident": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "cugp_covsum_solve_resident": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "cugp_covsum_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "cugp_covsum_profile_read": (C.c_int, [C.c_void_p, dp, dp, C.POINTER(C.c_long)]),
    "cugp_cholesky": (C.c_int, [dp, dp, C.c_int]),
    "cugp_chol_and_det": (C.c_int, [dp, dp, C.c_int, dp, dp]),
    "cugp_kinv_y": (C.c_int, [dp, dp, dp, C.c_int]),
    "cugp_k_inverse": (C.c_int, [dp, dp, C.c_int]),
    "cugp_tri_solve_matrix": (C.c_int, [dp, dp, dp, C.c_int, C.c_int]),
    "cugp_bcm_dims": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "cugp_bcm_create": (C.c_int, [dp, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "cugp_bcm_destroy": (C.c_int, [C.c_void_p]),
    "cugp_bcm_set_loghyper": (C.c_int, [C.c_void_p, dp]),
    "cugp_bcm_get_loghyper": (C.c_int, [C.c_void_p, dp]),
    "cugp_bcm_loglik_grad_local": (C.c_int, [C.c_void_p, C.c_int, dp]),
    "cugp_bcm_local_experts": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), dp]),
    "cugp_bcm_predict_moments_dev": (C.c_int, [C.c_void_p, dp, C.c_int, C.c_void_p]),
    "cugp_bcm_predict_moments": (C.c_int, [C.c_void_p, dp, C.c_int, dp]),
    "cugp_poe_finalize_dev": (C.c_int, [C.c_void_p, C.c_int, dp, dp]),
    "cugp_poe_finalize": (C.c_int, [dp, C.c_int, dp, dp]),
    "cugp_bcm_predict": (C.c_int, [C.c_void_p, dp, C.c_int, dp, dp]),
    "cugp_nccl_unique_id": (C.c_int, [C.POINTER(C.c_ubyte)]),
    "cugp_bcm_comm_init": (C.c_int, [C.c_void_p, C.POINTER(C.c_ubyte)]),
    "cugp_bcm_comm_init_file": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "cugp_bcm_has_comm": (C.c_int, [C.c_void_p]),
    "cugp_bcm_collectives": (C.c_long, [C.c_void_p]),
    "cugp_bcm_exchange_kind": (C.c_int, [C.c_void_p]),
    "cugp_bcm_loglik_grad": (C.c_int, [C.c_void_p, C.c_int, dp]),
    "cugp_shardstream_open_files": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.c_size_t, C.POINTER(C.c_void_p)]),
    "cugp_shardstream_open_memory": (C.c_int, [dp, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                               C.POINTER(C.c_void_p)]),
    "cugp_shardstream_close": (C.c_int, [C.c_void_p]),
    "cugp_shardstream_set_loghyper": (C.c_int, [C.c_void_p, dp]),
    "cugp_shardstream_get_loghyper": (C.c_int, [C.c_void_p, dp]),
    "cugp_shardstream_layout": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "cugp_shardstream_loglik_grad_local": (C.c_int, [C.c_void_p, C.c_int, dp, dp]),
    "cugp_shardstream_predict_moments_dev": (C.c_int, [C.c_void_p, dp, C.c_int, C.c_void_p]),
    "cugp_shardstream_predict_moments": (C.c_int, [C.c_void_p, dp, C.c_int, dp]),
    "cugp_shardstream_parse_file": (C.c_int, [C.c_char_p, C.c_int, C.c_size_t, dp]),
    "cugp_shardstream_get_stats": (C.c_int, [C.c_void_p, C.POINTER(ShardStreamStats)]),
    "cugp_probe_fp64_peak": (C.c_int, [C.c_float, dp, dp]),
    "cugp_probe_dmma": (C.c_int, [C.c_float, dp, dp]),
    "cugp_probe_gemm": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, dp]),
    "cugp_debug_gemm": (C.c_int, [dp, dp, dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int,
                                  C.c_int, C.c_int, dp]),
    "cugp_debug_diag_phases": (C.c_int, [dp, C.POINTER(C.c_longlong), C.c_int]),
    "cugp_debug_step_stamps": (C.c_int, [C.c_void_p, C.POINTER(C.c_longlong), C.c_int, C.POINTER(C.c_float)]),
    "cugp_probe_copy": (C.c_int, [C.c_size_t, C.c_int, dp]),
}


def lib():
    """Load libcugp.so.  Raises if it has not been built: the CUDA extension IS the product."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m cugp_b200.build` (nvcc, sm_100a). "
                "cugp_b200 has no CPU fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int) -> None:
    if rc != OK:
        raise CugpError(rc, lib().cugp_last_error().decode())


def f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def ptr(a: np.ndarray):
    return a.ctypes.data_as(dp)
