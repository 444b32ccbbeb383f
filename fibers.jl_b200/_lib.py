"""ctypes binding of libfibers_cuda.so (include/fibers_cuda.h).

The library is built in-tree by `fibers.jl_b200/build.py`; if it is missing it is built on
first use (nvcc must be present).  There is NO CPU fallback: every compute entry point fails
loudly (`FibersCudaError`) when the library or a CUDA device is unavailable.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# FIBERS_CUDA_LIB: load another build of the library (A/B timing of kernel variants inside one GPU session; the Julia wrapper
# honours the same variable)
LIB_PATH = os.environ.get("FIBERS_CUDA_LIB") or os.path.join(HERE, "libfibers_cuda.so")

F32, F64, I16, U16, I32, U8 = 0, 1, 2, 3, 4, 5
I8, U32, I64 = 6, 7, 8              # volume I/O only
KERNEL_AUTO, KERNEL_SIMT, KERNEL_TC = 0, 1, 2
_DTYPES = {np.dtype(np.float32): F32, np.dtype(np.float64): F64, np.dtype(np.int16): I16,
           np.dtype(np.uint16): U16, np.dtype(np.int32): I32, np.dtype(np.uint8): U8}

ERR_NAMES = {1: "ARG", 2: "TABLE", 3: "NODEV", 4: "CUDA", 5: "NOMEM"}


class FibersCudaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(msg)
        self.code = code


_lock = threading.Lock()
_lib = None

_p = C.c_void_p
_i = C.c_int
_i64 = C.c_int64
_f = C.c_float

# name -> (restype, argtypes); mirrors include/fibers_cuda.h one to one
SIGNATURES = {
    "fibers_cuda_version": (_i, []),
    "fibers_cuda_device_count": (_i, []),
    "fibers_cuda_last_error": (C.c_char_p, []),
    "fibers_cuda_set_devices": (_i, [_p, _i]),
    "fibers_cuda_set_kernel": (_i, [_i]),
    "fibers_cuda_release_cache": (None, []),
    "fibers_dti_fit": (_i, [_p, _p, _i, _i, _i, _i, _p, _p] + [_p] * 11 + [_i]),
    "fibers_adc_fit": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p, _i]),
    "fibers_gqi_rec": (_i, [_p, _i, _p, _i, _i, _i, _i, _p, _p, _p, _i, _p, _i, _f] + [_p] * 8 + [_i]),
    "fibers_dsi_rec": (_i, [_p, _i, _p, _i, _i, _i, _i, _p, _p, _p, _i, _p, _i, _i] + [_p] * 9 + [_i]),
    "fibers_dti_gqi_fit": (_i, [_p, _p, _i, _i, _i, _i, _p, _p] + [_p] * 10 + [_p, _i, _p, _i, _f] + [_p] * 7 + [_i]),
    "fibers_dti_gqi_fit_batch": (_i, [_i, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _i, _p, _i, _f, _p, _i]),
    "fibers_rumba_rec": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _i, _f, _i, _f, _f, _f, _f, _i, _i, _i, _i] + [_p] * 13 + [_i]),
    "fibers_st_eigen": (_i, [_p] * 6 + [_i, _i, _i, _p, _p, _i]),
    "fibers_st_recon": (_i, [_p, _i, _i, _i, _f, _f, _p, _p, _i]),
    "fibers_st_eigen_device": (_i, [_p] * 6 + [_i64, _p, _p, _p]),
    "fibers_stream": (_i, [_p, _i, _i, _i, _i, _p, _f, _p, _f, _p, _p, _p, _i, _i, _i, _f, _f, _f, _p, _f, _i, _p, _p, _p]),
    "fibers_stream_device": (_i, [_p, _i, _i, _i, _i, _p, _f, _p, _f, _p, _p, _p, _i, _i, _i, _f, _f, _f, _p, _f, _p, _p, _p]),
    "fibers_stream_lcm": (_i, [_p, _i, _i, _i, _i, _p, _f, _p, _f, _p, _p, _p, _i, _i, _i, _f, _f, _p, C.c_double, _i, _i, C.c_uint64, _i, _p, _p, _p]),
    "fibers_stream_fetch_scalars": (_i, [_p, _p]),
    "fibers_stream_fetch": (_i, [_p, _p, _p]),
    "fibers_stream_free": (None, [_p]),
    "fibers_mri_read_info": (_i, [C.c_char_p, _p]),
    "fibers_mri_read_data": (_i, [C.c_char_p, _p, _p, _i64]),
    "fibers_mri_write": (_i, [C.c_char_p, _p, _i, _p, _p, _p, _f, _f, _f, _f, _f, _f, _i]),
    "fibers_trk_write": (_i, [C.c_char_p, _p, _p, _p, _i64, _p, _p]),
    "fibers_trk_write_ex": (_i, [C.c_char_p, _p, _p, _p, _i64, _p, _p, _i, _p, _i, _p]),
    "fibers_trk_read_info": (_i, [C.c_char_p, _p]),
    "fibers_trk_read_data": (_i, [C.c_char_p, _p, _p, _p, _p, _p]),
    "fibers_cuda_host_register": (_i, [_p, C.c_size_t]),
    "fibers_cuda_host_unregister": (_i, [_p]),
    "fibers_dti_plan_create": (_i, [_p, _i, _i, _p, _p]),
    "fibers_adc_plan_create": (_i, [_p, _i, _i, _p]),
    "fibers_gqi_plan_create": (_i, [_p, _i, _i, _p, _p, _p, _i, _p, _i, _f]),
    "fibers_dsi_plan_create": (_i, [_p, _i, _i, _p, _p, _p, _i, _p, _i, _i]),
    "fibers_plan_destroy": (None, [_p]),
    "fibers_plan_matrix": (_i, [_p, _p, _i64]),
    "fibers_plan_kernel": (_i, [_p]),
    "fibers_dti_fit_device": (_i, [_p, _p, _i64, _p, _i64, _i64] + [_p] * 11 + [_p]),
    "fibers_adc_fit_device": (_i, [_p, _p, _i64, _p, _i64, _p, _p, _p]),
    "fibers_recon_device": (_i, [_p, _p, _i64, _p, _i64, _i64] + [_p] * 10 + [_i, _p]),
    "fibers_stats_init_device": (_i, [_p, _p]),
    "fibers_qa_scale_device": (_i, [_p, _p, _p, _i64, _p, _f, _p]),
    "fibers_stats_decode_max": (_f, [C.c_int32]),
    "fibers_cuda_launch_count": (_i64, []),
    "fibers_host_build_matrix": (_i, [_i, _i, _p, _p, _p, _i, _f, _i, _p, _i64, _p, _p]),
    "fibers_host_build_rumba": (_i, [_i, _p, _p, _p, _i, _f, _f, _f, _f, _f, _p, _i64, _p, _p]),
    "fibers_host_build_neighbours": (_i, [_p, _i, _i, _p]),
    "fibers_host_partition_slabs": (_i, [_p, _i64, _i, _i, _p]),
}


def lib():
    """Load (building if necessary) libfibers_cuda.so.  Raises if it cannot be had."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                from . import build as _build
                _build.build()
            L = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(L, name)          # AttributeError if the symbol is not exported
                fn.restype = res
                fn.argtypes = args
            _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        msg = lib().fibers_cuda_last_error().decode("utf-8", "replace")
        raise FibersCudaError(rc, msg)


def ptr(a):
    """Host pointer of a numpy array (None -> NULL)."""
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def dtype_code(dt) -> int:
    try:
        return _DTYPES[np.dtype(dt)]
    except KeyError:
        raise FibersCudaError(1, f"unsupported dwi element type {dt}")


def device_count() -> int:
    return int(lib().fibers_cuda_device_count())


def require_device():
    if device_count() <= 0:
        raise FibersCudaError(3, "no CUDA device available (libfibers_cuda has no CPU fallback)")
