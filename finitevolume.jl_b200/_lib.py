"""ctypes binding of libfvb200.so (include/fvb200.h).  No fallbacks: if the shared library
has not been built, or no CUDA device is present, every entry point raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfvb200.so")
UNIQUE_ID_BYTES = 128
NSLOT = 8
PEER_BLOB_BYTES = 128

# every symbol include/fvb200.h declares (tests check the built library exports them all)
SYMBOLS = [
    "fvb_version", "fvb_last_error", "fvb_device_count", "fvb_create", "fvb_destroy",
    "fvb_comm_unique_id", "fvb_comm_init", "fvb_assemble", "fvb_update_values", "fvb_sizes",
    "fvb_get_csr", "fvb_get_b", "fvb_get_diag", "fvb_get_freenode", "fvb_get_nodei2freenodei",
    "fvb_get_halo_cols", "fvb_set_halo_plan", "fvb_peer_export", "fvb_peer_import", "fvb_gradient_begin", "fvb_gradient_accumulate",
    "fvb_gradient_end", "fvb_solve", "fvb_spmv", "fvb_vec_upload",
    "fvb_vec_download", "fvb_vec_copy", "fvb_vec_load_b", "fvb_vec_diffnorm", "fvb_set_storage",
    "fvb_step", "fvb_vec_to_nodes", "fvb_time_spmv", "fvb_device_alloc", "fvb_device_free", "fvb_device_copy", "fvb_regulargrid",
    "fvb_nodehycos2neighborhycos", "fvb_set_preconditioner", "fvb_get_preconditioner", "fvb_set_spmv_format", "fvb_get_spmv_format", "fvb_set_pcg_scaling", "fvb_get_pcg_scaling", "fvb_set_profiling", "fvb_get_timings", "fvb_sync",
    "fvb_set_assembly", "fvb_get_assembly", "fvb_assemble_regulargrid",
    "fvb_solve_shifted", "fvb_integrate", "fvb_multi_create", "fvb_multi_destroy", "fvb_multi_set_preconditioner", "fvb_multi_assemble",
    "fvb_multi_assemble_regulargrid", "fvb_multi_sizes", "fvb_multi_solve", "fvb_multi_get_csr", "fvb_multi_get_b",
    "fvb_multi_get_freenode", "fvb_multi_device_handle",
]


class Timings(C.Structure):
    _fields_ = [("h2d_ms", C.c_double), ("assemble_ms", C.c_double), ("solve_ms", C.c_double),
                ("d2h_ms", C.c_double), ("spmv_ms_total", C.c_double), ("spmv_samples", C.c_int64),
                ("kernel_launches", C.c_int64)]


GETB_FN = C.CFUNCTYPE(None, C.c_double, C.POINTER(C.c_double), C.c_void_p)
STEP_CALLBACK_FN = C.CFUNCTYPE(None, C.c_double, C.c_double, C.c_void_p)


class IntegrateOptions(C.Structure):
    _fields_ = [("atol", C.c_double), ("dt0", C.c_double), ("fixed_step", C.c_int), ("adjoint", C.c_int),
                ("rtol", C.c_double), ("maxiter", C.c_int64), ("getb", GETB_FN), ("getb_ctx", C.c_void_p),
                ("callback", STEP_CALLBACK_FN), ("callback_ctx", C.c_void_p)]


class FVBError(RuntimeError):
    """Non-zero fvb_status.  `.status` holds the code (1 = bad input: the cases where the
    reference calls error())."""

    def __init__(self, status, msg):
        super().__init__(msg)
        self.status = status


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  finitevolume.jl_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        L.fvb_last_error.restype = C.c_char_p
        for name in SYMBOLS:
            if name != "fvb_last_error":
                getattr(L, name).restype = C.c_int
        _lib = L
    return _lib


def check(status):
    if status != 0:
        raise FVBError(status, lib().fvb_last_error().decode("utf-8", "replace"))


def ptr(a):
    """void* of a numpy array (must be C-contiguous), an int device address, or None."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    assert a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)
