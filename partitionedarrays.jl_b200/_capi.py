"""ctypes binding of libpa_b200.so (the C ABI declared in include/pa_b200.h).

The product path has no CPU fallback: if the CUDA library cannot be built/loaded, or no GPU is
visible when a context is created, the call fails loudly."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_lib = None

PA_SPMV_DEFAULT = 0
PA_SPMV_EXPLICIT_EXCHANGE = 1
PA_SPMV_SKIP_GHOST_REFRESH = 2
PA_CG_REFERENCE_OPS = 4
PA_SPMV_INLINE_PEER_LOADS = 8
PA_SPMV_OVERLAP = 16
PA_SPMV_FUSED_EXCHANGE = 32
PA_CG_TIMING = 64
PA_GS_LEXICOGRAPHIC, PA_GS_MULTICOLOR = 0, 1
PA_OP_SUM, PA_OP_MAX, PA_OP_MIN, PA_OP_ABSSUM, PA_OP_ABSMAX, PA_OP_ABSPOW, PA_OP_INSERT = range(7)


class PAError(RuntimeError):
    pass


class CGResult(C.Structure):
    _fields_ = [("iters", C.c_int32), ("converged", C.c_int32), ("residual0", C.c_double), ("residual", C.c_double)]


# name -> argtypes ; every entry point returns int unless listed in _RESTYPE
_P, _I32, _I64, _U32, _U64, _D = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_double
SIGNATURES = {
    "pa_abi_version": [],
    "pa_last_error": [],
    "pa_ctx_create": [_I32, _I32, _P, _I32, _U64, _P, _P],
    "pa_ctx_create_multi": [_I32, _P, _U64, _P],
    "pa_ctx_destroy": [_P],
    "pa_ctx_sync": [_P],
    "pa_ctx_stream": [_P, _P],
    "pa_ctx_launch_count": [_P, _P],
    "pa_ctx_arena_export": [_P, _I32, _P],
    "pa_ctx_arena_import": [_P, _I32, _P],
    "pa_nccl_unique_id": [_P],
    "pa_ctx_nccl_init": [_P, _P, _I32, _I32],
    "pa_plan_create": [_P, _P],
    "pa_plan_set_part": [_P, _I32, _I64, _I64, _P, _P, _I32, _P, _P, _P, _P, _I32, _P, _P, _P, _P],
    "pa_plan_commit": [_P, _I64],
    "pa_plan_destroy": [_P],
    "pa_vec_create": [_P, _P],
    "pa_vec_destroy": [_P],
    "pa_vec_upload": [_P, _I32, _P, _I64],
    "pa_vec_download": [_P, _I32, _P, _I64],
    "pa_vec_upload_async": [_P, _I32, _P, _I64, _P],
    "pa_vec_download_async": [_P, _I32, _P, _I64, _P],
    "pa_vec_fill": [_P, _D],
    "pa_vec_copy": [_P, _P],
    "pa_vec_scale": [_P, _D],
    "pa_vec_axpby": [_P, _D, _P, _D],
    "pa_vec_waxpby": [_P, _D, _P, _D, _P],
    "pa_vec_dot": [_P, _P, _P],
    "pa_vec_norm2": [_P, _P],
    "pa_vec_sum": [_P, _P],
    "pa_vec_consistent": [_P],
    "pa_vec_assemble": [_P],
    "pa_vec_assemble_op": [_P, _I32],
    "pa_vec_reduce_parts": [_P, _I32, _D, _P],
    "pa_sort_perm_u64": [_P, _P, _I64, _P],
    "pa_xchg_create": [_P, _P],
    "pa_xchg_set_elem_size": [_P, _I32],
    "pa_xchg_set_part": [_P, _I32, _I32, _P, _P, _I32, _P, _P, _P],
    "pa_xchg_commit": [_P, _I64],
    "pa_xchg_destroy": [_P],
    "pa_xchg_upload_snd": [_P, _I32, _P, _I64],
    "pa_xchg_exchange": [_P],
    "pa_xchg_download_rcv": [_P, _I32, _P, _I64],
    "pa_vec_fill_hash_box": [_P, _I32, _P, _P, _P, _U64],
    "pa_mat_create": [_P, _P, _P],
    "pa_mat_destroy": [_P],
    "pa_mat_set_csr": [_P, _I32, _I64, _I64, _I32, _I32, _I32, _P, _P, _P],
    "pa_mat_set_csr_split": [_P, _I32, _I64, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P],
    "pa_mat_set_csc": [_P, _I32, _I64, _I64, _I32, _I32, _I32, _P, _P, _P],
    "pa_mat_set_csc_split": [_P, _I32, _I64, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P],
    "pa_mat_set_coo": [_P, _I32, _I64, _I32, _P, _P, _P],
    "pa_mat_update_coo_values": [_P, _I32, _P, _I64],
    "pa_mat_set_stencil": [_P, _I32, _I32, _P, _P, _P, _I64, _P, _P, _P],
    "pa_mat_commit": [_P],
    "pa_mat_nnz": [_P, _I32, _P],
    "pa_mat_nrows": [_P, _I32, _P],
    "pa_mat_download_csr": [_P, _I32, _P, _P, _P],
    "pa_mat_fill_stored": [_P, _D],
    "pa_mat_spmm_local": [_P, _P, _P],
    "pa_mat_transpose_local": [_P, _P],
    "pa_spmv": [_P, _P, _P, _D, _D, _U32],
    "pa_spmv_transpose": [_P, _P, _P, _D, _D],
    "pa_cg": [_P, _P, _P, _I32, _D, _U32, _P, _P],
    "pa_cg_timings": [_P, _P],
    "pa_ctx_release_workspaces": [_P],
    "pa_gs_create": [_P, _P],
    "pa_gs_set_box": [_P, _I32, _I32, _P],
    "pa_gs_commit": [_P],
    "pa_gs_set_order": [_P, _I32],
    "pa_gs_destroy": [_P],
    "pa_gs_smooth": [_P, _P, _P, _I32],
    "pa_mg_create": [_I32, _P, _P, _P, _P],
    "pa_mg_destroy": [_P],
    "pa_mg_apply": [_P, _P, _P],
    "pa_cg_precond": [_P, _P, _P, _P, _I32, _D, _U32, _P, _P],
    "pa_local_spmv": [_P, _I32, _I32, _I32, _I32, _I64, _I64, _P, _P, _P, _P, _I64, _P],
    "pa_host_alloc": [_P, C.c_size_t],
    "pa_host_free": [_P],
    # not in the public header: tuning knob used by bench/tests
    "pa_ctx_set_knob": [_P, C.c_char_p, _I64],
}
_RESTYPE = {"pa_last_error": C.c_char_p}


def library_path() -> str:
    return _build.SO


def lib():
    """Load (building if needed) libpa_b200.so.  Raises if the CUDA library is unavailable."""
    global _lib
    if _lib is None:
        path = _build.build() if (os.environ.get("PA_B200_NO_BUILD") != "1") else _build.SO
        if not os.path.exists(path):
            raise PAError(f"libpa_b200.so not found at {path}: build it with nvcc (python -m ... build) — no CPU fallback exists")
        L = C.CDLL(path, mode=C.RTLD_GLOBAL)
        for name, args in SIGNATURES.items():
            f = getattr(L, name)
            f.argtypes = args
            f.restype = _RESTYPE.get(name, C.c_int)
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise PAError(f"pa_b200 error {rc}: {lib().pa_last_error().decode()}")


def ptr(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)
