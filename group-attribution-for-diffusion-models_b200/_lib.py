"""ctypes binding of libgadm.so (C ABI declared in include/gadm.h).

There is no fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from .build import LIB_PATH

c_i64 = C.c_int64
c_u64 = C.c_uint64
c_vp = C.c_void_p

_SIGNATURES = {
    "gadm_version": (C.c_int, []),
    "gadm_last_error": (C.c_char_p, []),
    "gadm_create": (C.c_int, [C.POINTER(c_vp), C.c_int]),
    "gadm_destroy": (C.c_int, [c_vp]),
    "gadm_launch_count": (c_i64, [c_vp]),
    "gadm_watchdog_code": (C.c_int, [c_vp, C.POINTER(C.c_uint)]),
    "gadm_set_watchdog_ns": (C.c_int, [c_vp, c_u64]),
    "gadm_project_workspace_bytes": (c_i64, [c_vp, c_i64, c_i64, c_i64, C.c_int]),
    "gadm_pack_block": (C.c_int, [c_vp, c_vp, C.c_int, c_i64, c_i64, c_i64, c_vp, c_i64, c_i64, c_i64, c_i64, C.c_float,
                                  c_vp]),
    "gadm_stage_scale_count": (c_i64, [c_i64]),
    "gadm_stage_rows": (C.c_int, [c_vp, c_vp, C.c_int, C.c_int, c_i64, C.c_float, c_vp, C.c_int, c_i64, c_i64, c_i64, c_vp,
                                  C.c_int, c_vp]),
    "gadm_accumulate_rows": (C.c_int, [c_vp, c_vp, C.c_int, C.c_int, c_i64, C.c_float, c_vp, c_i64, c_i64, c_i64, C.c_int,
                                       c_vp]),
    "gadm_project_staged": (C.c_int, [c_vp, c_vp, C.c_int, c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_u64, C.c_int, c_vp,
                                      c_i64, C.c_int, c_vp, c_i64, C.c_int, c_vp]),
    "gadm_materialize_p": (C.c_int, [c_vp, c_i64, c_i64, c_i64, c_u64, C.c_int, C.c_int, c_vp, c_vp]),
    "gadm_gemm_tn": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_i64, C.c_float, C.c_float,
                               C.c_float, C.c_int, c_vp]),
    "gadm_gemm_tn_batched": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_i64, c_i64, c_i64,
                                       c_i64, C.c_float, C.c_float, C.c_float, C.c_int, c_vp]),
    "gadm_cholesky_solve_vec_workspace_bytes": (c_i64, [c_i64]),
    "gadm_cholesky_solve_vec": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "gadm_transpose": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp]),
    "gadm_cholesky_workspace_bytes": (c_i64, [c_i64]),
    "gadm_cholesky": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, C.POINTER(C.c_int), c_vp]),
    "gadm_tri_inverse_workspace_bytes": (c_i64, [c_i64]),
    "gadm_tri_inverse": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp]),
    "gadm_solve_rows": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_i64, c_vp]),
    "gadm_row_norms": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, C.c_int, c_vp, c_vp]),
    "gadm_matvec_rows": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "gadm_diag_minmax": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp]),
    "gadm_col_mean_scaled": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "gadm_scale_rows_cols": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "gadm_pack_masks": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "gadm_mask_gram": (C.c_int, [c_vp, c_vp, c_i64, c_i64, C.c_int, c_vp, c_vp]),
    "gadm_mask_xty": (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, C.c_double, C.c_double, c_vp, c_vp]),
    "gadm_mask_times_matrix": (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "gadm_sym_pinv_workspace_bytes": (c_i64, [c_i64]),
    "gadm_sym_pinv": (C.c_int, [c_vp, c_vp, c_i64, C.c_double, c_vp, c_vp, c_i64, C.POINTER(C.c_int), c_vp]),
    "gadm_dgemm_dk": (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, C.c_double, c_vp, c_vp]),
    "gadm_center_columns": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "gadm_dgemm": (C.c_int, [c_vp, C.c_int, c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp]),
    "gadm_sym_eig_workspace_bytes": (c_i64, [c_i64]),
    "gadm_sym_eig": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, C.POINTER(C.c_int), c_vp]),
    "gadm_ridge_gcv_workspace_bytes": (c_i64, [c_i64, c_i64, c_i64, c_i64]),
    "gadm_ridge_gcv": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp, c_vp]),
    "gadm_ridge_select": (C.c_int, [c_vp, c_vp, c_i64, c_i64, C.c_int, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "gadm_ridge_intercept": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp]),
    "gadm_datamodel_slot_bytes": (c_i64, [c_i64]),
    "gadm_datamodel_ridge_systems": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "gadm_shapley_rhs": (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "gadm_lds_spearman": (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_vp, c_vp]),
    "gadm_lds_mean": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp]),
    "gadm_group_reduce": (C.c_int, [c_vp, c_vp, C.c_int, c_vp, c_i64, c_i64, C.c_int, c_vp, c_vp]),
    "gadm_stable_rank_desc": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    "gadm_project": (C.c_int, [c_vp, c_vp, C.c_int, C.c_int, c_i64, C.c_float, c_vp, C.c_int, c_vp, c_i64, c_i64, c_i64, c_u64,
                               C.c_int, c_vp, c_i64, C.c_int, c_vp, c_i64, C.c_int, c_vp]),
    "gadm_gram": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp, c_i64, C.c_float, C.c_int, c_vp]),
    "gadm_score": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_vp, c_i64,
                             c_vp, c_i64, c_vp, c_vp, c_vp]),
    "gadm_shapley_workspace_bytes": (c_i64, [c_i64, c_i64]),
    "gadm_shapley": (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp]),
    "gadm_banzhaf": (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp]),
    "gadm_lds": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_i64, c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_vp]),
    "gadm_row_mean": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp]),
}

_lib = None
_lock = threading.Lock()


class Block(C.Structure):
    """gadm_block (include/gadm.h)."""

    _fields_ = [("ptr", c_vp), ("numel_per_example", c_i64), ("example_stride", c_i64), ("row_offset", c_i64)]


class GadmError(RuntimeError):
    """A libgadm call failed (CUDA error, kernel watchdog, workspace too small)."""


def load_library(path: str | None = None):
    """dlopen libgadm.so and attach argument types.  Raises if it is not built."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = path or os.environ.get("GADM_LIBRARY", LIB_PATH)
        if not os.path.exists(path):
            raise ImportError(
                f"{path} not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU or PyTorch fallback for this package.")
        lib = C.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def declared_symbols():
    return sorted(_SIGNATURES)


def last_error() -> str:
    return load_library().gadm_last_error().decode("utf-8", "replace")


def check(rc: int):
    """Map a C return code to the exception the reference-facing API documents."""
    if rc >= 0:
        return rc
    msg = last_error()
    if rc == -1:
        raise ValueError(msg)
    raise GadmError(f"gadm error {rc}: {msg}")


_handles: dict[int, "Handle"] = {}


class Handle:
    """One gadm context per (process, device)."""

    def __init__(self, device_index: int):
        lib = load_library()
        h = c_vp()
        check(lib.gadm_create(C.byref(h), device_index))
        self.lib = lib
        self.ptr = h
        self.device_index = device_index
        if os.environ.get("GADM_WATCHDOG_SEC") is not None:  # 0 disables (e.g. under ncu --set full)
            check(lib.gadm_set_watchdog_ns(h, int(float(os.environ["GADM_WATCHDOG_SEC"]) * 1e9)))

    def launch_count(self) -> int:
        return int(self.lib.gadm_launch_count(self.ptr))

    def watchdog_code(self) -> int:
        code = C.c_uint(0)
        check(self.lib.gadm_watchdog_code(self.ptr, C.byref(code)))
        return int(code.value)


def get_handle(device) -> Handle:
    import torch

    device = torch.device(device)
    if device.type != "cuda":
        raise ValueError(f"gadm kernels run on CUDA devices only (sm_100a); got device '{device}'")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _handles:
        _handles[idx] = Handle(idx)
    return _handles[idx]


def stream_ptr(device=None) -> int:
    import torch

    return int(torch.cuda.current_stream(device).cuda_stream)
