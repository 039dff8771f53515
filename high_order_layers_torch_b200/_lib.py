"""ctypes binding of include/hol_b200.h (the C-ABI shared library csrc/libhol_b200.so).

There is no CPU implementation and no fallback: if the library is missing or a call fails the
caller gets an exception.
"""
from __future__ import annotations

import ctypes
import os
import threading

from ._build import LIB_PATH

HOL_MAX_N = 10
HOL_MAX_SEGMENTS = 4096
HOL_WS_PREPARED = 1
HOL_GY_STATS_VALID = 2
HOL_COUNT_FUSED_NORM = 4
# hol_set_plan_overrides bits (diagnostic: tests / A-B runs force a kernel family)
HOL_PLAN_DISABLE_TENSOR_CORES = 1
HOL_PLAN_DISABLE_DENSE_DW = 2
HOL_PLAN_NO_SHARED_PHI = 4
HOL_PLAN_NO_CLUSTER = 16
HOL_PLAN_FLUSH_SHIFT = 8

_c_f32p = ctypes.c_void_p
_lock = threading.Lock()
_lib = None

# every symbol include/hol_b200.h declares: (restype, argtypes)
_i32, _i64, _f32, _vp, _sz, _u32 = (ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p,
                                    ctypes.c_size_t, ctypes.c_uint32)
SIGNATURES = {
    "hol_abi_version": (ctypes.c_int, []),
    "hol_error_string": (ctypes.c_char_p, [ctypes.c_int]),
    "hol_set_plan_overrides": (_u32, [_u32]),
    "hol_pw_workspace_bytes": (_sz, [_i64, _i32, _i32, _i32, _i32, _i32]),
    "hol_pw_segment_index_i64": (ctypes.c_int, [_vp, _vp, _i64, _i32, _f32, _f32, _vp]),
    "hol_pw_forward_f32": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _f32, _i32, _f32,
                                          _vp, _sz, _vp]),
    "hol_pw_forward_norm_f32": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _f32, _i64, _i32, _i32, _i32, _i32, _f32,
                                               _i32, _f32, _vp, _sz, _vp]),
    "hol_pw_norm_backward_f32": (ctypes.c_int, [_i32, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _sz,
                                                _vp]),
    "hol_pw_backward_f32": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _f32, _i32,
                                           _f32, _vp, _sz, _u32, _vp]),
    "hol_pw_conv2d_workspace_bytes": (_sz, [_i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32]),
    "hol_pw_conv2d_forward_f32": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32,
                                                 _i32, _i32, _i32, _f32, _i32, _f32, _vp, _sz, _vp]),
    "hol_pw_conv2d_backward_f32": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32,
                                                  _i32, _i32, _i32, _i32, _i32, _f32, _i32, _f32, _vp, _sz, _u32,
                                                  _vp]),
    "hol_pw_conv_workspace_bytes": (_sz, [_i64, _vp, _i32, _i32, _i32, _i32]),
    "hol_pw_conv_forward_f32": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _i32, _i32, _i32, _f32, _i32, _f32,
                                               _vp, _sz, _vp]),
    "hol_pw_conv_backward_f32": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _i32, _i32, _i32, _f32,
                                                _i32, _f32, _vp, _sz, _u32, _vp]),
    "hol_fold_rows_f32": (ctypes.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "hol_unfold_rows_f32": (ctypes.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "hol_p2p_flags_bytes": (_sz, []),
    "hol_p2p_allreduce_f32": (ctypes.c_int, [_vp, _vp, _i32, _i32, _i64, _i64, _f32, _vp]),
    "hol_pw_profile_f32": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _f32, _i32,
                                          _f32, _vp, _sz, _vp, _vp]),
    "hol_row_norm_forward_f32": (ctypes.c_int, [_i32, _vp, _vp, _vp, _i64, _i32, _f32, _vp]),
    "hol_row_norm_backward_f32": (ctypes.c_int, [_i32, _vp, _vp, _vp, _vp, _i64, _i32, _vp]),
    "hol_sparse_lion_step_f32": (ctypes.c_int, [_vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _vp]),
    "hol_pw_plan_flags": (ctypes.c_int, [_i64, _i32, _i32, _i32, _i32, _i32]),
    "hol_pw_launch_count": (ctypes.c_int, [ctypes.c_int, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _u32]),
}


class HolError(RuntimeError):
    pass


def load():
    """Load csrc/libhol_b200.so (built by __graft_entry__.build() / _build.build())."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise HolError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        if lib.hol_abi_version() != 1:
            raise HolError(f"ABI version mismatch: library reports {lib.hol_abi_version()}, binding expects 1")
        _lib = lib
    return _lib


def check(code: int, what: str):
    if code != 0:
        msg = load().hol_error_string(code).decode()
        if code == -2:
            raise ValueError(f"{what}: {msg}")
        raise HolError(f"{what} failed with code {code}: {msg}")
