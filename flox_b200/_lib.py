"""ctypes binding of libflox_b200.so (the C ABI declared in include/flox_b200.h).

There is NO fallback: if the shared library is missing or CUDA is unavailable the first kernel
call raises.  Build with `python -m flox_b200.build` (or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libflox_b200.so")

F32, F64, I32, I64 = 0, 1, 2, 3
OP_SUM, OP_PROD, OP_SUMSQ = 0, 1, 2
FIN_SUM, FIN_NANLEN, FIN_MEAN, FIN_MAX, FIN_MIN, FIN_VAR, FIN_STD, FIN_ARG, FIN_PICK = range(9)
COMB_SUM, COMB_MAX, COMB_MIN, COMB_PROD, COMB_NANMAX, COMB_NANMIN = range(6)


class Chunk(C.Structure):
    _fields_ = [("values", C.c_void_p), ("codes", C.c_void_p), ("n", C.c_int64)]


class FloxB200Error(RuntimeError):
    pass


_i64, _i32, _vp, _f64 = C.c_int64, C.c_int, C.c_void_p, C.c_double
_chunks_p = C.POINTER(Chunk)

# name -> (restype, argtypes); must list EVERY symbol include/flox_b200.h declares
SIGNATURES = {
    "floxb_version": (_i32, []),
    "floxb_strerror": (C.c_char_p, [_i32]),
    "floxb_last_cuda_error": (C.c_char_p, []),
    "floxb_smem_group_limit": (_i64, [_i32, _i32]),
    "floxb_max_chunks_per_launch": (_i32, []),
    "floxb_moments": (_i32, [_chunks_p, _i64, _i32, _i32, _i64, _i64, _i64, _i32, _vp, _vp, _vp, _i32, _i32, _vp, _i64, _vp]),
    "floxb_moments_workspace_bytes": (_i64, [_i32, _i64, _i64, _i64]),
    "floxb_minmax": (_i32, [_chunks_p, _i64, _i32, _i32, _i64, _i64, _i64, _i32, _vp, _vp, _vp, _i32, _i32, _vp]),
    "floxb_group_index": (_i32, [_vp, _i32, _i64, _i64, _vp, _vp, _vp, _vp]),
    "floxb_reduce_cols": (_i32, [_vp, _i32, _vp, _vp, _i64, _i64, _i64, _i64, _i32, _vp, _vp, _vp, _i32, _i32, _vp]),
    "floxb_sqdev": (_i32, [_chunks_p, _i64, _i32, _i32, _i64, _i64, _i64, _vp, _i32, _vp, _i32, _vp]),
    "floxb_group_means": (_i32, [_i32, _vp, _vp, _vp, _i64, _i32, _vp, _vp]),
    "floxb_argext": (_i32, [_chunks_p, _i64, _i32, _i32, _i64, _i64, _i64, _vp, _vp, _i32, C.POINTER(_i64), _vp, _i32, _vp]),
    "floxb_firstlast": (_i32, [_chunks_p, _i64, _i32, _i32, _i64, _i64, _i64, _i32, _i32, C.POINTER(_i64), _vp, _i32, _vp]),
    "floxb_take": (_i32, [_chunks_p, _i64, _i32, _i32, _i64, _i64, _i64, _vp, C.POINTER(_i64), _vp, _i32, _vp]),
    "floxb_finalize": (_i32, [_i32, _i32, _i32, _vp, _vp, _vp, _vp, _i64, _i32, _f64, _i64, _i64, _f64, _i64, _i64, _vp, _i64, _vp]),
    "floxb_combine": (_i32, [_i32, _i32, _vp, _i64, _i64, _vp, _vp]),
    "floxb_var_combine": (_i32, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp]),
    "floxb_digitize": (_i32, [_vp, _i32, _i64, _vp, _i64, _i32, _vp, _vp]),
    "floxb_ravel_codes": (_i32, [_vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    "floxb_label_minmax": (_i32, [_vp, _i32, _i64, _vp, _vp]),
    "floxb_label_presence": (_i32, [_vp, _i32, _i64, _i64, _i64, _vp, _vp]),
    "floxb_lut_codes": (_i32, [_vp, _i32, _i64, _i64, _i64, _vp, _vp, _vp]),
    "floxb_host_pack_codes": (_i32, [_vp, _i32, _i64, _i64, _vp, _i32]),
    "floxb_widen_codes": (_i32, [_vp, _i64, _vp, _vp]),
    "floxb_segment_valid": (_i32, [_vp, _i32, _i64, _i64, _vp, _i64, _vp, _vp]),
    "floxb_segment_quantile": (_i32, [_vp, _i32, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _i32, _i32, _f64, _i32, _vp, _vp]),
    "floxb_segment_scan_scratch_bytes": (_i64, [_i64, _i64]),
    "floxb_segment_scan": (_i32, [_i32, _vp, _i32, _vp, _vp, _i64, _i64, _i32, _vp, _i64, _vp, _vp]),
    "floxb_ipc_alloc": (_i32, [_i64, C.POINTER(_vp), _vp]),
    "floxb_ipc_open": (_i32, [_vp, C.POINTER(_vp)]),
    "floxb_ipc_close": (_i32, [_vp]),
    "floxb_ipc_free": (_i32, [_vp]),
    "floxb_xchg_allreduce": (_i32, [C.POINTER(_vp), _i32, _i32, C.c_uint64, _i64, _i32, C.POINTER(_i64), C.POINTER(_i32),
                                    C.POINTER(_vp), _vp]),
    "floxb_xchg_status": (_i32, [_vp, C.POINTER(C.c_uint64)]),
}

SCAN_CUMSUM, SCAN_FFILL = 0, 1
XCHG_MAX_RANKS, XCHG_MAX_SEGMENTS, XCHG_HEADER_BYTES = 8, 4, 128
XCHG_SUM_F64, XCHG_SUM_I64, XCHG_MAX_I64, XCHG_MIN_I64, XCHG_PROD_F64, XCHG_PROD_I64 = range(6)

_lock = threading.Lock()
_lib = None


def load():
    """Load (once) and return the ctypes library handle.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                try:  # in-tree build with nvcc if the artefact did not travel with the checkout
                    from . import build as _build

                    _build.build()
                except Exception:
                    pass
            if not os.path.exists(LIB_PATH):
                raise FloxB200Error(
                    f"{LIB_PATH} not found: the CUDA extension has not been built "
                    "(run `python -m flox_b200.build`). flox_b200 has no CPU fallback."
                )
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        lib = load()
        msg = lib.floxb_strerror(rc).decode()
        if rc == -3:
            msg += ": " + lib.floxb_last_cuda_error().decode()
        raise FloxB200Error(f"{what or 'flox_b200 kernel'} failed ({rc}): {msg}")
