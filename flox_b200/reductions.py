"""Named reductions -> kernel passes + typed epilogue.

The names are the reference's engine-function names (flox/aggregate_flox.py:210-266 and the
chunk funcs listed in flox/aggregations.py:304-632); `evaluate` is what both the standalone engine
functions (aggregate_b200.py) and the fused chunk driver (core.chunk_reduce) call.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib as L
from .accumulate import GroupedAccumulators
from .arrays import torch_dtype

_SUMLIKE = {
    "sum": (L.OP_SUM, False), "nansum": (L.OP_SUM, True),
    "prod": (L.OP_PROD, False), "nanprod": (L.OP_PROD, True),
    "sum_of_squares": (L.OP_SUMSQ, False), "nansum_of_squares": (L.OP_SUMSQ, True),
}
_MINMAX = {"max": (True, False), "nanmax": (True, True), "min": (False, False), "nanmin": (False, True)}
_ARG = {"argmax": (True, False), "nanargmax": (True, True), "argmin": (False, False), "nanargmin": (False, True)}
_PICK = {"first": (False, False), "last": (True, False), "nanfirst": (False, True), "nanlast": (True, True)}
_VAR = {"var": (False, False), "nanvar": (True, False), "std": (False, True), "nanstd": (True, True)}

SUPPORTED = tuple(_SUMLIKE) + tuple(_MINMAX) + tuple(_ARG) + tuple(_VAR) + tuple(_PICK) + ("nanlen", "mean", "nanmean")

_KERNEL_OUT = (torch.float32, torch.float64, torch.int32, torch.int64)


def _is_nan(v) -> bool:
    try:
        return v is not None and math.isnan(float(v))
    except (TypeError, ValueError):
        return False


def _kernel_out_dtype(np_dt: np.dtype) -> torch.dtype:
    np_dt = np.dtype(np_dt)
    if np_dt.kind == "f":
        return torch.float64 if np_dt.itemsize >= 8 else torch.float32
    if np_dt.kind in "iub":
        return torch.int32 if np_dt == np.dtype("int32") else torch.int64
    if np_dt.kind in "mM":
        return torch.int64
    raise TypeError(f"cannot produce dtype {np_dt}")


def evaluate(acc: GroupedAccumulators, func: str, *, out_dtype, fill_value=None, min_count: int = 0, mc_fill=None,
             ddof: int = 0) -> torch.Tensor:
    """Run (or reuse) the pass `func` needs and return the typed flat result [A*B*G] on device.

    `out_dtype` is a numpy dtype.  Absent groups receive `fill_value`; groups with fewer than
    `min_count` non-NaN values receive `mc_fill` (core.py:438-459).
    """
    out_dtype = np.dtype(out_dtype)
    kdt = _kernel_out_dtype(out_dtype)
    if not kdt.is_floating_point:
        for name, v in (("fill_value", fill_value), ("min_count fill", mc_fill if min_count > 0 else None)):
            if _is_nan(v):
                raise ValueError(f"cannot convert float NaN {name} to integer dtype {out_dtype}")
    kw = dict(out_dtype=kdt, fill=fill_value, min_count=min_count, mc_fill=mc_fill)
    as_float = kdt.is_floating_point and not acc.value_dtype.is_floating_point
    fdt = torch.float64 if as_float else None  # accumulator holds f64 for int values feeding a float result
    if func in _SUMLIKE:
        op, skipna = _SUMLIKE[func]
        a, cnt, nnan = acc.moments(op, as_float)
        res = acc.finalize(L.FIN_SUM, a, None, cnt, nnan, skipna=skipna, value_dtype=fdt, **kw)
    elif func == "nanlen":
        cnt, nnan = acc.counts()
        res = acc.finalize(L.FIN_NANLEN, None, None, cnt, nnan, skipna=True, **kw)
    elif func in ("mean", "nanmean"):
        a, cnt, nnan = acc.moments(L.OP_SUM, as_float)
        res = acc.finalize(L.FIN_MEAN, a, None, cnt, nnan, skipna=func == "nanmean", value_dtype=fdt, **kw)
    elif func in _MINMAX:
        is_max, skipna = _MINMAX[func]
        key, cnt, nnan = acc.minmax(is_max)
        res = acc.finalize(L.FIN_MAX if is_max else L.FIN_MIN, key, None, cnt, nnan, skipna=skipna, **kw)
    elif func in _ARG:
        is_max, skipna = _ARG[func]
        idx = acc.argext(is_max, skipna)
        _, cnt, nnan = acc.minmax(is_max)
        res = acc.finalize(L.FIN_ARG, idx, None, cnt, nnan, skipna=skipna, **kw)
    elif func in _PICK:
        is_last, skipna = _PICK[func]
        skipna = skipna and acc.value_dtype.is_floating_point
        val, cnt, nnan = acc.picked(is_last, skipna)
        res = acc.finalize(L.FIN_PICK, val, None, cnt, nnan, skipna=skipna, **kw)
    elif func in _VAR:
        skipna, std = _VAR[func]
        m2, s, cnt, nnan = acc.var(skipna)
        res = acc.finalize(L.FIN_STD if std else L.FIN_VAR, m2, s, cnt, nnan, skipna=skipna, ddof=ddof,
                           value_dtype=torch.float64, **kw)
    else:
        raise NotImplementedError(f"reduction {func!r} is not implemented by the b200 engine")
    want = torch_dtype(out_dtype) if out_dtype.kind not in "mM" else torch.int64
    if res.dtype != want:
        res = res.to(want)
    return res


def var_triple(acc: GroupedAccumulators, skipna: bool, out_dtype) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """var_chunk's (M2, S, n), all in the float intermediate dtype (aggregations.py:348-389)."""
    kdt = _kernel_out_dtype(np.dtype(out_dtype))
    m2, s, cnt, nnan = acc.var(skipna)
    # M2: NaN-propagating already (NaN deviations add NaN); absent groups hold 0 like fill_value=0
    o_m2 = acc.finalize(L.FIN_SUM, m2, None, cnt, nnan, skipna=True, out_dtype=kdt, fill=0, value_dtype=torch.float64)
    o_s = acc.finalize(L.FIN_SUM, s, None, cnt, nnan, skipna=skipna, out_dtype=kdt, fill=0,
                       value_dtype=None if acc.value_dtype.is_floating_point else torch.float64)
    o_n = acc.finalize(L.FIN_NANLEN, None, None, cnt, nnan, skipna=True, out_dtype=kdt, fill=0)
    return o_m2, o_s, o_n
