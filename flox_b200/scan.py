"""Grouped scans on the GPU: `groupby_scan`, signature and semantics of flox.groupby_scan for in-memory arrays
(reference: flox/scan.py:101-309; the kernels replace aggregate_flox.cumsum / nancumsum / ffill, aggregate_flox.py:269-329).

Each element of the result is the running `cumsum` / `nancumsum` (or, for `ffill` / `bfill`, the last / next valid value)
over the EARLIER elements of its own group along `axis`, in order of appearance; the output has the input's shape and order.
"""
from __future__ import annotations

import numpy as np
import torch

from . import aggregate_b200
from .arrays import is_torch, np_dtype_of
from .factorize import _factorize_multiple

SCANS = ("cumsum", "nancumsum", "ffill", "bfill")


def _flip(x):
    return x.flip(-1) if is_torch(x) else x[..., ::-1]


def groupby_scan(array, *by, func: str, expected_groups=None, axis=-1, dtype=None, method=None, engine=None):
    """GroupBy scan along one axis.  `array` / `by`: numpy arrays or torch tensors; `func` in {"cumsum", "nancumsum",
    "ffill", "bfill"}.  Returns an array shaped like `array`."""
    if engine not in (None, "b200"):
        raise NotImplementedError("Setting `engine` is not supported for scans yet.")          # scan.py:187-188
    if func not in SCANS:
        raise NotImplementedError(f"Scan {func!r} is not implemented")
    if expected_groups is not None:
        raise NotImplementedError("Setting `expected_groups` with scans is not supported yet.")  # scan.py:204-205
    axes = (axis,) if np.isscalar(axis) else tuple(axis)
    if len(axes) > 1:
        raise NotImplementedError("Scans are only supported along a single dimension.")          # scan.py:177-178
    if not is_torch(array):
        array = np.asarray(array)
    (ax,) = axes
    ax = ax + array.ndim if ax < 0 else ax
    if ax != array.ndim - 1:
        raise NotImplementedError("the b200 engine scans along the last axis")
    bys = tuple(b if is_torch(b) else np.asarray(b) for b in by)
    array_dt = np_dtype_of(array)
    if func in ("ffill", "bfill") and array_dt.kind != "f":
        return array                                                                             # scan.py:199-201: no NaNs to fill
    (codes,), _, grp_shape = _factorize_multiple(bys, (None,) * len(bys), sort=False)
    if codes.ndim != 1 or codes.shape[0] != array.shape[-1]:
        raise NotImplementedError("scans group by 1-D labels along the scanned (last) axis")

    cast_to = None
    if array_dt.kind in "Mm":
        cast_to = array_dt
        array = array.view(np.int64)
    elif array_dt.kind == "b":
        array = array.to(torch.int8) if is_torch(array) else array.view(np.int8)
    work_dt = np_dtype_of(array)
    if func in ("cumsum", "nancumsum") and work_dt.kind in "iu":                                 # scan.py:271-278
        out_dt = np.result_type(work_dt, np.int_ if work_dt.kind == "i" else np.uint)
    else:
        out_dt = work_dt if dtype is None else np.dtype(dtype)

    def cast(x):
        if is_torch(x):
            from .arrays import torch_dtype

            x = x.to(torch_dtype(out_dt))
            return x
        x = x.astype(out_dt)
        return x.astype(cast_to) if cast_to is not None else x

    if codes.shape[-1] == 1 or tuple(codes.shape) == tuple(grp_shape):                           # scan.py:284-289
        return cast(array)

    if func == "bfill":                                                                          # reverse, ffill, reverse
        res = _flip(aggregate_b200.ffill(_flip(codes), _flip(array), axis=array.ndim - 1))
    else:
        res = getattr(aggregate_b200, func)(codes, array, axis=array.ndim - 1)
    return cast(res)
