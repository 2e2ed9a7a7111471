"""engine="b200": drop-in for the reference's engine modules.

Every function has the engine contract of flox/aggregate_flox.py:133-143 (identical in
aggregate_npg.py:58 and aggregate_numbagg.py:42-52):

    f(group_idx, array, *, axis=-1, size=None, fill_value=None, dtype=None, **kw) -> array[..., size]

`group_idx` is 1-D integer codes over the LAST axis of `array`; codes outside [0, size) are
ignored; groups without rows hold `fill_value`; the result dtype is `dtype` or `array.dtype`
(aggregate_flox.py:171-180).  Inputs may be numpy arrays (copied to the GPU) or torch tensors
(used in place when already on the GPU); the result comes back in the same container kind.
Unlike engine="flox" the codes need NOT be sorted: the CUDA kernels are scatter reductions, so
core.chunk_reduce skips `_prepare_for_flox` (flox/core.py:336-339) for this engine.

`generic_aggregate(..., engine="b200")` resolves these by name exactly as
flox/aggregations.py:85-90 does for engine="flox".
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .accumulate import GroupedAccumulators
from .arrays import from_device, is_torch, kernel_dtype, np_dtype_of, to_device, torch_dtype
from .kernels import default_device
from .reductions import evaluate

__all__ = [
    "sum", "nansum", "prod", "nanprod", "max", "nanmax", "min", "nanmin", "nanlen", "mean", "nanmean",
    "sum_of_squares", "nansum_of_squares", "argmax", "argmin", "nanargmax", "nanargmin", "var", "nanvar", "std", "nanstd",
    "quantile", "nanquantile", "median", "nanmedian", "cumsum", "nancumsum", "ffill", "first", "last", "nanfirst", "nanlast",
]


def _codes_to_device(group_idx, device):
    t = to_device(group_idx, device)
    if t.dtype not in (torch.int32, torch.int64):
        t = t.to(torch.int64)
    return t.reshape(-1)


def make_accumulators(group_idx, array, size=None) -> tuple[GroupedAccumulators, tuple[int, ...], int]:
    """Device accumulators for one `(…, N)` array + `(N,)` codes problem."""
    device = default_device()
    np_dt = np_dtype_of(array)
    # A view whose reduced (last) axis is OUTERMOST in memory — e.g. a (time, lat, lon) array handed
    # over as array.transpose(1, 2, 0), which is what moving the grouped dim to the end produces
    # (flox/core.py:1014-1018, xarray.py:361-373) — is consumed in place as values[1, n, B]; the
    # reference gathers / copies it instead.
    time_major = None
    if np.ndim(array) >= 2:
        moved = array.movedim(-1, 0) if is_torch(array) else np.moveaxis(array, -1, 0)
        contiguous = (lambda t: t.is_contiguous()) if is_torch(array) else (lambda t: t.flags.c_contiguous)
        if not contiguous(array) and contiguous(moved):
            time_major = moved
    vals = to_device(time_major if time_major is not None else array, device)
    kdt = kernel_dtype(vals.dtype)
    if vals.dtype != kdt:
        vals = vals.to(kdt)
    codes = _codes_to_device(group_idx, device)
    n_reduced = vals.shape[0] if time_major is not None else (vals.shape[-1] if vals.ndim else -1)
    if vals.ndim == 0 or n_reduced != codes.numel():
        raise ValueError("array.shape[-1] must equal len(group_idx)")
    if size is None:
        size = int(codes.max().item()) + 1 if codes.numel() else 0
    if time_major is not None:
        lead = tuple(vals.shape[1:])
        A, B = 1, math.prod(lead)
        acc = GroupedAccumulators([(vals.reshape(1, -1, B), codes)], 1, B, int(size), kdt, device)
        acc.np_dtype = np_dt
        return acc, lead, int(size)
    lead = tuple(vals.shape[:-1])
    A = math.prod(lead)
    acc = GroupedAccumulators([(vals.reshape(A, -1, 1), codes)], A, 1, int(size), kdt, device)
    acc.np_dtype = np_dt
    return acc, lead, int(size)


def _engine(func: str, default_int_to_float: bool = False, fixed_dtype=None):
    def kernel(group_idx, array, *, axis=-1, size=None, fill_value=None, dtype=None, **kwargs):
        if axis not in (-1, np.ndim(array) - 1):
            raise ValueError("the b200 engine reduces along the last axis (as chunk_reduce always asks)")
        acc, lead, size = make_accumulators(group_idx, array, size)
        if dtype is not None:
            out_dt = np.dtype(dtype)
        elif fixed_dtype is not None:
            out_dt = np.dtype(fixed_dtype)
        else:
            out_dt = acc.np_dtype
            if default_int_to_float and out_dt.kind in "iub":
                out_dt = np.dtype("float64")
            elif out_dt.kind == "b":
                out_dt = np.dtype("int64")
        ddof = kwargs.pop("ddof", 0)
        res = evaluate(acc, func, out_dtype=out_dt, fill_value=fill_value, ddof=ddof)
        return from_device(res.reshape(lead + (size,)), array, out_dt)

    kernel.__name__ = func
    kernel.__qualname__ = func
    kernel.__doc__ = f"Grouped `{func}` on the GPU; engine contract of flox/aggregate_flox.py:133-143."
    return kernel


sum = _engine("sum")
nansum = _engine("nansum")
prod = _engine("prod")
nanprod = _engine("nanprod")
max = _engine("max")
nanmax = _engine("nanmax")
min = _engine("min")
nanmin = _engine("nanmin")
sum_of_squares = _engine("sum_of_squares")
nansum_of_squares = _engine("nansum_of_squares")
nanlen = _engine("nanlen", fixed_dtype=np.int64)
mean = _engine("mean", default_int_to_float=True)
nanmean = _engine("nanmean", default_int_to_float=True)
var = _engine("var", default_int_to_float=True)
nanvar = _engine("nanvar", default_int_to_float=True)
std = _engine("std", default_int_to_float=True)
nanstd = _engine("nanstd", default_int_to_float=True)
argmax = _engine("argmax", fixed_dtype=np.int64)
argmin = _engine("argmin", fixed_dtype=np.int64)
nanargmax = _engine("nanargmax", fixed_dtype=np.int64)
nanargmin = _engine("nanargmin", fixed_dtype=np.int64)
first = _engine("first")
last = _engine("last")
nanfirst = _engine("nanfirst")
nanlast = _engine("nanlast")


# ---- sorted-segment family: quantile / median ---------------------------------------------------

def _sort_rows_by_group_then_value(vals: torch.Tensor, codes: torch.Tensor, size: int) -> torch.Tensor:
    """Every row of `vals[R, n]` ordered by (code, value), NaNs last within their group, rows without a group (code
    outside [0, size)) behind everything — the order the reference reaches by partitioning `code + 1j * value`
    (flox/aggregate_flox.py:100-113).  f32: ONE 64-bit radix sort of the composite key (code << 32 | order-preserving value
    bits); f64: a stable sort by value followed by a stable sort by code (the device radix sorts behind torch.sort)."""
    codes = torch.where((codes >= 0) & (codes < size), codes, torch.full_like(codes, size))
    if vals.dtype == torch.float32:
        bits = vals.contiguous().view(torch.int32)
        key = bits ^ ((bits >> 31) & 0x7FFFFFFF)                       # signed order == numeric order
        key = torch.where(torch.isnan(vals), torch.full_like(key, 0x7FFFFFFF), key)
        comp = (codes.to(torch.int64) << 32)[None, :] + (key.to(torch.int64) + (1 << 31))
        comp = torch.sort(comp, dim=-1).values
        key = ((comp & 0xFFFFFFFF) - (1 << 31)).to(torch.int32)
        return (key ^ ((key >> 31) & 0x7FFFFFFF)).view(torch.float32)
    by_value = torch.sort(vals, dim=-1, stable=True)                   # NaN sorts last
    order = torch.sort(codes[by_value.indices], dim=-1, stable=True).indices
    return torch.gather(by_value.values, -1, order)


def _quantile_engine(func: str, skipna: bool, fixed_q=None):
    def kernel(group_idx, array, *, q=fixed_q, axis=-1, size=None, fill_value=None, dtype=None, **kwargs):
        if q is None:
            raise ValueError(f"{func} needs q")
        if axis not in (-1, np.ndim(array) - 1):
            raise ValueError("the b200 engine reduces along the last axis (as chunk_reduce always asks)")
        from .kernels import get_kernels

        device = default_device()
        np_dt = np_dtype_of(array)
        if np_dt.kind in "Mm":
            raise NotImplementedError("quantiles of datetime / timedelta arrays are not implemented by the b200 engine")
        vals = to_device(array, device)
        if vals.dtype not in (torch.float32, torch.float64):
            vals = vals.to(torch.float64)                              # integers: exact below 2^53, like the reference's float64 view
        codes = _codes_to_device(group_idx, device).to(torch.int64)
        n = codes.numel()
        if vals.ndim == 0 or vals.shape[-1] != n:
            raise ValueError("array.shape[-1] must equal len(group_idx)")
        if size is None:
            size = int(codes.max().item()) + 1 if n else 0
        size = int(size)
        lead = tuple(vals.shape[:-1])
        R = math.prod(lead)
        qs = np.atleast_1d(np.asarray(q, dtype=np.float64))
        out_dt = np.dtype(dtype) if dtype is not None else (np_dt if np_dt.kind == "f" else np.dtype("float64"))
        kdt = torch.float32 if out_dt == np.dtype("float32") else torch.float64
        fill = float("nan") if fill_value is None else float(fill_value)
        out = torch.empty((len(qs), R, size), dtype=kdt, device=device)
        if n == 0 or R == 0 or size == 0:
            out.fill_(fill)
        else:
            k = get_kernels()
            srt = _sort_rows_by_group_then_value(vals.reshape(R, n), codes, size).contiguous()
            counts = torch.bincount(codes[(codes >= 0) & (codes < size)], minlength=size)
            starts = torch.zeros(size + 1, dtype=torch.int64, device=device)
            starts[1:] = torch.cumsum(counts, 0)
            nvalid = torch.empty((R, size), dtype=torch.int64, device=device)
            k.segment_valid(srt, starts, nvalid)
            before = torch.cumsum(nvalid, dim=1) - nvalid               # valid values of the row's earlier groups
            k.segment_quantile(srt, starts, nvalid, before.contiguous(), torch.as_tensor(qs, device=device), skipna, fill, out)
        res = out.reshape((len(qs),) + lead + (size,))
        return from_device(res, array, out_dt)

    kernel.__name__ = func
    kernel.__qualname__ = func
    kernel.__doc__ = (f"Grouped `{func}` on the GPU (flox/aggregate_flox.py:50-130,218-221): result `(nq,) + array.shape[:-1] + "
                      "(size,)`, bit-identical to the reference's.")
    return kernel


quantile = _quantile_engine("quantile", False)
nanquantile = _quantile_engine("nanquantile", True)
median = _quantile_engine("median", False, fixed_q=0.5)
nanmedian = _quantile_engine("nanmedian", True, fixed_q=0.5)


# ---- grouped scans: cumsum / nancumsum / ffill (flox/aggregate_flox.py:269-329) --------------------------------------------

def _scan_engine(func: str, kind: int, skipna: bool):
    def kernel(group_idx, array, *, axis=-1, dtype=None, **kwargs):
        if axis not in (-1, np.ndim(array) - 1):
            raise ValueError("the b200 engine scans along the last axis")
        from .kernels import get_kernels

        device = default_device()
        np_dt = np_dtype_of(array)
        if np_dt.kind in "Mm":
            raise NotImplementedError("scans of datetime / timedelta arrays are not implemented by the b200 engine")
        vals = to_device(array, device)
        kdt = kernel_dtype(vals.dtype)
        if kind == L.SCAN_CUMSUM and kdt == torch.int32:
            kdt = torch.int64                                          # np.nancumsum widens small integers to the platform int
        if vals.dtype != kdt:
            vals = vals.to(kdt)
        codes = _codes_to_device(group_idx, device).to(torch.int64)
        n = codes.numel()
        if vals.ndim == 0 or vals.shape[-1] != n:
            raise ValueError("array.shape[-1] must equal len(group_idx)")
        lead = tuple(vals.shape[:-1])
        R = math.prod(lead)
        out = torch.empty((R, n), dtype=kdt, device=device)
        if R * n:
            order = torch.sort(codes, stable=True)                     # labels only: the reference's _prepare_for_flox argsort
            get_kernels().segment_scan(kind, vals.reshape(R, n).contiguous(), order.indices.contiguous(), order.values.contiguous(),
                                       skipna, out)
        out_np = _TORCH2NP_SCAN[kdt] if dtype is None else np.dtype(dtype)
        return from_device(out.reshape(lead + (n,)), array, out_np)

    kernel.__name__ = func
    kernel.__qualname__ = func
    kernel.__doc__ = f"Grouped `{func}` along the last axis, in order of appearance (flox/aggregate_flox.py:269-329)."
    return kernel


from . import _lib as L  # noqa: E402

_TORCH2NP_SCAN = {torch.float32: np.dtype("float32"), torch.float64: np.dtype("float64"), torch.int32: np.dtype("int32"),
                  torch.int64: np.dtype("int64")}
cumsum = _scan_engine("cumsum", L.SCAN_CUMSUM, False)
nancumsum = _scan_engine("nancumsum", L.SCAN_CUMSUM, True)
ffill = _scan_engine("ffill", L.SCAN_FFILL, True)
