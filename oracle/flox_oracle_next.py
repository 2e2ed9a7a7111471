"""CPU oracle for the NEXT rows of the hot-path table: the sorted-segment family of the
reference's engine="flox" — grouped quantile / median and the grouped scans.  TEST INFRASTRUCTURE ONLY.

No CUDA counterpart exists yet: `flox_b200.groupby_reduce` raises NotImplementedError for these reductions.
This module and its fixtures (`tests/golden/golden_next.*`, written by `tests/golden/make_golden_next.py` from
the UNMODIFIED reference) exist so that the kernels of the next round have a pinned checker from their first
line.  `tests/test_oracle_next_pinning.py` requires bit-exact agreement with the reference's outputs.

Reference lines are `file:line` relative to `/root/reference/`.
"""
from __future__ import annotations

import numpy as np

from .flox_oracle import isnull, notnull


def _segments(codes):
    """Start offsets and labels of the runs of a SORTED code vector (aggregate_flox.py:164-166)."""
    codes = np.asarray(codes)
    starts = np.flatnonzero(np.concatenate(([True], codes[1:] != codes[:-1])))
    return starts, codes[starts]


def _lerp(lo, hi, t, dtype):
    """numpy's linear interpolation between order statistics (aggregate_flox.py:26-47): lo + (hi - lo) * t, taken
    from the upper end (hi - (hi - lo) * (1 - t)) where t >= 0.5."""
    out = np.empty_like(lo, dtype=dtype)
    with np.errstate(invalid="ignore"):
        d = np.subtract(hi, lo)
    np.add(lo, d * t, out=out)
    np.subtract(hi, d * (1 - t), out=out, where=t >= 0.5)
    return out


def engine_quantile(codes, array, *, q, skipna: bool, size=None, fill_value=None, dtype=None):
    """`aggregate_flox.quantile / nanquantile / median / nanmedian` (aggregate_flox.py:50-130, 133-192, 218-221) on
    SORTED codes.  Per row and group: the valid (non-NaN) values in ascending order v[0..n-1], the linear-method
    quantile at virtual index q*(n-1).  The reference compacts the valid values of ALL groups of a row to the front
    of one partitioned array (NaNs sort behind everything), so its virtual index carries the running count of valid
    values of the earlier groups — in float64, which decides the last bits of the interpolation weight; the same
    quantity is formed here.  A group without valid values is NaN; with skipna=False a group holding any NaN is NaN.
    Groups that do not occur keep `fill_value`.  Output `(nq,) + lead + (size,)`; at the engine boundary a scalar q
    still yields nq = 1 (aggregate_flox.py:156-160, 177-181) — the caller squeezes it."""
    codes = np.asarray(codes)
    array = np.asarray(array)
    qs = np.atleast_1d(np.asarray(q, dtype=np.float64))
    if dtype is None:
        dtype = array.dtype
    lead = array.shape[:-1]
    if codes.size == 0:
        size = 0 if size is None else size
        shape = (len(qs),) + lead + (size,)
        return np.full(shape, fill_value, dtype=dtype)
    starts, labels = _segments(codes)
    if size is None:
        size = int(labels[-1]) + 1
    bounds = np.concatenate((starts, [codes.size]))
    flat = array.reshape((-1, array.shape[-1]))
    values = flat.view(np.int64) if array.dtype.kind in "Mm" else flat
    nseg = len(labels)
    res = np.empty((len(qs), flat.shape[0], nseg), dtype=dtype)
    for r in range(flat.shape[0]):
        row = values[r]
        valid_sorted, n_valid, has_nan = [], np.empty(nseg, dtype=np.int64), np.empty(nseg, dtype=bool)
        for s in range(nseg):
            seg = row[bounds[s]:bounds[s + 1]]
            ok = notnull(flat[r, bounds[s]:bounds[s + 1]])
            v = np.sort(seg[ok], kind="stable")
            valid_sorted.append(v)
            n_valid[s] = v.size
            has_nan[s] = v.size != seg.size
        packed = np.concatenate(valid_sorted) if nseg else row[:0]
        before = np.concatenate(([0], np.cumsum(n_valid)[:-1]))          # valid values of the earlier groups
        for iq, qq in enumerate(qs):
            virtual = qq * (n_valid - 1)                                      # float64
            virtual[1:] += before[1:]
            lo = np.floor(virtual).astype(np.int64)
            hi = np.ceil(virtual).astype(np.int64)
            safe = packed if packed.size else np.zeros(1, dtype=row.dtype)
            # an empty group reads a neighbour's slot (or wraps); it is masked below, like in the reference
            lov = safe.take(lo, mode="wrap")
            hiv = safe.take(hi, mode="wrap")
            out = _lerp(lov, hiv, virtual - lo, dtype)
            bad = n_valid == 0
            if not skipna:
                bad = bad | has_nan
            out[bad] = np.nan
            res[iq, r] = out
    res = res.reshape((len(qs),) + lead + (nseg,))
    full = np.full((len(qs),) + lead + (size,), fill_value, dtype=dtype)
    full[..., labels] = res
    return full


def _invert(perm):
    return np.argsort(perm, kind="stable")


def engine_cumsum(codes, array, *, skipna: bool):
    """`aggregate_flox.cumsum / nancumsum` (aggregate_flox.py:296-329) on UNSORTED codes: a running sum within each
    group in order of appearance, returned in the input's order.  The reference forms ONE nancumsum over the
    code-sorted array and subtracts the total that precedes each group, which fixes the rounding of every element;
    the same two steps are taken here.  skipna=False: from the first NaN of a group onwards the result is NaN."""
    codes = np.asarray(codes)
    array = np.asarray(array)
    perm = np.argsort(codes, kind="stable")
    sc, sa = codes[perm], array[..., perm]
    starts, _ = _segments(sc) if sc.size else (np.zeros(0, dtype=np.int64), None)
    acc = np.nancumsum(sa, axis=-1)
    lengths = np.diff(np.concatenate((starts, [sc.size])))
    if len(starts) > 1:
        carried = acc[..., starts[1:] - 1]
        acc[..., starts[1]:] -= np.repeat(carried, lengths[1:], axis=-1)
    if not skipna and sa.dtype.kind in "cfO":
        seen = np.cumsum(isnull(sa), axis=-1, dtype=np.uint64)
        if len(starts) > 1:
            carried = seen[..., starts[1:] - 1]
            seen[..., starts[1]:] -= np.repeat(carried, lengths[1:], axis=-1)
        acc[seen > 0] = np.nan
    return acc[..., _invert(perm)]


def engine_ffill(codes, array):
    """`aggregate_flox.ffill` (aggregate_flox.py:269-293): every NaN takes the last valid value that precedes it
    WITHIN ITS GROUP (order of appearance); a group's leading NaNs stay NaN."""
    codes = np.asarray(codes)
    array = np.asarray(array)
    perm = np.argsort(codes, kind="stable")
    sc, sa = codes[perm], array[..., perm]
    n = sa.shape[-1]
    starts, _ = _segments(sc) if sc.size else (np.zeros(0, dtype=np.int64), None)
    hole = isnull(sa).copy()
    hole[..., starts] = False                       # a group's first slot is its own source even when it is NaN
    src = np.where(hole, 0, np.arange(n))
    np.maximum.accumulate(src, axis=-1, out=src)
    filled = np.take_along_axis(sa, src, axis=-1)
    return filled[..., _invert(perm)]
