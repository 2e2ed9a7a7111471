"""Array-level API and per-chunk driver of the B200 engine.

Behavioural mirror of the hot-path half of flox/core.py:
  chunk_reduce (:214-394)  chunk_argreduce (:157-211)  _finalize_results (:410-475)
  _reduce_blockwise (:478-524)  _validate_reindex (:527-588)  groupby_reduce (:739-1222)
with engine="b200" as the only engine.  The differences are the ones the GPU design buys:

* chunk_reduce builds ONE set of device accumulators per call and every requested reduction is
  an epilogue over it, so `("sum", "nanlen")` (mean) or `("nanmax", "nanlen")` read the chunk once
  where the reference loops `generic_aggregate` per func (core.py:348-391);
* there is no `_prepare_for_flox` sort (core.py:336-339) and no NaN-sentinel column
  (core.py:381-383): the kernels are scatter reductions that skip codes outside [0, size);
* inputs may be numpy arrays, torch tensors (CPU or CUDA) or a `ChunkedArray` (the stand-in for a
  dask array: a list of slabs along the reduced axis, see chunked.py); results come back in the
  container kind of `array`.
"""
from __future__ import annotations

import logging
import math
from collections.abc import Sequence
from functools import partial
from typing import Any

import numpy as np
import pandas as pd
import torch

from . import xrdtypes
from .aggregate_b200 import make_accumulators
from .aggregations import (
    ENGINES, QUANTILE_FUNCS, Aggregation, _atleast_1d, _initialize_aggregation, _is_arg_reduction, blockwise_or_numpy_var,
    is_var_chunk_reduction, quantile_new_dims_func, var_chunk,
)
from .arrays import from_device, is_torch, np_dtype_of
from .factorize import _factorize_multiple, factorize_
from .multiarray import MultiArray
from .reductions import evaluate, var_triple
from .reindex import ReindexArrayType, ReindexStrategy, reindex_

logger = logging.getLogger("flox_b200")

DUMMY_AXIS = -2  # axis inserted per chunk and reduced by _simple_combine (core.py:97-100)
DEFAULT_ENGINE = "b200"


# ---------------------------------------------------------------------------------------------
# small container-agnostic helpers
# ---------------------------------------------------------------------------------------------

def _asarray(x):
    return x if (is_torch(x) or isinstance(x, np.ndarray)) else np.asarray(x)


def _transpose(x, order):
    return x.permute(*order) if is_torch(x) else x.transpose(order)


def _broadcast_to(x, shape):
    if tuple(x.shape) == tuple(shape):
        return x
    if not (is_torch(x) or isinstance(x, np.ndarray)):
        raise NotImplementedError("chunked labels must already have the broadcast shape of all `by`")
    return x.broadcast_to(shape) if is_torch(x) else np.broadcast_to(x, shape)


def _astype(x, dt):
    if isinstance(x, MultiArray):
        return x.astype(dt)
    if is_torch(x):
        from .arrays import torch_dtype

        return x.to(torch_dtype(dt))
    return x.astype(dt, copy=False)


def _is_nanlen(reduction) -> bool:
    return isinstance(reduction, str) and reduction == "nanlen"


def _is_first_last_reduction(func) -> bool:
    name = func.name if isinstance(func, Aggregation) else func
    return name in ("nanfirst", "nanlast", "first", "last")


def _is_minmax_reduction(func) -> bool:
    return not _is_arg_reduction(func) and isinstance(func, str) and ("max" in func or "min" in func)


def _issorted(arr) -> bool:
    return bool((arr[:-1] <= arr[1:]).all())


def _move_reduce_dims_to_end(arr, axis):
    axis = tuple(axis)
    order = tuple(ax for ax in range(arr.ndim) if ax not in axis) + axis
    return _transpose(arr, order)


def _collapse_axis(arr, naxis: int):
    return arr.reshape(tuple(arr.shape[:-naxis]) + (math.prod(arr.shape[-naxis:]),))


# ---------------------------------------------------------------------------------------------
# per-chunk driver
# ---------------------------------------------------------------------------------------------

def chunk_reduce(array, by, func, expected_groups, axis=None, fill_value=None, dtype=None, reindex: bool = False,
                 engine: str = DEFAULT_ENGINE, kwargs=None, sort: bool = True, user_dtype=None) -> dict:
    """Grouped reduction(s) of one in-memory chunk (reference: flox/core.py:214-394).

    Returns `{"groups": ..., "intermediates": [one array per func]}`, each intermediate shaped
    `array.shape[:-nax] + (1,)*(nax-1) + groups.shape`.
    """
    if engine not in ENGINES:
        raise ValueError(f"Expected engine to be one of {list(ENGINES)}. Received {engine} instead.")
    array, by = _asarray(array), _asarray(by)
    funcs = _atleast_1d(func)
    nfuncs = len(funcs)
    dtypes = _atleast_1d(dtype, nfuncs)
    fill_values = _atleast_1d(fill_value, nfuncs) if not isinstance(fill_value, tuple) else fill_value
    kwargss = ({},) * nfuncs if kwargs is None else kwargs

    if isinstance(axis, Sequence):
        axes = tuple(axis)
        nax = len(axes)
    else:
        nax = by.ndim
        axes = () if axis is None else (axis,) * nax
    assert by.ndim <= array.ndim
    final_array_shape = tuple(array.shape[: array.ndim - nax]) + (1,) * (nax - 1)
    final_groups_shape = (1,) * (nax - 1)
    if 1 < nax < by.ndim:
        by = _collapse_axis(by, nax)
        array = _collapse_axis(array, nax)
        axes, nax = (-1,), 1

    group_idx, grps, found_groups_shape, _, size, props = factorize_(
        (by,), axes, expected_groups=(expected_groups,), reindex=bool(reindex), sort=sort)
    (groups,) = grps
    trailing = tuple(array.shape[array.ndim - by.ndim:])
    if tuple(group_idx.shape) != trailing:  # size-1 label dims (core.py:298-312; no F-order trick needed)
        group_idx = _broadcast_to(group_idx, trailing)
    array2 = array.reshape(tuple(array.shape[: array.ndim - by.ndim]) + (math.prod(trailing),))
    group_idx = group_idx.reshape(-1)

    empty = props.nanmask is not None and bool(np.all(props.nanmask))
    results: dict[str, Any] = {"groups": [], "intermediates": []}
    if reindex and expected_groups is not None:
        results["groups"] = expected_groups
    else:
        results["groups"] = np.array([np.nan]) if empty else groups
    final_array_shape += tuple(results["groups"].shape)
    final_groups_shape += tuple(results["groups"].shape)
    out_shape = final_array_shape[:-1] + tuple(found_groups_shape)

    acc = None
    lead = tuple(array2.shape[:-1])
    array_np_dtype = np_dtype_of(array)
    previous = ""
    for reduction, fv, kw, dt in zip(funcs, fill_values, kwargss, dtypes):
        is_var = is_var_chunk_reduction(reduction)
        if empty and not is_var:
            shape = final_array_shape
            if is_torch(array):
                from .arrays import torch_dtype

                result = torch.full(shape, fv, dtype=torch_dtype(dt if dt is not None else array_np_dtype), device=array.device)
            else:
                result = np.full(shape, fv, dtype=dt if dt is not None else array_np_dtype)
        elif _is_nanlen(reduction) and _is_nanlen(previous):
            result = results["intermediates"][-1]
        else:
            is_quantile = isinstance(reduction, str) and reduction in QUANTILE_FUNCS
            if (is_var or not (callable(reduction) or is_quantile)) and acc is None:
                acc, lead, size = make_accumulators(group_idx, array2, size)
            if is_var:
                base, pkw = (reduction.func, reduction.keywords) if isinstance(reduction, partial) else (reduction, {})
                skipna = bool(pkw.get("skipna", False))
                out_dt = np.dtype(dt) if dt is not None else (array_np_dtype if array_np_dtype.kind == "f" else np.dtype("float64"))
                if base is var_chunk:
                    result = MultiArray(tuple(
                        from_device(t.reshape(lead + (size,)), array, out_dt).reshape(out_shape)
                        for t in var_triple(acc, skipna, out_dt)))
                else:  # blockwise_or_numpy_var: finalize at once
                    name = ("nan" if skipna else "") + ("std" if pkw.get("std", False) else "var")
                    ddof = kw.get("ddof", pkw.get("ddof", 0))
                    t = evaluate(acc, name, out_dtype=out_dt, fill_value=np.nan, ddof=ddof)
                    result = from_device(t.reshape(lead + (size,)), array, out_dt).reshape(out_shape)
            elif isinstance(reduction, str) and reduction in QUANTILE_FUNCS:
                # sorted-segment family: one sort of the chunk by (group, value), then per-group order statistics
                from . import aggregate_b200

                out_dt = np.dtype(dt) if dt is not None else (array_np_dtype if array_np_dtype.kind == "f" else np.dtype("float64"))
                result = getattr(aggregate_b200, reduction)(group_idx, array2, size=size, dtype=out_dt, fill_value=fv, **kw)
                if reduction in ("quantile", "nanquantile"):  # core.py:385-389
                    new_dims_shape = tuple(d.size for d in quantile_new_dims_func(**kw) if not d.is_scalar)
                else:
                    new_dims_shape = ()
                result = result.reshape(new_dims_shape + out_shape)
            elif callable(reduction):
                # custom Aggregation chunk function: it does its own grouped reduction (core.py:360-366)
                result = reduction(group_idx, array2, size=size, dtype=dt, fill_value=fv, **kw)
                result = result.reshape(out_shape)
            else:
                out_dt = np.dtype(dt) if dt is not None else array_np_dtype
                kw = dict(kw)
                t = evaluate(acc, reduction, out_dtype=out_dt, fill_value=fv, ddof=kw.pop("ddof", 0))
                result = from_device(t.reshape(lead + (size,)), array, out_dt).reshape(out_shape)
        results["intermediates"].append(result)
        previous = reduction
    if not (isinstance(results["groups"], pd.RangeIndex) and tuple(final_groups_shape) == (len(results["groups"]),)):
        # a RangeIndex stays one (O(1)); materialising, copying and comparing 1e6 group labels on the host cost more
        # per call than the reduction kernel of BASELINE config 5 (tools/c5_hostprofile.py)
        results["groups"] = np.broadcast_to(np.asarray(results["groups"]), final_groups_shape)
    return results


def chunk_argreduce(array_plus_idx, by, func, expected_groups, axis, fill_value, dtype=None, reindex: bool = False,
                    engine: str = DEFAULT_ENGINE, sort: bool = True, user_dtype=None) -> dict:
    """Per-chunk (extremum, global position) pair (reference: flox/core.py:157-211)."""
    array, idx = array_plus_idx
    by = _broadcast_to(_asarray(by), tuple(array.shape))
    results = chunk_reduce(array, by, func, expected_groups=None, axis=axis, fill_value=fill_value, dtype=dtype,
                           engine=engine, sort=sort, user_dtype=user_dtype)
    groups = np.asarray(results["groups"])
    if not np.all(pd.isnull(groups)):
        idx = np.broadcast_to(np.asarray(idx), tuple(array.shape))
        local = results["intermediates"][1]
        local = local.cpu().numpy() if is_torch(local) else np.asarray(local)
        results["intermediates"][1] = idx[np.unravel_index(local, tuple(array.shape))]
    if reindex and expected_groups is not None:
        results["intermediates"][1] = reindex_(results["intermediates"][1], groups.squeeze(), expected_groups, fill_value=0)
    assert tuple(results["intermediates"][0].shape) == tuple(results["intermediates"][1].shape)
    return results


# ---------------------------------------------------------------------------------------------
# finalize
# ---------------------------------------------------------------------------------------------

def _squeeze_results(results: dict, axis) -> dict:
    groups = results["groups"]
    if isinstance(groups, pd.RangeIndex):
        out = {"groups": groups}
    else:
        out = {"groups": np.squeeze(groups, axis=tuple(ax for ax in range(groups.ndim - 1) if groups.shape[ax] == 1))}
    squeezed = []
    for v in results["intermediates"]:
        for ax in sorted((a for a in sorted(axis)[:-1] if v.shape[a] == 1), reverse=True):
            v = v.squeeze(ax)
        squeezed.append(v)
    out["intermediates"] = squeezed
    return out


def _finalize_results(results: dict, agg: Aggregation, axis, expected_groups, reindex: ReindexStrategy) -> dict:
    """Squeeze, agg.finalize, min_count mask, optional reindex, cast (reference: core.py:410-475)."""
    squeezed = _squeeze_results(results, tuple(agg.num_new_vector_dims + ax for ax in axis))
    min_count = agg.min_count
    counts = None
    if min_count > 0:
        counts = squeezed["intermediates"][-1]
        squeezed["intermediates"] = squeezed["intermediates"][:-1]
    if agg.finalize is None:
        value = squeezed["intermediates"][0]
    else:
        value = agg.finalize(*squeezed["intermediates"], **agg.finalize_kwargs)

    fill_value = agg.fill_value["user"]
    if counts is not None:
        mask = counts < min_count
        if bool(mask.any()):
            if fill_value is None:
                raise ValueError("Filling is required but fill_value is None.")
            if fill_value == xrdtypes.NA:
                new_dtype, fill_value = xrdtypes.maybe_promote(np_dtype_of(value))
                value = _astype(value, new_dtype)
            if is_torch(value):
                value = torch.where(mask, torch.as_tensor(fill_value, dtype=value.dtype, device=value.device), value)
            else:
                value = np.where(mask, fill_value, value)

    finalized: dict[str, Any] = {}
    if not reindex.blockwise and expected_groups is not None:
        value = reindex_(value, squeezed["groups"], expected_groups, fill_value=fill_value, array_type=reindex.array_type)
        finalized["groups"] = expected_groups
    else:
        finalized["groups"] = squeezed["groups"]
    finalized[agg.name] = _astype(value, agg.dtype["final"])
    return finalized


def _reduce_blockwise(array, by, agg: Aggregation, *, axis, expected_groups, fill_value, engine: str, sort: bool,
                      reindex: ReindexStrategy) -> dict:
    """In-memory inputs and method="blockwise" blocks: straight to the final value
    (reference: core.py:478-524)."""
    agg.finalize = None
    kwargs_ = (agg.finalize_kwargs,) + ({},) + ({},)
    results = chunk_reduce(array, by, func=agg.numpy, axis=axis, expected_groups=expected_groups,
                           fill_value=agg.fill_value["numpy"], dtype=agg.dtype["numpy"], kwargs=kwargs_, engine=engine,
                           sort=sort, reindex=bool(reindex.blockwise), user_dtype=agg.dtype["user"])
    if _is_arg_reduction(agg):
        first = results["intermediates"][0]
        host = first.cpu().numpy() if is_torch(first) else np.asarray(first)
        local = np.unravel_index(host, tuple(array.shape))[-1]
        results["intermediates"][0] = torch.as_tensor(local, device=first.device) if is_torch(first) else local
    return _finalize_results(results, agg, axis, expected_groups, reindex=reindex)


# ---------------------------------------------------------------------------------------------
# argument validation (reference: core.py:527-682)
# ---------------------------------------------------------------------------------------------

def _validate_reindex(reindex, func, method, expected_groups, any_by_dask: bool, is_dask_array: bool, array_dtype):
    def first_or_last():
        return func in ("first", "last") or (_is_first_last_reduction(func) and np.dtype(array_dtype).kind != "f")

    all_eager = not is_dask_array and not any_by_dask
    if reindex is True and not all_eager:
        if _is_arg_reduction(func):
            raise NotImplementedError
        if method == "cohorts" or (method == "blockwise" and not any_by_dask):
            raise ValueError("reindex=True is not a valid choice for method='blockwise' or method='cohorts'.")
        if first_or_last():
            raise ValueError("reindex must be None or False when func is 'first' or 'last.")
    strategy = reindex if isinstance(reindex, ReindexStrategy) else ReindexStrategy(blockwise=reindex)
    if strategy.blockwise is None:
        if method is None:
            return ReindexStrategy(blockwise=None)
        if all_eager:
            return ReindexStrategy(blockwise=True)
        if first_or_last():
            return ReindexStrategy(blockwise=False)
        if method == "blockwise":
            strategy = ReindexStrategy(blockwise=any_by_dask)
        elif _is_arg_reduction(func) or method == "cohorts":
            strategy = ReindexStrategy(blockwise=False)
        elif method == "map-reduce":
            strategy = ReindexStrategy(blockwise=not (expected_groups is None and any_by_dask))
    return strategy


def _assert_by_is_aligned(shape, by) -> None:
    assert all(b.ndim == by[0].ndim for b in by[1:])
    for idx, b in enumerate(by):
        if not all(j in (i, 1) for i, j in zip(shape[-b.ndim:], b.shape)):
            raise ValueError(
                "`array` and `by` arrays must be 'aligned' so that such that by_ is broadcastable to "
                "array.shape[-by.ndim:] for every array `by_` in `by`. "
                f"Received array of shape {tuple(shape)} but array {idx} in `by` has shape {tuple(b.shape)}.")


def _convert_expected_groups_to_index(expected_groups, isbin, sort: bool):
    out = []
    for ex, isbin_ in zip(expected_groups, isbin):
        if isinstance(ex, pd.IntervalIndex) or (isinstance(ex, pd.Index) and not isbin_):
            out.append(ex.sort_values() if sort else ex)
        elif ex is not None:
            if isbin_:
                out.append(pd.IntervalIndex.from_breaks(ex))
            else:
                out.append(pd.Index(np.sort(ex) if sort else ex))
        else:
            out.append(None)
    return tuple(out)


def _validate_expected_groups(nby: int, expected_groups):
    if expected_groups is None:
        return (None,) * nby
    if nby == 1 and not isinstance(expected_groups, tuple):
        if isinstance(expected_groups, (pd.Index, np.ndarray)):
            return (expected_groups,)
        arr = np.asarray(expected_groups)
        if np.issubdtype(arr.dtype, np.integer):
            arr = arr.astype(np.int64)
        return (arr,)
    if nby > 1 and not isinstance(expected_groups, tuple):
        raise ValueError(
            "When grouping by multiple variables, expected_groups must be a tuple of either arrays or objects "
            f"convertible to an array (like lists). Received a {type(expected_groups).__name__} instead.")
    if len(expected_groups) != nby:
        raise ValueError(f"Must have same number of `expected_groups` (received {len(expected_groups)}) "
                         f" and variables to group by (received {nby}).")
    return expected_groups


_arange_seen: tuple = (None, False)


def _is_arange(ex: np.ndarray) -> bool:
    """0..n-1?  The O(n) check is remembered for the last array OBJECT seen (a map-reduce loop passes the same
    expected_groups every call; at 1e6 groups the check costs more than the kernels)."""
    global _arange_seen
    seen = _arange_seen  # one read: other threads replace the tuple as a whole
    if seen[0] is ex and not ex.flags.writeable:
        return seen[1]
    n = ex.size
    ok = bool(ex[0] == 0 and ex[-1] == n - 1 and (n < 3 or np.all(ex[1:] - ex[:-1] == 1)))
    _arange_seen = (ex, ok)
    return ok


def _choose_engine(by, agg: Aggregation) -> str:
    """The reference picks between CPU engines by sortedness (core.py:712-736); a scatter kernel
    does not care, so the answer here is always the GPU engine."""
    return DEFAULT_ENGINE


def _normalize_axis_tuple(axis, ndim: int):
    axis = (axis,) if np.isscalar(axis) else tuple(axis)
    out = tuple(a + ndim if a < 0 else a for a in axis)
    if any(a < 0 or a >= ndim for a in out) or len(set(out)) != len(out):
        raise ValueError(f"invalid axis {axis} for array of dimension {ndim}")
    return out


# ---------------------------------------------------------------------------------------------
# public API
# ---------------------------------------------------------------------------------------------

def groupby_reduce(array, *by, func, expected_groups=None, sort: bool = True, isbin=False, axis=None, fill_value=None,
                   dtype=None, min_count: int | None = None, method: str | None = None, engine: str | None = None,
                   reindex=None, finalize_kwargs: dict | None = None):
    """GroupBy reduction on the GPU; signature and semantics of flox.groupby_reduce
    (reference: flox/core.py:739-1222).

    `array` / `by`: numpy arrays, torch tensors, or `flox_b200.ChunkedArray` (chunked along the
    reduced axis: the dask-array stand-in; `method` = "map-reduce" | "blockwise" | "cohorts"
    applies to it).  `engine`: None or "b200".  Returns `(result, *groups)`.
    """
    from .chunked import ChunkedArray, is_chunked

    if engine is not None and engine not in ENGINES:
        raise ValueError(f"Expected engine to be one of {list(ENGINES)}. Received {engine} instead.")
    if func in ("mode", "nanmode", "all", "any"):
        raise NotImplementedError(f"Reduction {func!r} is outside the b200 engine's scope")
    if func in ("quantile", "nanquantile") and (finalize_kwargs is None or "q" not in finalize_kwargs):
        raise ValueError("Please pass `q` for quantile calculations.")  # core.py:869-871

    bys = tuple(b if (is_torch(b) or is_chunked(b)) else np.asarray(b) for b in by)
    nby = len(bys)
    # pandas widens datetime64[D] / [h] / [m] labels to [s]; discovered groups are handed back in the labels' own unit
    by_datetime_units = {i: b.dtype for i, b in enumerate(bys) if isinstance(b, np.ndarray) and b.dtype.kind == "M"}
    by_is_chunked = tuple(is_chunked(b) for b in bys)
    any_by_chunked = any(by_is_chunked)
    provided_expected = expected_groups is not None
    if method == "cohorts" and any_by_chunked and expected_groups is None:
        raise ValueError(f"method={method!r} needs expected_groups when grouping by chunked labels.")
    if not (is_torch(array) or is_chunked(array)):
        array = np.asarray(array)
    array_dtype = np_dtype_of(array)

    reindex = _validate_reindex(reindex, func, method, expected_groups, any_by_chunked, is_chunked(array), array_dtype)

    is_bool_array = array_dtype == np.dtype(bool)
    if is_bool_array:
        array = array.astype(np.int_) if not is_torch(array) and not is_chunked(array) else array.to(torch.int64)
        array_dtype = np.dtype(np.int_)

    isbins = _atleast_1d(isbin, nby)
    _assert_by_is_aligned(tuple(array.shape), bys)
    expected_groups = _validate_expected_groups(nby, expected_groups)
    # 0..G-1 given as an integer array is a RangeIndex in disguise: treat it as one (no sort, no hash, no
    # lookup table; labels pass straight through as codes) and hand the caller back an array at the end
    arange_like: dict[int, np.dtype] = {}
    if not any(isbins):
        eg = list(expected_groups)
        for i, ex in enumerate(eg):
            # only for INTEGER labels: float labels with fractional values must be dropped by the
            # searchsorted + isin lookup (factorize.py:84-94), not truncated into a neighbouring group
            if (isinstance(ex, np.ndarray) and ex.ndim == 1 and ex.dtype.kind in "iu" and ex.size > 0
                    and np_dtype_of(bys[i]).kind in "iu" and _is_arange(ex)):
                arange_like[i] = ex.dtype
                eg[i] = pd.RangeIndex(ex.size)
        expected_groups = tuple(eg)
    for idx, (expect, chunked_) in enumerate(zip(expected_groups, by_is_chunked)):
        if chunked_ and expect is None:
            raise ValueError(f"`expected_groups` for array {idx} in `by` cannot be None since it is chunked.")
    expected_groups = _convert_expected_groups_to_index(expected_groups, isbins, sort)

    # ---- early factorization: labels -> codes in a RangeIndex of prod(grp_shape) (core.py:935-953)
    if nby > 1:
        shape = np.broadcast_shapes(*[tuple(b.shape) for b in bys])
        bys = tuple(_broadcast_to(b, shape) for b in bys)
    if any_by_chunked:
        bys, final_groups, grp_shape = ChunkedArray.factorize_chunks(bys, expected_groups, sort=sort)
    else:
        bys, final_groups, grp_shape = _factorize_multiple(bys, expected_groups, sort=sort)
    expected_ = pd.RangeIndex(math.prod(grp_shape))
    (by_,) = bys

    if axis is None:
        axis_ = tuple(range(array.ndim - by_.ndim, array.ndim))
    else:
        axis_ = _normalize_axis_tuple(axis, array.ndim)
    nax = len(axis_)

    is_npdatetime = array_dtype.kind in "Mm"
    datetime_dtype = array_dtype
    if is_npdatetime and func != "count":
        array = array.view(np.int64)
        array_dtype = np.dtype(np.int64)
    elif is_npdatetime and isinstance(array, np.ndarray):
        # count of datetimes = number of non-NaT values: hand the kernels a float array that is NaN exactly where the
        # input is NaT.  timedelta64 is an np.integer subclass, which the reference's notnull() treats as "cannot
        # be missing" (xrutils.py:149-160): every timedelta, NaT included, counts — reproduced as is.
        array = np.where(np.isnat(array), np.nan, 0.0) if array.dtype.kind == "M" else np.zeros(array.shape)
        array_dtype = np.dtype(np.float64)

    assert nax <= by_.ndim
    if nax < by_.ndim:
        by_ = _move_reduce_dims_to_end(by_, tuple(-array.ndim + ax + by_.ndim for ax in axis_))
        array = _move_reduce_dims_to_end(array, axis_)
        if not is_chunked(array):
            array = array.contiguous() if is_torch(array) else np.ascontiguousarray(array)
        axis_ = tuple(range(array.ndim - nax, array.ndim))

    if min_count is None:  # core.py:1026-1032
        min_count_ = 1 if (nax < by_.ndim or (fill_value is not None and provided_expected)) else 0
    else:
        min_count_ = min_count
    if min_count_ > 0 and func in ("nansum", "nanprod") and fill_value is None:
        fill_value = np.nan

    agg = _initialize_aggregation(func, dtype, array_dtype, fill_value, min_count_, finalize_kwargs)
    engine = DEFAULT_ENGINE

    if not is_chunked(array):
        reindex.set_blockwise_for_numpy()
        results = _reduce_blockwise(array, by_, agg, axis=axis_, expected_groups=expected_, fill_value=fill_value,
                                    engine=engine, sort=sort, reindex=reindex)
        groups = (results["groups"],)
        result = results[agg.name]
    else:
        from .mapreduce import chunked_groupby_agg

        if nax != by_.ndim:
            raise NotImplementedError("Chunked arrays must be reduced along all dimensions of `by`.")
        if func in QUANTILE_FUNCS:
            raise NotImplementedError(f"{func!r} needs every group in one block: concatenate the chunks (the reference "
                                      "supports it for method='blockwise' only, aggregations.py:672-673)")
        method = method or "map-reduce"
        if method not in ("map-reduce", "blockwise", "cohorts"):
            raise ValueError(f"Unknown method {method!r}")
        if _is_arg_reduction(agg) and method == "blockwise" and array.nchunks > 1:
            raise NotImplementedError("arg-reductions are not supported with method='blockwise', use 'cohorts' instead.")
        reindex = _validate_reindex(reindex, func, method, expected_, any_by_chunked, True, array_dtype)
        if fill_value is None:
            fill_value = agg.fill_value[agg.name]
        result, groups = chunked_groupby_agg(array=array, by=by_, agg=agg, expected_groups=expected_, axis=axis_,
                                             fill_value=fill_value, method=method, reindex=reindex, engine=engine, sort=sort)

    # ---- final reindex: drops the -1 group, fills absent groups, restores grp_shape (core.py:1184-1203)
    (groups_,) = groups
    result = reindex_(result, from_=groups_, to=expected_, fill_value=fill_value, array_type=ReindexArrayType.AUTO)
    result = result.reshape(tuple(result.shape[:-1]) + tuple(grp_shape))
    groups = final_groups

    if is_bool_array and (_is_minmax_reduction(func) or _is_first_last_reduction(func)):
        result = result.to(torch.bool) if is_torch(result) else result.astype(bool)
    if is_npdatetime and func != "count":
        result = result.astype(datetime_dtype)
    groups = tuple(
        np.arange(len(g), dtype=arange_like[i]) if i in arange_like
        else (g.to_numpy() if isinstance(g, pd.Index) and not isinstance(g, pd.RangeIndex) else g)
        for i, g in enumerate(groups))
    if by_datetime_units and not provided_expected:
        groups = tuple(g.astype(by_datetime_units[i]) if i in by_datetime_units and isinstance(g, np.ndarray)
                       and g.dtype.kind == "M" and g.dtype != by_datetime_units[i] else g for i, g in enumerate(groups))
    return (result, *groups)
