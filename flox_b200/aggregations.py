"""Aggregation specs and engine dispatch (behavioural mirror of flox/aggregations.py).

Kept: `generic_aggregate` (:60-133, the engine plug-in seam), `Aggregation` (:161-301), the
built-in specs for the scatter-reduction family (:304-632), `var_chunk` / `_var_combine` /
`_var_finalize` / `_std_finalize` (:348-477) and `_initialize_aggregation` (:925-1030).
Not kept (out of scope): scans, first/last, mode, quantile/median, all/any.
"""
from __future__ import annotations

import copy
import logging
from functools import partial
from typing import Any, Callable

import numpy as np
import torch

from . import aggregate_b200, xrdtypes as dtypes
from .arrays import from_device, is_torch
from .multiarray import MultiArray
from .xrutils import is_scalar

logger = logging.getLogger("flox_b200")

ENGINES = ("b200",)


def _is_arg_reduction(func) -> bool:
    if isinstance(func, str):
        return func in ("argmin", "argmax", "nanargmax", "nanargmin")
    return isinstance(func, Aggregation) and func.reduction_type == "argreduce"


def generic_aggregate(group_idx, array, *, engine: str, func: str, axis=-1, size=None, fill_value=None, dtype=None,
                      **kwargs):
    """Look `func` up by NAME in the engine module and call it (flox/aggregations.py:60-133)."""
    if func == "identity":
        return array
    if engine not in ENGINES:
        raise ValueError(f"Expected engine to be one of {list(ENGINES)}. Received {engine} instead.")
    try:
        method = getattr(aggregate_b200, func)
    except AttributeError:
        raise NotImplementedError(f"Reduction {func!r} is not implemented by engine={engine!r}") from None
    return method(group_idx, array, axis=axis, size=size, fill_value=fill_value, dtype=dtype, **kwargs)


def _atleast_1d(inp, min_length: int = 1):
    if is_scalar(inp) or callable(inp):
        inp = (inp,) * min_length
    assert len(inp) >= min_length
    return tuple(inp)


class Dim:
    """A dimension a reduction adds to its result (aggregations.py:148-158): the quantile axis."""

    def __init__(self, name, values):
        self.name, self.values = name, values

    @property
    def is_scalar(self) -> bool:
        return is_scalar(self.values)

    @property
    def size(self) -> int:
        return 0 if self.is_scalar else len(self.values)


def quantile_new_dims_func(q) -> tuple:
    return (Dim(name="quantile", values=q),)


QUANTILE_FUNCS = ("quantile", "nanquantile", "median", "nanmedian")


class Aggregation:
    """Blueprint of a grouped reduction: per-chunk funcs, combine funcs, finalize, fills, dtypes.

    Same constructor and attributes as flox.aggregations.Aggregation (aggregations.py:161-301);
    custom instances (tests/test_core.py:1509-1548 style) work with engine="b200".
    """

    def __init__(self, name, *, numpy=None, chunk, combine, preprocess=None, finalize=None, fill_value=None,
                 final_fill_value=dtypes.NA, dtypes=None, final_dtype=None, reduction_type="reduce",
                 new_dims_func=None, preserves_dtype=False):
        self.name = name
        self.preprocess = preprocess
        self.reduction_type = reduction_type
        self.numpy = (numpy,) if numpy is not None else (name,)
        self.chunk = _atleast_1d(chunk)
        self.combine = _atleast_1d(combine)
        self.simple_combine: tuple = ()
        self.finalize = finalize
        self.fill_value: dict[str, Any] = {name: final_fill_value}
        self.fill_value["intermediate"] = self._per_chunk(fill_value, "fill_value")
        self.dtype_init = {"final": final_dtype, "intermediate": self._per_chunk(dtypes, "dtype")}
        self.dtype: dict[str, Any] = None  # set by _initialize_aggregation
        self.finalize_kwargs: dict = {}
        self.min_count = 0
        self.new_dims_func = new_dims_func or (lambda **kw: ())
        self.preserves_dtype = preserves_dtype

    def _per_chunk(self, value, what):
        value = value if isinstance(value, tuple) else (value,)
        if len(value) == 1 and len(self.chunk) > 1:
            value = value * len(self.chunk)
        if len(value) != len(self.chunk):
            raise ValueError(f"Bad {what} specified for Aggregation {self.name}.")
        return tuple(value)

    @property
    def new_dims(self):
        return self.new_dims_func(**self.finalize_kwargs)

    @property
    def num_new_vector_dims(self) -> int:
        return len(tuple(d for d in self.new_dims if not getattr(d, "is_scalar", True)))

    def __repr__(self):
        return (f"{self.name!r}, fill: {list(self.fill_value.values())!r}, dtype: {self.dtype}\n"
                f"chunk: {self.chunk!r}\ncombine: {self.combine!r}\nfinalize: {self.finalize!r}\nmin_count: {self.min_count!r}")


# ---- mean ------------------------------------------------------------------------------------

def _mean_finalize(sum_, count):
    """sum / count; 0/0 -> NaN silently (aggregations.py:323-325)."""
    if is_torch(sum_):
        return sum_ / count
    with np.errstate(invalid="ignore", divide="ignore"):
        return sum_ / count


# ---- variance --------------------------------------------------------------------------------

def var_chunk(group_idx, array, *, skipna: bool, engine: str, axis=-1, size=None, fill_value=None, dtype=None):
    """Per-chunk `(M2, S, n)` (aggregations.py:348-389): n = nanlen, S = (nan)sum,
    M2 = (nan)sum((x - S/n)^2); all three in the float intermediate dtype.  On the GPU this is
    two passes over the chunk (sum+count, then centred squares) instead of the reference's three."""
    if engine not in ENGINES:
        raise ValueError(f"Expected engine to be one of {list(ENGINES)}. Received {engine} instead.")
    from .reductions import var_triple

    acc, lead, size = aggregate_b200.make_accumulators(group_idx, array, size)
    out_dt = np.dtype(dtype) if dtype is not None else (acc.np_dtype if acc.np_dtype.kind == "f" else np.dtype("float64"))
    shape = lead + (size,)
    return MultiArray(tuple(from_device(t.reshape(shape), array, out_dt) for t in var_triple(acc, skipna, out_dt)))


def _var_combine(array: MultiArray, axis, keepdims=True):
    """Chan merge of stacked triples along `axis` (aggregations.py:392-451)."""
    from .combine import var_combine_stacked

    for ax in axis:
        if array.shape[ax] == 1:
            continue
        array = MultiArray(var_combine_stacked(*array.arrays, axis=ax, keepdims=keepdims))
    return array


def is_var_chunk_reduction(agg) -> bool:
    if isinstance(agg, partial):
        agg = agg.func
    return agg is blockwise_or_numpy_var or agg is var_chunk


def _var_finalize(multiarray, ddof=0):
    """M2 / (n - ddof); NaN where n - ddof < 0 (aggregations.py:460-468)."""
    m2, _, n = multiarray.arrays
    den = n - ddof
    if is_torch(m2):
        ret = m2 / den
        return torch.where(den < 0, torch.full_like(ret, float("nan")), ret)
    with np.errstate(invalid="ignore", divide="ignore"):
        ret = m2 / den
    ret = np.asarray(ret)
    ret[den < 0] = np.nan
    return ret


def _std_finalize(multiarray, ddof=0):
    v = _var_finalize(multiarray, ddof)
    return torch.sqrt(v) if is_torch(v) else np.sqrt(v)


def blockwise_or_numpy_var(*args, skipna: bool, ddof=0, std=False, **kwargs):
    """numpy-input / blockwise path: var_chunk then finalize at once (aggregations.py:475-477)."""
    res = _var_finalize(var_chunk(*args, skipna=skipna, **kwargs), ddof)
    if std:
        res = torch.sqrt(res) if is_torch(res) else np.sqrt(res)
    return res


def _pick_second(*x):
    return x[1]


# ---- the spec table (aggregations.py:304-632) ------------------------------------------------

def _build_reductions() -> dict[str, Aggregation]:
    A = Aggregation
    NA, INF, NINF = dtypes.NA, dtypes.INF, dtypes.NINF
    table = {
        "count": A("count", numpy="nanlen", chunk="nanlen", combine="sum", fill_value=0, final_fill_value=0,
                   dtypes=np.intp, final_dtype=np.intp),
        "sum": A("sum", chunk="sum", combine="sum", fill_value=0),
        "nansum": A("nansum", chunk="nansum", combine="sum", fill_value=0),
        "prod": A("prod", chunk="prod", combine="prod", fill_value=1, final_fill_value=1),
        "nanprod": A("nanprod", chunk="nanprod", combine="prod", fill_value=1),
        "mean": A("mean", chunk=("sum", "nanlen"), combine=("sum", "sum"), finalize=_mean_finalize, fill_value=(0, 0),
                  dtypes=(None, np.intp), final_dtype=np.floating),
        "nanmean": A("nanmean", chunk=("nansum", "nanlen"), combine=("sum", "sum"), finalize=_mean_finalize,
                     fill_value=(0, 0), dtypes=(None, np.intp), final_dtype=np.floating),
        "max": A("max", chunk="max", combine="max", fill_value=NINF, preserves_dtype=True),
        "nanmax": A("nanmax", chunk="nanmax", combine="nanmax", fill_value=NINF, final_fill_value=NA, preserves_dtype=True),
        "min": A("min", chunk="min", combine="min", fill_value=INF, preserves_dtype=True),
        "nanmin": A("nanmin", chunk="nanmin", combine="nanmin", fill_value=INF, final_fill_value=NA, preserves_dtype=True),
    }
    # first / last (aggregations.py:634-651): positional picks; the nan-skipping pair combines across chunks
    table["first"] = A("first", chunk=None, combine=None, fill_value=None, preserves_dtype=True)
    table["last"] = A("last", chunk=None, combine=None, fill_value=None, preserves_dtype=True)
    table["nanfirst"] = A("nanfirst", chunk="nanfirst", combine="nanfirst", fill_value=NA, preserves_dtype=True)
    table["nanlast"] = A("nanlast", chunk="nanlast", combine="nanlast", fill_value=NA, preserves_dtype=True)
    # sorted-segment family: only in-memory / blockwise (aggregations.py:672-712: "Support statistical quantities only
    # blockwise")
    table["median"] = A("median", chunk=None, combine=None, fill_value=NA, final_dtype=np.floating)
    table["nanmedian"] = A("nanmedian", chunk=None, combine=None, fill_value=NA, final_dtype=np.floating)
    table["quantile"] = A("quantile", chunk=None, combine=None, fill_value=NA, final_dtype=np.float64,
                          new_dims_func=quantile_new_dims_func)
    table["nanquantile"] = A("nanquantile", chunk=None, combine=None, fill_value=NA, final_dtype=np.float64,
                             new_dims_func=quantile_new_dims_func)
    for name, skipna, std in (("var", False, False), ("nanvar", True, False), ("std", False, True), ("nanstd", True, True)):
        table[name] = A(
            name, chunk=partial(var_chunk, skipna=skipna),
            numpy=partial(blockwise_or_numpy_var, skipna=skipna, **({"std": True} if std else {})),
            combine=(_var_combine,), finalize=_std_finalize if std else _var_finalize, fill_value=((0, 0, 0),),
            final_fill_value=np.nan, dtypes=(None,), final_dtype=np.floating)
    for name, val, comb, fill in (("argmax", "max", ("max", "argmax"), NINF), ("argmin", "min", ("min", "argmin"), INF),
                                  ("nanargmax", "nanmax", ("max", "argmax"), NINF),
                                  ("nanargmin", "nanmin", ("min", "argmin"), INF)):
        table[name] = A(name, chunk=(val, name), combine=comb, reduction_type="argreduce", fill_value=(fill, 0),
                        final_fill_value=-1, finalize=_pick_second, dtypes=(None, np.intp), final_dtype=np.intp)
    return table


REDUCTIONS: dict[str, Aggregation] = _build_reductions()
AGGREGATIONS = REDUCTIONS


def _initialize_aggregation(func, dtype, array_dtype, fill_value, min_count: int, finalize_kwargs) -> Aggregation:
    """Resolve dtypes / fills, append the min_count `nanlen`, bind combine functions
    (aggregations.py:925-1030)."""
    from . import combine as _combine

    if isinstance(func, Aggregation):
        agg = copy.deepcopy(func)
        func = agg.name
    elif isinstance(func, str):
        if func not in AGGREGATIONS:
            raise NotImplementedError(f"Reduction {func!r} not implemented yet")
        # prototypes hold only immutable tuples / callables besides these dicts: a shallow clone is enough and
        # saves the ~50-object deepcopy on every call
        proto = AGGREGATIONS[func]
        agg = copy.copy(proto)
        agg.fill_value = dict(proto.fill_value)
        agg.finalize_kwargs = dict(proto.finalize_kwargs) if getattr(proto, "finalize_kwargs", None) else {}
        agg.dtype = None
    else:
        raise ValueError("Bad type for func. Expected str or Aggregation")

    user_dtype = np.dtype(dtype) if dtype is not None and not isinstance(dtype, np.dtype) else dtype
    array_dtype = np.dtype(array_dtype)
    final = dtypes._normalize_dtype(user_dtype or agg.dtype_init["final"], array_dtype, agg.preserves_dtype, fill_value)
    wide = np.result_type(array_dtype, final)
    inter = tuple(
        # NB the reference passes the intermediate FILL VALUE in `preserves_dtype`'s slot
        # (aggregations.py:961-967); only its truthiness matters and we keep that behaviour.
        dtypes._normalize_dtype(None, wide, bool(ifv) if not dtypes.is_sentinel(ifv) else True) if idt is None else np.dtype(idt)
        for idt, ifv in zip(agg.dtype_init["intermediate"], agg.fill_value["intermediate"])
    )
    agg.dtype = {"user": dtype, "final": final, "numpy": (final,), "intermediate": inter}

    agg.fill_value["user"] = fill_value
    agg.fill_value["intermediate"] = tuple(
        tuple(dtypes._get_fill_value(dt, f) for f in fv) if isinstance(fv, tuple) else dtypes._get_fill_value(dt, fv)
        for dt, fv in zip(inter, agg.fill_value["intermediate"])
    )
    agg.fill_value[func] = dtypes._get_fill_value(final, agg.fill_value[func])
    agg.fill_value["numpy"] = (0,) if _is_arg_reduction(agg) else (agg.fill_value[func],)

    if finalize_kwargs is not None:
        assert isinstance(finalize_kwargs, dict)
        agg.finalize_kwargs = finalize_kwargs

    if agg.name in ("nanmin", "nanmax") and min_count == 0:  # aggregations.py:997-1003
        min_count = 1
        agg.fill_value["user"] = agg.fill_value["user"] or agg.fill_value[agg.name]

    if min_count > 0:  # aggregations.py:1005-1014
        agg.min_count = min_count
        agg.numpy += ("nanlen",)
        if agg.chunk != (None,):
            agg.chunk += ("nanlen",)
            agg.combine += ("sum",)
        agg.fill_value["intermediate"] += (0,)
        agg.fill_value["numpy"] += (0,)
        agg.dtype["intermediate"] += (np.dtype(np.intp),)
        agg.dtype["numpy"] += (np.dtype(np.intp),)
    else:
        agg.min_count = 0

    # arg-reductions have no stack-and-reduce form; they go through the (value, index) pair merge
    agg.simple_combine = tuple(
        (_combine.resolve(c) if c in _combine.SIMPLE else None) if isinstance(c, str) else c for c in agg.combine)
    return agg
