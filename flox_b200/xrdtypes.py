"""dtype and fill-value rules of the hot path (behavioural mirror of flox/xrdtypes.py:9-209).

Only what `_initialize_aggregation`, `reindex_` and `_finalize_results` consult is kept:
sentinels NA / INF / NINF, `maybe_promote`, `_normalize_dtype`, `_maybe_promote_int`,
`_get_fill_value`.
"""
from __future__ import annotations

import numpy as np


class _Sentinel:
    __slots__ = ("_label",)

    def __init__(self, label: str):
        self._label = label

    def __repr__(self) -> str:
        return self._label

    def __eq__(self, other) -> bool:  # identity may be lost through pickling
        return isinstance(other, _Sentinel) and other._label == self._label

    def __hash__(self) -> int:
        return hash(self._label)


NA = _Sentinel("<NA>")      # dtype-appropriate missing marker          (xrdtypes.py:9)
INF = _Sentinel("<+INF>")   # largest representable value of the dtype  (xrdtypes.py:31)
NINF = _Sentinel("<-INF>")  # smallest representable value of the dtype (xrdtypes.py:32)
_SENTINELS = (NA, INF, NINF)


def is_sentinel(v) -> bool:
    return isinstance(v, _Sentinel)


def maybe_promote(dtype):
    """(dtype that can hold a missing value, that missing value)   — xrdtypes.py:35-66."""
    dtype = np.dtype(dtype)
    kind = dtype.kind
    if kind == "f":
        return dtype, np.nan
    if kind == "m":
        return dtype, np.timedelta64("NaT")
    if kind in "iu":
        return np.dtype("float32" if dtype.itemsize <= 2 else "float64"), np.nan
    if kind == "c":
        return dtype, complex(np.nan, np.nan)
    if kind == "M":
        return dtype, np.datetime64("NaT")
    return np.dtype(object), np.nan


def _maybe_promote_int(dtype) -> np.dtype:
    """Narrow ints widen to the platform int like np.sum/np.prod do — xrdtypes.py:175-185."""
    dtype = np.dtype(dtype)
    if dtype.kind == "i":
        return np.result_type(dtype, np.int_)
    if dtype.kind == "u":
        return np.result_type(dtype, np.uint)
    return dtype


def _normalize_dtype(dtype, array_dtype, preserves_dtype: bool, fill_value=None) -> np.dtype:
    """Result dtype of a reduction — xrdtypes.py:153-172."""
    array_dtype = np.dtype(array_dtype)
    if dtype is None:
        dtype = array_dtype if preserves_dtype else _maybe_promote_int(array_dtype)
    if dtype is np.floating:
        dtype = array_dtype if array_dtype.kind in "fcmM" else np.dtype("float64")
    dtype = np.dtype(dtype)
    if fill_value is not None and not is_sentinel(fill_value):
        dtype = np.result_type(dtype, fill_value)
    return dtype


def _extreme(dtype: np.dtype, positive: bool):
    if dtype.kind in "iu":
        info = np.iinfo(dtype)
        return info.max if positive else info.min
    if dtype.kind == "c":
        return complex(np.inf, np.inf) if positive else complex(-np.inf, -np.inf)
    if dtype.kind in "mM" and not positive:
        return dtype.type(np.iinfo(np.int64).min + 1, np.datetime_data(dtype)[0])
    if dtype.kind in "mM":
        return np.iinfo(np.int64).max
    return np.inf if positive else -np.inf


def _get_fill_value(dtype, fill_value):
    """Concrete fill for `dtype`; None means +INF — xrdtypes.py:188-209."""
    dtype = np.dtype(dtype)
    if dtype.kind in "US" and (fill_value is None or fill_value == NA):
        return ""
    if fill_value is None or fill_value == INF:
        return _extreme(dtype, True)
    if fill_value == NINF:
        return _extreme(dtype, False)
    if fill_value == NA:
        if dtype.kind in "fc":
            return np.nan
        if dtype.kind == "m":
            return np.timedelta64("NaT")
        if dtype.kind == "M":
            return np.datetime64("NaT")
        if dtype.kind in "iu":
            return _extreme(dtype, False)
        return None
    return fill_value
