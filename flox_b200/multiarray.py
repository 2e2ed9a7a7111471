"""`(M2, S, n)` container for the variance intermediates (mirror of flox/multiarray.py:9-97).

Holds same-shaped arrays (numpy or torch) and forwards the handful of array operations the
chunk / combine / finalize pipeline applies to intermediates.
"""
from __future__ import annotations

import numpy as np

_HANDLED = {}


class MultiArray:
    def __init__(self, arrays):
        arrays = tuple(arrays)
        first = tuple(arrays[0].shape)
        if any(tuple(a.shape) != first for a in arrays):
            raise AssertionError("Expect all arrays to have the same shape")
        self.arrays = arrays

    def _map(self, fn):
        return type(self)(tuple(fn(a) for a in self.arrays))

    def astype(self, dt, **kwargs):
        if hasattr(self.arrays[0], "numpy") and not isinstance(self.arrays[0], np.ndarray):
            from .arrays import torch_dtype

            return self._map(lambda a: a.to(torch_dtype(dt)))
        return self._map(lambda a: a.astype(dt, **kwargs))

    def reshape(self, shape, **kwargs):
        return self._map(lambda a: a.reshape(shape, **kwargs))

    def squeeze(self, axis=None):
        return self._map(lambda a: a.squeeze(axis) if axis is not None else a.squeeze())

    def __getitem__(self, key):
        return self._map(lambda a: a[key])

    def __setitem__(self, key, value):
        assert len(value) == len(self.arrays)
        for a, v in zip(self.arrays, value):
            a[key] = v

    def __array_function__(self, func, types, args, kwargs):
        if func not in _HANDLED:
            return NotImplemented
        return _HANDLED[func](*args, **kwargs)

    @property
    def dtype(self):
        return self.arrays[0].dtype

    @property
    def shape(self):
        return tuple(self.arrays[0].shape)

    @property
    def ndim(self):
        return self.arrays[0].ndim


def _implements(np_func):
    def deco(fn):
        _HANDLED[np_func] = fn
        return fn

    return deco


@_implements(np.expand_dims)
def _expand_dims(ma, axis):
    return ma._map(lambda a: np.expand_dims(a, axis))


@_implements(np.concatenate)
def _concatenate(mas, axis):
    width = len(mas[0].arrays)
    assert all(len(m.arrays) == width for m in mas)
    return MultiArray(tuple(np.concatenate([m.arrays[i] for m in mas], axis) for i in range(width)))


@_implements(np.transpose)
def _transpose(ma, axes):
    return ma._map(lambda a: np.transpose(a, axes))


@_implements(np.squeeze)
def _squeeze(ma, axis):
    return ma._map(lambda a: np.squeeze(a, axis))


@_implements(np.full)
def _full(shape, fill_values, *args, **kwargs):
    return MultiArray(tuple(np.full(shape, fv, *args, **kwargs) for fv in fill_values))
