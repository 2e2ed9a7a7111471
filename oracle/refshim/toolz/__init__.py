"""Minimal stand-in for the five toolz names the reference imports.

TEST INFRASTRUCTURE ONLY (golden-vector generation in the build container).
"""
import itertools as _it
from functools import reduce as _reduce

__all__ = ["accumulate", "groupby", "compose", "partition_all", "last"]


def accumulate(binop, seq, initial=None):
    if initial is None:
        return _it.accumulate(seq, binop)
    return _it.accumulate(seq, binop, initial=initial)


def groupby(key, seq):
    out = {}
    for item in seq:
        out.setdefault(key(item), []).append(item)
    return out


def compose(*funcs):
    if not funcs:
        return lambda x: x

    def composed(*args, **kwargs):
        *rest, first = funcs
        val = first(*args, **kwargs)
        for f in reversed(rest):
            val = f(val)
        return val

    return composed


def partition_all(n, seq):
    it = iter(seq)
    while True:
        chunk = tuple(_it.islice(it, n))
        if not chunk:
            return
        yield chunk


def last(seq):
    return _reduce(lambda _, x: x, seq)
