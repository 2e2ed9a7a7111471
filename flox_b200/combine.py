"""Combine stage: reduce per-chunk intermediates stacked on an axis (flox/dask.py:90-144).

`resolve(name)` gives the callable `_initialize_aggregation` binds into `agg.simple_combine`
(the reference binds np.sum / np.max / … there, aggregations.py:1018-1028); it has numpy's
reduction signature `f(array, axis=..., keepdims=...)` but runs the `floxb_combine` CUDA kernel.
In the fused GPU path the chunks of one GPU accumulate into the same bins, so this module is
only exercised for explicit per-chunk intermediates and for the cross-GPU merge.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L
from .arrays import from_device, is_torch, np_dtype_of, to_device
from .kernels import get_kernels

_KINDS = {"sum": L.COMB_SUM, "prod": L.COMB_PROD, "max": L.COMB_MAX, "min": L.COMB_MIN,
          "nanmax": L.COMB_NANMAX, "nanmin": L.COMB_NANMIN}


SIMPLE = tuple(_KINDS)


def _normalize_axes(axis, ndim):
    if axis is None:
        return tuple(range(ndim))
    if np.isscalar(axis):
        axis = (axis,)
    return tuple(sorted(a % ndim for a in axis))


def _stack_front(t: torch.Tensor, axes: tuple[int, ...]):
    keep = [d for d in range(t.ndim) if d not in axes]
    perm = list(axes) + keep
    moved = t.permute(perm).contiguous()
    nblk = 1
    for a in axes:
        nblk *= t.shape[a]
    kept_shape = [t.shape[d] for d in keep]
    return moved.reshape(nblk, -1), kept_shape, keep


def _restore(flat: torch.Tensor, t_shape, axes, kept_shape, keepdims: bool):
    out = flat.reshape(kept_shape)
    if keepdims:
        full = list(t_shape)
        for a in axes:
            full[a] = 1
        out = out.reshape(full)
    return out


def reduce_stacked(kind: int, array, axis, keepdims: bool = False):
    """Reduce `array` over `axis` with the floxb_combine kernel; numpy in -> numpy out."""
    np_dt = np_dtype_of(array)
    t = to_device(array)
    axes = _normalize_axes(axis, t.ndim)
    wide = torch.float64 if t.dtype.is_floating_point else torch.int64
    stacked, kept_shape, _ = _stack_front(t.to(wide), axes)
    out = torch.empty(stacked.shape[1], dtype=wide, device=t.device)
    if stacked.shape[0] == 0:
        raise ValueError("cannot combine zero blocks")
    get_kernels().combine(kind, stacked, out)
    res = _restore(out, t.shape, axes, kept_shape, keepdims).to(t.dtype)
    return from_device(res, array, np_dt)


def resolve(name: str):
    """Callable for a combine spec string ('sum', 'max', 'nanmin', ...)."""
    if name not in _KINDS:
        raise NotImplementedError(f"combine={name!r} has no simple (stack-and-reduce) form")
    kind = _KINDS[name]

    def combine(array, axis=None, keepdims=False):
        return reduce_stacked(kind, array, axis, keepdims)

    combine.__name__ = name
    return combine


def var_combine_stacked(m2, s, n, axis: int, keepdims: bool = True):
    """Chan merge of (M2, S, n) blocks stacked on `axis` (flox/aggregations.py:392-451)."""
    np_dt = np_dtype_of(m2)
    like = m2
    tm2, ts, tn = (to_device(x) for x in (m2, s, n))
    axes = _normalize_axes(axis, tm2.ndim)
    parts = []
    kept_shape = None
    for t in (tm2, ts, tn):
        st, kept_shape, _ = _stack_front(t.to(torch.float64), axes)
        parts.append(st)
    outs = [torch.empty(parts[0].shape[1], dtype=torch.float64, device=tm2.device) for _ in range(3)]
    get_kernels().var_combine(parts[0], parts[1], parts[2], *outs)
    # dtypes as the reference produces them: S and n keep the intermediate dtype, the merged M2 is float64 as soon as
    # more than one block is merged (the adjustment terms divide by an integer-augmented denominator,
    # aggregations.py:427-436: float32 / (float32 + int64) -> float64)
    nblk = parts[0].shape[0]
    dts = [(torch.float64, np.dtype("float64")) if (k == 0 and nblk > 1) else (tm2.dtype, np_dt) for k in range(3)]
    return tuple(from_device(_restore(o, tm2.shape, axes, kept_shape, keepdims).to(tdt), like, ndt)
                 for o, (tdt, ndt) in zip(outs, dts))
