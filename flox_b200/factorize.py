"""Label encoding: `by` -> integer codes in [0, G) with -1 for "no group".

Behavioural mirror of flox/factorize.py (`_factorize_single` :42, `_ravel_factorized` :102,
`offset_labels` :24, `factorize_` :147, `_factorize_multiple` :221) with two differences that the
CUDA kernels make possible:

* the kernels skip any code outside [0, G), so the reference's "-1 -> sentinel bin G, size += 1,
  strip the last column afterwards" (factorize.py:201-210, core.py:381-383) is not needed: codes
  keep their -1 and `size` is always the number of real groups;
* integer labels that are already codes (RangeIndex expected groups, factorize.py:44-53) and
  binning (np.digitize, factorize.py:55-82) stay on the GPU when `by` is a CUDA tensor.

Hash-based discovery of unknown / non-numeric labels stays on the host in pandas, as SURVEY §2
scopes it.
"""
from __future__ import annotations

import math
from collections import namedtuple
from functools import reduce

import numpy as np
import pandas as pd
import torch

from .arrays import is_torch, to_device
from .kernels import get_kernels
from .xrutils import isnull

FactorProps = namedtuple("FactorProps", "offset_group nan_sentinel nanmask")


def offset_labels(labels, ngroups: int):
    """row * ngroups + label over the last axis, -1 preserved (factorize.py:24-39)."""
    assert labels.ndim > 1
    nrows = math.prod(labels.shape[:-1])
    if is_torch(labels):
        rows = torch.arange(nrows, device=labels.device, dtype=labels.dtype).reshape(*labels.shape[:-1], 1)
        out = torch.where(labels < 0, labels.new_full((), -1), labels + rows * ngroups)
    else:
        rows = np.arange(nrows).reshape(labels.shape[:-1] + (1,))
        out = np.where(labels < 0, -1, labels + rows * ngroups)
    return out, nrows * ngroups


def _digitize_device(flat: torch.Tensor, edges: np.ndarray, right: bool) -> torch.Tensor:
    src = flat
    if src.dtype not in (torch.float32, torch.float64, torch.int32, torch.int64):
        src = src.to(torch.float64 if src.dtype.is_floating_point else torch.int64)
    src = src.contiguous()
    codes = torch.empty(src.numel(), dtype=torch.int64, device=src.device)
    e = torch.as_tensor(np.asarray(edges, dtype=np.float64), device=src.device)
    get_kernels().digitize(src, e, right, codes)
    return codes


def _is_int_index(expect) -> bool:
    return (isinstance(expect, pd.Index) and not isinstance(expect, (pd.RangeIndex, pd.IntervalIndex))
            and expect.dtype.kind in "iu" and len(expect) > 0 and expect.is_unique)


class IntLabelEncoder:
    """Integer labels against KNOWN integer groups, without leaving the device:

    * groups == 0..G-1 -> the labels ARE the codes ("passthrough"; kernels drop anything outside);
    * otherwise        -> a lookup table over [min(groups), max(groups)] ("lut").
    The reference does a host searchsorted + isin here (factorize.py:84-94).  Built once per
    `groupby_reduce` call and reused for every chunk.
    """

    def __init__(self, expect: pd.Index, sort: bool):
        ex = np.asarray(expect, dtype=np.int64)
        order = np.argsort(ex, kind="stable")
        already_sorted = bool(np.all(order == np.arange(len(ex))))
        self.groups = expect[order] if (sort and not already_sorted) else expect
        self.lo, self.hi = int(ex.min()), int(ex.max())
        self.mode = None
        self._table = None
        self._lut: dict = {}
        if self.lo == 0 and self.hi == len(ex) - 1 and (sort or already_sorted):
            self.mode = "passthrough"
        elif self.hi - self.lo < (1 << 26):
            self.mode = "lut"
            table = np.full(self.hi - self.lo + 1, -1, dtype=np.int64)
            table[(ex[order] if sort else ex) - self.lo] = np.arange(len(ex))
            self._table = table

    def __call__(self, flat: torch.Tensor) -> torch.Tensor:
        if self.mode == "passthrough":
            return flat if flat.dtype in (torch.int32, torch.int64) else flat.to(torch.int64)
        lut = self._lut.get(flat.device)
        if lut is None:
            lut = self._lut[flat.device] = torch.as_tensor(self._table, device=flat.device)
        src = flat if flat.dtype in (torch.int32, torch.int64) else flat.to(torch.int64)
        src = src.contiguous()
        codes = torch.empty(src.numel(), dtype=torch.int64, device=src.device)
        get_kernels().lut_codes(src, self.lo, lut, codes)
        return codes.reshape(flat.shape)


_last_encoder: tuple = (None, None, None)


def int_label_encoder(expect: pd.Index, sort: bool) -> IntLabelEncoder:
    global _last_encoder
    cached = _last_encoder  # ONE read: other threads replace the tuple as a whole (dask's threaded scheduler)
    if cached[0] is expect and cached[1] == sort:
        return cached[2]
    enc = IntLabelEncoder(expect, sort)
    _last_encoder = (expect, sort, enc)
    return enc


_DISCOVER_MAX_SPAN = 1 << 26  # presence table of at most 64 MiB


def _discover_int_labels(by):
    """Unknown integer (or datetime) labels: sorted unique values and codes WITHOUT the host hash pass of
    pd.factorize (flox/factorize.py:95-99): min/max -> presence table over [min, max] -> groups = min + nonzero(table)
    -> codes through a lookup table, all on the device (floxb_label_minmax / _presence / _lut_codes).  Returns None
    when it does not apply (non-integer labels, empty input, value span above 2^26)."""
    np_dt = None
    if is_torch(by):
        if by.dtype not in (torch.int8, torch.int16, torch.int32, torch.int64, torch.uint8) or by.numel() == 0:
            return None
        dev = by.reshape(-1)
    else:
        a = np.asarray(by)
        if a.dtype.kind not in "iuM" or a.size == 0 or (a.dtype.kind == "u" and a.dtype.itemsize == 8):
            return None
        np_dt = a.dtype
        dev = to_device(a.reshape(-1))   # datetime64 -> int64 view, small unsigned -> int64
    if dev.dtype not in (torch.int32, torch.int64):
        dev = dev.to(torch.int64)
    dev = dev.contiguous()
    kern = get_kernels()
    lo, hi = kern.label_minmax(dev)
    if np_dt is not None and np_dt.kind == "M" and lo == np.iinfo(np.int64).min:
        return None  # NaT labels: leave the null handling to the host path
    span = hi - lo + 1
    if span > _DISCOVER_MAX_SPAN:
        return None
    present = kern.label_presence(dev, lo, span).cpu().numpy()
    rel = np.flatnonzero(present)
    lut_np = np.full(span, -1, dtype=np.int64)
    lut_np[rel] = np.arange(rel.size)
    codes = torch.empty(dev.numel(), dtype=torch.int64, device=dev.device)
    kern.lut_codes(dev, lo, torch.as_tensor(lut_np, device=dev.device), codes)
    values = rel.astype(np.int64) + lo
    if np_dt is not None:
        values = values.view(np_dt) if np_dt.kind == "M" else values.astype(np_dt)
    elif is_torch(by) and by.dtype != torch.int64:
        values = values.astype(np.dtype(str(by.dtype).replace("torch.", "")))
    return pd.Index(values), codes


def _factorize_single(by, expect, *, sort: bool, reindex: bool):
    """(groups Index, codes shaped like `by`) for one grouper (factorize.py:42-99)."""
    shape = tuple(by.shape)
    on_device = is_torch(by)
    flat = by.reshape(-1)
    if expect is None and sort:
        found = _discover_int_labels(by)
        if found is not None:
            groups, codes = found
            # 1-D host labels: the codes were produced on the device and the reduction kernels want them there — hand
            # them on as a device tensor instead of copying 8 bytes per row back to the host and up again (BASELINE
            # config 1's call: 2 x 8 MB over PCIe for a 4 MB array)
            return groups, (codes if (on_device or by.ndim == 1) else codes.cpu().numpy()).reshape(shape)
    if on_device and not flat.dtype.is_floating_point and flat.dtype != torch.bool and _is_int_index(expect):
        enc = int_label_encoder(expect, sort)
        if enc.mode is not None:
            return enc.groups, enc(flat).reshape(shape)
    if isinstance(expect, pd.RangeIndex):
        # labels ARE codes; anything above the range is "no group".  The reference copies and
        # overwrites (factorize.py:44-53); the kernels' range check makes both unnecessary.
        groups = expect
        if on_device:
            codes = flat if flat.dtype in (torch.int32, torch.int64) else flat.to(torch.int64)
        else:
            codes = np.asarray(flat)
            if codes.dtype.kind not in "iu":
                bad = isnull(codes)
                codes = np.where(bad, -1, codes).astype(np.int64)
            top = expect[-1] if len(expect) else -1
            if len(codes) and (codes.max() > top):
                codes = np.where(codes > top, -1, codes)
    elif isinstance(expect, pd.IntervalIndex):
        if expect.closed == "both":
            raise NotImplementedError
        edges = np.concatenate([expect.left.to_numpy(), expect.right.to_numpy()[[-1]]])
        right = expect.closed_right
        if edges.dtype.kind == "M":
            edges = edges.view(np.int64)
            flat = flat.view(np.int64) if not on_device else flat
        groups = expect
        if len(edges) > 1:
            dev = to_device(flat) if not on_device else flat
            codes = _digitize_device(dev, edges, right)
            if not on_device:
                codes = codes.cpu().numpy()
        else:
            codes = (torch.full((flat.numel(),), -1, dtype=torch.int64, device=flat.device) if on_device
                     else np.full(flat.shape, -1, dtype=np.int64))
    else:
        host = flat.cpu().numpy() if on_device else np.asarray(flat)
        if expect is not None and reindex:
            expect_arr = np.asarray(expect)
            sorter = np.argsort(expect_arr, kind="stable")
            groups = expect[sorter] if sort else expect
            pos = np.searchsorted(expect_arr, host, sorter=sorter)
            missing = ~np.isin(host, expect_arr) | isnull(host) | (pos == len(expect_arr))
            if not sort:
                pos[pos == len(expect_arr)] = -1
                pos = sorter[pos]
            codes = pos.astype(np.int64)
            codes[missing] = -1
        else:
            codes, groups = pd.factorize(host, sort=sort)
            codes = codes.astype(np.int64)
        groups = pd.Index(groups) if not isinstance(groups, pd.Index) else groups
        if on_device:
            codes = torch.as_tensor(codes, device=by.device)
    return groups, codes.reshape(shape)


def _ravel_factorized(*factorized, grp_shape):
    """One code from several (factorize.py:102-108); -1 if any component is -1."""
    if len(factorized) == 1:
        return factorized[0]
    if all(is_torch(f) for f in factorized):
        out = factorized[0].reshape(-1).to(torch.int64).contiguous()
        size = grp_shape[0]
        for f, s in zip(factorized[1:], grp_shape[1:]):
            nxt = torch.empty_like(out)
            get_kernels().ravel_codes(out, f.reshape(-1).to(torch.int64).contiguous(), size, s, nxt)
            out, size = nxt, size * s
        return out.reshape(factorized[0].shape)
    host = [f.cpu().numpy() if is_torch(f) else np.asarray(f) for f in factorized]
    bad = reduce(np.logical_or, [h < 0 for h in host])
    combined = np.ravel_multi_index(tuple(np.where(bad, 0, h) for h in host), grp_shape, mode="wrap")
    combined[bad] = -1
    return combined


def factorize_(by, axes, *, expected_groups=None, reindex: bool = False, sort: bool = True, fastpath: bool = False):
    """Codes + groups for one or more groupers (factorize.py:147-213).

    Returns `(group_idx, found_groups, grp_shape, ngroups, size, props)`.  Unlike the reference
    `group_idx` keeps -1 for missing labels and `size == ngroups` (times the number of offset rows);
    `props.nanmask` is None for device codes (no host sync) and a bool array for host codes.
    """
    if expected_groups is None:
        expected_groups = (None,) * len(by)
    results = [_factorize_single(b, e, sort=sort, reindex=reindex) for b, e in zip(by, expected_groups)]
    found_groups = tuple(r[0] for r in results)
    grp_shape = tuple(len(g) for g in found_groups)
    ngroups = math.prod(grp_shape)
    group_idx = _ravel_factorized(*[r[1] for r in results], grp_shape=grp_shape)
    if fastpath:
        return group_idx, found_groups, grp_shape, ngroups, ngroups, None
    if len(axes) == 1 and by[0].ndim > 1:
        group_idx, size = offset_labels(group_idx.reshape(tuple(by[0].shape)), ngroups)
        offset_group = True
    else:
        size, offset_group = ngroups, False
    nanmask = None if is_torch(group_idx) else (np.asarray(group_idx) < 0)
    return group_idx, found_groups, grp_shape, ngroups, size, FactorProps(offset_group, -1, nanmask)


def _factorize_multiple(by, expected_groups, any_by_dask: bool = False, sort: bool = True):
    """Early, whole-array factorization (factorize.py:221-273): `by` tuple -> one code array."""
    group_idx, found_groups, grp_shape, *_ = factorize_(by, axes=(), expected_groups=expected_groups, reindex=True,
                                                        sort=sort, fastpath=True)
    return (group_idx,), found_groups, grp_shape
