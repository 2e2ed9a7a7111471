"""`ChunkedArray`: the stand-in for a dask array on the map-reduce path.

flox hands `dask_groupby_agg` (flox/dask.py:325) a dask array whose blocks along the reduced axis
become one `chunk_reduce` task each.  dask is not a dependency here; a ChunkedArray is simply the
list of those blocks — slabs `[..., n_k]` along the LAST axis (numpy arrays, or torch tensors that
may already live on the GPU) — plus the matching slabs of the labels.  `groupby_reduce` accepts it
wherever the reference accepts a dask array.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import pandas as pd
import torch

from .arrays import is_torch, np_dtype_of


class ChunkedArray:
    def __init__(self, chunks: Sequence):
        chunks = list(chunks)
        if not chunks:
            raise ValueError("ChunkedArray needs at least one chunk")
        lead = tuple(chunks[0].shape[:-1])
        for c in chunks:
            if tuple(c.shape[:-1]) != lead:
                raise ValueError("chunks must agree on all but the last axis")
        self.chunks = chunks
        self._sizes = tuple(int(c.shape[-1]) for c in chunks)
        # For a ChunkedArray of LABELS: set True to promise that the label values stay as they are, on every rank, for as
        # long as this object is used — label-derived facts (which groups straddle ranks in the group-sharded merge) are
        # then computed once instead of on every call.
        self.static_labels = False

    # -- construction ------------------------------------------------------------------------
    @classmethod
    def from_array(cls, array, chunksize: int) -> "ChunkedArray":
        n = array.shape[-1]
        return cls([array[..., lo:min(lo + chunksize, n)] for lo in range(0, max(n, 1), chunksize)])

    # -- array-ish surface -------------------------------------------------------------------
    @property
    def nchunks(self) -> int:
        return len(self.chunks)

    @property
    def chunk_sizes(self) -> tuple[int, ...]:
        return self._sizes

    @property
    def shape(self) -> tuple[int, ...]:
        return tuple(self.chunks[0].shape[:-1]) + (sum(self.chunk_sizes),)

    @property
    def ndim(self) -> int:
        return self.chunks[0].ndim

    @property
    def dtype(self):
        return np_dtype_of(self.chunks[0])

    def map(self, fn) -> "ChunkedArray":
        return ChunkedArray([fn(c) for c in self.chunks])

    def astype(self, dt):
        return self.map(lambda c: c.astype(dt))

    def to(self, dt):
        return self.map(lambda c: c.to(dt) if is_torch(c) else c.astype(str(dt).replace("torch.", "")))

    def view(self, dt):
        return self.map(lambda c: c.view(dt))

    def rechunk_like(self, other: "ChunkedArray") -> "ChunkedArray":
        if self.chunk_sizes == other.chunk_sizes:
            return self
        whole = self.concatenate()
        out, lo = [], 0
        for n in other.chunk_sizes:
            out.append(whole[..., lo:lo + n])
            lo += n
        return ChunkedArray(out)

    def concatenate(self):
        if is_torch(self.chunks[0]):
            return torch.cat(list(self.chunks), dim=-1)
        return np.concatenate(self.chunks, axis=-1)

    # -- labels ------------------------------------------------------------------------------
    @staticmethod
    def factorize_chunks(bys, expected_groups, sort: bool = True):
        """Per-chunk factorization against KNOWN expected groups (flox/factorize.py:241-268:
        the `any_by_dask` branch maps `_lazy_factorize_wrapper` over blocks)."""
        from .factorize import _factorize_multiple, _is_int_index, int_label_encoder

        ex0 = expected_groups[0]
        if (len(bys) == 1 and is_chunked(bys[0]) and isinstance(ex0, pd.RangeIndex) and ex0.start == 0 and ex0.step == 1
                and bys[0].dtype.kind in "iu" and bys[0].dtype.itemsize <= (8 if bys[0].dtype.kind == "i" else 4)):
            # integer labels against 0..G-1: the labels ARE the codes and the kernels skip whatever is outside [0, G)
            # (the reference's copy + mask, factorize.py:44-53, fused away) — no per-chunk work at all
            return (bys[0],), (ex0,), (len(ex0),)
        if len(bys) == 1 and is_chunked(bys[0]) and _is_int_index(expected_groups[0]):
            first = bys[0].chunks[0]
            if is_torch(first) and first.dtype in (torch.int32, torch.int64):
                enc = int_label_encoder(expected_groups[0], sort)
                if enc.mode == "passthrough":  # labels already are the codes: nothing to do per chunk
                    return (bys[0],), (enc.groups,), (len(enc.groups),)
                if enc.mode == "lut":
                    return (bys[0].map(enc),), (enc.groups,), (len(enc.groups),)

        template = next(b for b in bys if is_chunked(b))
        per_by_chunks = []
        for b in bys:
            if is_chunked(b):
                per_by_chunks.append(b.rechunk_like(template).chunks)
            else:  # in-memory labels next to chunked ones: slice them by the template's chunk boundaries
                bounds = np.concatenate([[0], np.cumsum(template.chunk_sizes)])
                per_by_chunks.append([b[..., int(lo):int(hi)] for lo, hi in zip(bounds[:-1], bounds[1:])])
        out_chunks, found, grp_shape = [], None, None
        for parts in zip(*per_by_chunks):
            (codes,), found, grp_shape = _factorize_multiple(tuple(parts), expected_groups, sort=sort)
            out_chunks.append(codes)
        return (ChunkedArray(out_chunks),), found, grp_shape


def is_chunked(x) -> bool:
    return isinstance(x, ChunkedArray)
