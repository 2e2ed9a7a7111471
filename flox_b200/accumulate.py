"""Wide device-side accumulators for one grouped-reduction problem.

One `GroupedAccumulators` covers a list of chunks `values[A, n_k, B]` + `codes[n_k]` that are to
be reduced TOGETHER into `[A, B, G]` bins: a single numpy-style call has one chunk, flox's
map-reduce over the chunks resident on this GPU has many (the per-chunk `chunk_reduce` and the
`_simple_combine` of flox/dask.py:90-144 fuse into one accumulation).  Each kernel family runs at
most ONE pass over the data and is cached, so `mean` (= sum + nanlen in the reference,
flox/aggregations.py:328-336) or `nanmax` + its min_count `nanlen` (aggregations.py:997-1014)
cost one read of the inputs, not two.
"""
from __future__ import annotations

from typing import Sequence

import torch

from . import _lib as L
from .kernels import PreparedChunks, acc_dtype, get_kernels


class GroupedAccumulators:
    SPARSE_MAX_CONTESTED = 65536  # more straddling groups than this: a dense all-reduce is the better merge

    def __init__(self, chunks: Sequence[tuple[torch.Tensor, torch.Tensor]], A: int, B: int, G: int,
                 value_dtype: torch.dtype, device: torch.device, index_base: Sequence[int] | None = None, comm=None,
                 sparse_merge: bool = False):
        self.chunks = chunks if isinstance(chunks, PreparedChunks) else list(chunks)
        self.A, self.B, self.G = int(A), int(B), int(G)
        self.nbins = self.A * self.B * self.G
        self.value_dtype = value_dtype
        self.device = device
        if index_base is None:
            index_base, pos = [], 0
            for _, c in self.chunks:
                index_base.append(pos)
                pos += c.numel()
        self.index_base = list(index_base)
        self._moments: dict[tuple, tuple] = {}
        self._fchunks = None
        self._gindex: dict[int, tuple] = {}
        # accumulator layout: [A][G] when B == 1, group-major [G][A*B] when B > 1 (include/flox_b200.h)
        self.gm_ab = self.A * self.B if self.B > 1 else 0
        self._minmax: dict[bool, tuple] = {}
        self._var: dict[bool, tuple] = {}
        self._arg: dict[tuple, torch.Tensor] = {}
        self._pick: dict[tuple, tuple] = {}
        self.k = get_kernels()
        # cross-rank merge of every pass (parallel.Comm); None = single process.  Merging the
        # raw accumulators right after each pass IS flox's combine stage (flox/dask.py:90-144).
        self.comm = comm if (comm is not None and comm.enabled) else None
        # Group-sharded output (parallel.distributed(output="sharded")): merge only the groups whose rows straddle ranks.
        self.sparse_merge = bool(sparse_merge) and self.comm is not None
        self._contested_idx = False  # False = not determined yet; None = too many (dense merge); LongTensor otherwise

    # ---- allocation ----------------------------------------------------------------------
    def _i64(self):
        return torch.empty(self.nbins, dtype=torch.int64, device=self.device)

    def _f64(self):
        return torch.empty(self.nbins, dtype=torch.float64, device=self.device)

    def _count_pair(self):
        """(both, cnt, nnan): the two i64 counters share one buffer so that ranks merge them with ONE collective."""
        both = torch.empty(2 * self.nbins, dtype=torch.int64, device=self.device)
        return both, both[: self.nbins], both[self.nbins:]

    # ---- sparse (cohort-sharded) merge --------------------------------------------------------------------------
    def _contested(self, cnt, nnan):
        """Indices of the groups that may have rows on more than one rank, from the ranks' ranges of occupied groups
        [lo_r, hi_r] (one 16-byte all-gather + one host read, once per accumulator set).  With label-sorted shards —
        method="cohorts" on sorted labels, BASELINE config 5 — that is at most one group per rank boundary; ranges that
        overlap widely (random labels) give None and the merge falls back to the dense all-reduce."""
        if self._contested_idx is not False:
            return self._contested_idx
        # label-derived: a loop over the same collection would otherwise stall on the host read below every call (measured
        # 0.24 ms per call at BASELINE config 5's per-GPU share).  Reused only when the caller DECLARED the labels static
        # (`ChunkedArray.static_labels = True`, like dask graphs are built once for fixed labels and flox memoises its cohort
        # detection, flox/cache.py): the decision below involves a collective, so every rank must take the same branch —
        # a rank-local "my labels did not change" test cannot guarantee that.
        memo = getattr(self.chunks, "memo", None) if getattr(self.chunks, "static_labels", False) else None
        mkey = ("contested", self.comm.world_size, self.comm.rank, self.chunks.label_versions()) if memo is not None else None
        if memo is not None and mkey in memo:
            self._contested_idx = memo[mkey]
            return self._contested_idx
        G = self.G
        rows = cnt + nnan
        if self.B > 1:
            rows = rows.view(G, self.A * self.B).sum(1)   # group-major bins
        elif self.A > 1:
            rows = rows.view(self.A, G).sum(0)
        idx = torch.arange(G, device=self.device)
        occupied = rows > 0
        pair = torch.stack((torch.where(occupied, idx, G).min(), torch.where(occupied, idx, -1).max()))
        ranges = [(lo, hi) for lo, hi in self.comm.all_gather_pairs(pair) if hi >= lo]
        events = sorted([(lo, 1) for lo, _ in ranges] + [(hi + 1, -1) for _, hi in ranges])
        spans, depth, start = [], 0, None
        for pos, d in events:           # sweep: maximal intervals covered by >= 2 ranges
            if depth >= 2 and d < 0 and depth + d < 2 and start is not None:
                spans.append((start, pos))
                start = None
            depth += d
            if depth >= 2 and start is None:
                start = pos
        total = sum(b - a for a, b in spans)
        AB = self.A * self.B
        if total * AB > self.SPARSE_MAX_CONTESTED:
            self._contested_idx = None
        else:
            groups = (torch.cat([torch.arange(a, b, device=self.device) for a, b in spans]) if spans
                      else torch.empty(0, dtype=torch.int64, device=self.device))
            if self.B > 1:      # bins of a group: g * AB + (0 .. AB-1)
                groups = (groups[:, None] * AB + torch.arange(AB, device=self.device)[None, :]).reshape(-1)
            elif self.A > 1:    # bins of a group: a * G + g
                groups = (torch.arange(self.A, device=self.device)[:, None] * G + groups[None, :]).reshape(-1)
            self._contested_idx = groups
        if memo is not None:
            memo.clear()  # one entry: the labels changed or the process group did
            memo[mkey] = self._contested_idx
        return self._contested_idx

    def _reduce_small(self, t, how: str):
        {"sum": self.comm.sum_, "prod": self.comm.prod_, "max": self.comm.max_, "min": self.comm.min_}[how](t)
        return t

    def _acc_and_counts(self, dtype: torch.dtype):
        """(px, acc, both, cnt, nnan).  With several ranks and a small bin count the accumulators live in the NVLink
        peer-exchange buffer (px is the PeerExchange): the kernels accumulate straight into it and ONE launch merges
        value accumulator and counters across ranks (flox_b200/exchange.py)."""
        px = self.comm.peer(24 * self.nbins, self.device) if (self.comm and self.nbins and not self.sparse_merge) else None
        if px is not None:
            acc, both = px.views([(dtype, self.nbins), (torch.int64, 2 * self.nbins)])
        else:
            acc = torch.empty(self.nbins, dtype=dtype, device=self.device)
            both = torch.empty(2 * self.nbins, dtype=torch.int64, device=self.device)
        return px, acc, both, both[: self.nbins], both[self.nbins:]

    def _merge(self, px, acc, both, how: str):
        """Cross-rank merge of one pass (flox's combine stage, flox/dask.py:90-144): `how` for the value accumulator,
        sum for the counters.  Returns (acc, cnt, nnan)."""
        if self.comm and self.sparse_merge and self.nbins:
            n = self.nbins
            idx = self._contested(both[:n], both[n:])
            if idx is not None:
                if idx.numel():  # every rank computed the same index list: the small collectives line up
                    acc[idx] = self._reduce_small(acc[idx], how)
                    small = self._reduce_small(torch.cat((both[:n][idx], both[n:][idx])), "sum")
                    both[:n][idx] = small[: idx.numel()]
                    both[n:][idx] = small[idx.numel():]
                return acc, both[:n], both[n:]
        if self.comm:
            if px is not None:
                acc, both = px.merge([acc, both], [how, "sum"])
            else:
                {"sum": self.comm.sum_, "prod": self.comm.prod_, "max": self.comm.max_, "min": self.comm.min_}[how](acc)
                self.comm.sum_(both)
        return acc, both[: self.nbins], both[self.nbins:]

    def _single(self, dtype: torch.dtype):
        px = self.comm.peer(8 * self.nbins, self.device) if (self.comm and self.nbins and not self.sparse_merge) else None
        if px is not None:
            (t,) = px.views([(dtype, self.nbins)])
        else:
            t = torch.empty(self.nbins, dtype=dtype, device=self.device)
        return px, t

    def _merge_single(self, px, t, how: str):
        if self.comm and self.sparse_merge and self.nbins and self._contested_idx is not False and self._contested_idx is not None:
            idx = self._contested_idx  # determined by the first pass of this accumulator set (moments / minmax)
            if idx.numel():
                t[idx] = self._reduce_small(t[idx], how)
            return t
        if self.comm:
            if px is not None:
                (t,) = px.merge([t], [how])
            else:
                {"sum": self.comm.sum_, "min": self.comm.min_, "max": self.comm.max_}[how](t)
        return t

    # ---- kernel families -----------------------------------------------------------------
    def _chunks_as(self, as_float: bool):
        """Integer values that feed a FLOAT result are summed as float64, like the reference's
        `ufunc.reduceat(..., dtype=float64)` (flox/aggregate_flox.py:183-190): converted once, cached."""
        if not as_float or self.value_dtype.is_floating_point:
            return self.chunks, self.value_dtype
        if self._fchunks is None:
            self._fchunks = [(v.to(torch.float64), c) for v, c in self.chunks]
        return self._fchunks, torch.float64

    # ---- column-owner path (kept trailing dims, B > 1) ------------------------------------------
    COLS_MIN_COLUMNS = 32768  # enough independent columns to fill 148 SMs with 16-byte loads

    ROWS_MIN_ROWS = 296        # row-batched (A >> 1, B == 1): at least two warps' worth of rows per SM
    ROWS_MIN_GROUP_RUN = 32    # ... and groups long enough to keep a warp's lanes busy

    def _use_cols(self) -> bool:
        """Group-index (register-accumulating, atomic-free) kernels: column-owner for kept trailing
        dims, warp-per-row for many rows sharing one label vector."""
        if self.nbins == 0 or not any(c.numel() for _, c in self.chunks):
            return False
        if self.B > 1:
            return self.A * self.B >= self.COLS_MIN_COLUMNS and self.A <= 65535
        n_total = sum(c.numel() for _, c in self.chunks)
        return self.A >= self.ROWS_MIN_ROWS and n_total >= self.ROWS_MIN_GROUP_RUN * self.G * len(self.chunks)

    def _group_index(self, k: int):
        """(perm, goff) of chunk k: rows grouped by code (one counting sort over the LABELS only)."""
        if k not in self._gindex:
            codes = self.chunks[k][1]
            perm = torch.empty(codes.numel(), dtype=torch.int64, device=self.device)
            goff = torch.empty(self.G + 1, dtype=torch.int64, device=self.device)
            cursor = torch.empty(max(self.G, 1), dtype=torch.int64, device=self.device)
            self.k.group_index(codes, self.G, perm, goff, cursor)
            self._gindex[k] = (perm, goff)
        return self._gindex[k]

    def _reduce_cols(self, chunks, mode: int, acc, cnt, nnan, group_major: bool):
        first = True
        for k, (v, c) in enumerate(chunks):
            if c.numel() == 0:
                continue
            perm, goff = self._group_index(k)
            self.k.reduce_cols(v, perm, goff, self.A, c.numel(), self.B, self.G, mode, acc, cnt, nnan, init=first,
                               group_major=group_major)
            first = False

    def moments(self, op: int = L.OP_SUM, as_float: bool = False):
        """(acc, cnt, nnan): acc = op over non-NaN values, cnt = # non-NaN values, nnan > 0 iff NaNs."""
        as_float = bool(as_float) and not self.value_dtype.is_floating_point
        if (op, as_float) not in self._moments:
            chunks, vdt = self._chunks_as(as_float)
            px, acc, both, cnt, nnan = self._acc_and_counts(acc_dtype(vdt))
            if self._use_cols():
                self._reduce_cols(chunks, op, acc, cnt, nnan, self.B > 1)
            else:
                self.k.moments(chunks, self.A, self.B, self.G, op, acc, cnt, nnan, init=True)
            self._moments[(op, as_float)] = self._merge(px, acc, both, "prod" if op == L.OP_PROD else "sum")
        return self._moments[(op, as_float)]

    def minmax(self, is_max: bool):
        if is_max not in self._minmax:
            px, key, both, cnt, nnan = self._acc_and_counts(torch.int64)
            if self._use_cols():
                self._reduce_cols(self.chunks, 3 if is_max else 4, key, cnt, nnan, self.B > 1)
            else:
                self.k.minmax(self.chunks, self.A, self.B, self.G, is_max, key, cnt, nnan, init=True,
                              value_dtype=self.value_dtype)
            self._minmax[is_max] = self._merge(px, key, both, "max" if is_max else "min")
        return self._minmax[is_max]

    def counts(self):
        """(cnt, nnan) from whichever pass already ran, else from a SUM pass."""
        for store in (self._moments, self._minmax):
            for v in store.values():
                return v[1], v[2]
        _, cnt, nnan = self.moments(L.OP_SUM)
        return cnt, nnan

    def var(self, skipna: bool):
        """(m2, s_acc, cnt, nnan): two-pass centred sum of squared deviations (var_chunk).
        Always a float computation: integer values go through float64 (aggregations.py:352-387)."""
        if skipna not in self._var:
            chunks, vdt = self._chunks_as(True)
            acc, cnt, nnan = self.moments(L.OP_SUM, as_float=True)
            means = self._f64()
            px, m2 = self._single(torch.float64)
            self.k.group_means(vdt, acc, cnt, nnan, skipna, means)
            # means come from the ALREADY MERGED sums, so M2 is about the global group mean and the
            # cross-rank merge is a plain sum (no Chan adjustment terms needed, cf. aggregations.py:392-451)
            self.k.sqdev(chunks, self.A, self.B, self.G, means, skipna, m2, init=True, value_dtype=vdt)
            m2 = self._merge_single(px, m2, "sum")
            self._var[skipna] = (m2, acc, cnt, nnan)
        return self._var[skipna]

    def argext(self, is_max: bool, skipna: bool):
        key_ = (is_max, skipna)
        if key_ not in self._arg:
            key, cnt, nnan = self.minmax(is_max)
            px, idx = self._single(torch.int64)
            self.k.argext(self.chunks, self.A, self.B, self.G, key, nnan, skipna, self.index_base, idx, init=True,
                          value_dtype=self.value_dtype)
            # key / nnan are already global, so local candidates are comparable
            self._arg[key_] = self._merge_single(px, idx, "min")
        return self._arg[key_]

    def picked(self, is_last: bool, skipna: bool):
        """(val, cnt, nnan): the value of each group's first / last row (first / last NON-NaN row when skipna), in the wide
        accumulator dtype.  Positions are global along the reduced axis (index_base), so chunks and ranks merge by MIN / MAX
        of the position; only the winner's owner writes the value and ranks merge it by SUM over zeros."""
        key_ = (is_last, skipna)
        if key_ not in self._pick:
            cnt, nnan = self.counts()
            px, pos = self._single(torch.int64)
            self.k.firstlast(self.chunks, self.A, self.B, self.G, is_last, skipna, self.index_base, pos, init=True,
                             value_dtype=self.value_dtype)
            pos = self._merge_single(px, pos, "max" if is_last else "min")
            px, val = self._single(acc_dtype(self.value_dtype))
            self.k.take(self.chunks, self.A, self.B, self.G, pos, self.index_base, val, init=True, value_dtype=self.value_dtype)
            val = self._merge_single(px, val, "sum")
            self._pick[key_] = (val, cnt, nnan)
        return self._pick[key_]

    # ---- typed results -------------------------------------------------------------------
    def shape(self, lead: tuple[int, ...] = ()):
        return tuple(lead) + (self.G,)

    def finalize(self, kind: int, acc, acc2, cnt, nnan, *, skipna: bool, out_dtype: torch.dtype, fill=None,
                 min_count: int = 0, mc_fill=None, ddof: int = 0, value_dtype=None) -> torch.Tensor:
        out = torch.empty(self.nbins, dtype=out_dtype, device=self.device)
        self.k.finalize(kind, value_dtype or self.value_dtype, acc, acc2, cnt, nnan, skipna, fill, min_count, mc_fill, ddof, out,
                        group_major_ab=self.gm_ab)
        return out
