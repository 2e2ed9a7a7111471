"""Chunked (dask-like) inputs: map-reduce / blockwise / cohorts on the GPU(s).

Replaces `dask_groupby_agg` + `_simple_combine` + `_aggregate` + `_expand_dims`
(flox/dask.py:325-571, 90-175).  The reference builds a task graph: one `chunk_reduce` per block
(dask.py:395-429), a tree of `_simple_combine` nodes (dask.py:440-460) and a final `_aggregate`.
On a GPU the chunks owned by a rank stream through ONE persistent kernel launch into shared
accumulators (chunk_reduce and the combine tree fuse), ranks merge accumulators with an
all-reduce (parallel.py), and the reference's finalize semantics are applied to the merged result:

* method="map-reduce": intermediates are evaluated with the INTERMEDIATE fills / dtypes
  (aggregations.py:973-984), then `_finalize_results` runs agg.finalize and the min_count mask —
  an absent group therefore finishes as e.g. 0 for `sum`, exactly like the reference's reindexed
  map-reduce;
* method="blockwise" / "cohorts": every group is completed by the blocks / cohort that hold it,
  the result is the numpy-path value (dask.py:390-392,469-518) and absent groups get the final fill.

`_simple_combine` / `_aggregate` / `_expand_dims` are also provided for explicit per-chunk
intermediates (API parity and the golden tests).
"""
from __future__ import annotations

import math
import os
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pandas as pd
import torch

from . import parallel
from .accumulate import GroupedAccumulators
from .aggregations import Aggregation, _is_arg_reduction, is_var_chunk_reduction, var_chunk
from .arrays import count_h2d, from_device, is_torch, kernel_dtype, to_device
from .chunked import ChunkedArray, is_chunked
from .core import DUMMY_AXIS, _finalize_results
from .kernels import PreparedChunks, default_device, get_kernels
from .multiarray import MultiArray
from .reductions import evaluate, var_triple
from .reindex import ReindexStrategy


# ---- explicit intermediates (reference-shaped API) ---------------------------------------------

def _expand_dims(results: dict) -> dict:
    """Insert DUMMY_AXIS into every intermediate (flox/dask.py:154-158)."""
    def one(a):
        if isinstance(a, MultiArray):
            return MultiArray(tuple(one(x) for x in a.arrays))
        return a.unsqueeze(DUMMY_AXIS) if is_torch(a) else np.expand_dims(a, DUMMY_AXIS)

    results["intermediates"] = tuple(one(a) for a in results["intermediates"])
    return results


def _concat(parts, axis):
    if isinstance(parts[0], MultiArray):
        return MultiArray(tuple(_concat([p.arrays[i] for p in parts], axis) for i in range(len(parts[0].arrays))))
    return torch.cat(list(parts), dim=axis) if is_torch(parts[0]) else np.concatenate(parts, axis=axis)


def _simple_combine(x_chunk, agg: Aggregation, axis, keepdims: bool, reindex: ReindexStrategy | None = None,
                    is_aggregate: bool = False) -> dict:
    """Concatenate per-chunk intermediates on DUMMY_AXIS and reduce it (flox/dask.py:90-144).
    `x_chunk`: list of chunk_reduce dicts already passed through `_expand_dims`."""
    results = {"groups": x_chunk[0]["groups"], "intermediates": []}
    for idx, combine in enumerate(agg.simple_combine):
        if combine is None:
            raise NotImplementedError("arg-reductions need the grouped (value, index) combine")
        stacked = _concat([x["intermediates"][idx] for x in x_chunk], DUMMY_AXIS)
        nd = stacked.ndim
        result = combine(stacked, axis=(nd + DUMMY_AXIS,), keepdims=True)
        if is_aggregate:
            result = result.squeeze(nd + DUMMY_AXIS)
        results["intermediates"].append(result)
    return results


def _aggregate(x_chunk, combine, agg: Aggregation, expected_groups, axis, keepdims: bool, fill_value, reindex) -> dict:
    """Last node of the tree: combine, then finalize (flox/dask.py:161-175)."""
    results = combine(x_chunk, agg, axis, keepdims, is_aggregate=True)
    return _finalize_results(results, agg, axis, expected_groups, reindex=reindex)


# ---- fused GPU path -------------------------------------------------------------------------------

_CODE_DTYPES = (torch.int32, torch.int64)
_VALUE_DTYPES = (torch.float32, torch.float64, torch.int32, torch.int64)


_pack_pool: ThreadPoolExecutor | None = None
_pack_pool_lock = threading.Lock()


def _pool(world: int) -> ThreadPoolExecutor:
    global _pack_pool
    if _pack_pool is not None:
        return _pack_pool
    with _pack_pool_lock:  # concurrent first calls (dask threads) must end up sharing ONE pool
        if _pack_pool is not None:
            return _pack_pool
        # the pack is bound by host memory bandwidth: on the B200 box 16 threads are the optimum (107 ms per 8 GB of
        # int64 labels; 32 threads 131 ms, 128 threads 307 ms — tools/e2e_probe.py)
        want = int(os.environ.get("FLOXB_PACK_THREADS", "16"))
        _pack_pool = ThreadPoolExecutor(max_workers=max(1, min(want, (os.cpu_count() or 1) // max(world, 1))),
                                        thread_name_prefix="floxb-pack")
    return _pack_pool


def _host_codes(c):
    """A C-contiguous 1-D int32/int64 numpy view of a HOST label chunk, else None."""
    if is_torch(c):
        if c.device.type != "cpu" or c.dtype not in _CODE_DTYPES or not c.is_contiguous():
            return None
        c = c.numpy()
    elif not isinstance(c, np.ndarray):
        return None
    if c.dtype not in (np.dtype("int32"), np.dtype("int64")) or c.ndim != 1 or not c.flags.c_contiguous:
        return None
    return c


def _stage_packed_codes(label_chunks: dict, n_groups: int, device, world: int) -> dict:
    """Host-resident integer label chunks -> device int32 codes, moving 2 (4) instead of 8 bytes per row over
    PCIe when the groups fit 16 (32) bits: worker threads range-check + narrow each chunk into a pinned staging
    buffer (floxb_host_pack_codes), the copies are enqueued as the packs finish, int16 is widened on the device
    (floxb_widen_codes).  Returns {chunk index: device codes}; chunks it does not take are left to the caller."""
    policy = os.environ.get("FLOXB_PACK_LABELS", "auto")
    # The pack is bound by what the HOST can stream (~77 GB/s of label reads on the B200 box however many threads
    # or processes share it: profiles/r01_e2e_probe_session19.log, r01_bench_4gpu.log), so with several ranks per
    # host it is slower than letting every rank's own PCIe link copy the int64 labels: auto = single rank only.
    if device.type != "cuda" or n_groups > 2**31 - 1 or policy == "0" or (policy == "auto" and world > 1):
        return {}
    todo = {}
    for k, c in label_chunks.items():
        h = _host_codes(c)
        if h is not None and h.size and (n_groups <= 32767 or h.dtype == np.dtype("int64")):
            todo[k] = h
    if not todo:
        return {}
    sdt = torch.int16 if n_groups <= 32767 else torch.int32
    kern, pool = get_kernels(), _pool(world)
    try:
        staged = {k: torch.empty(h.size, dtype=sdt, pin_memory=True) for k, h in todo.items()}
    except RuntimeError:  # no pinned memory left for the staging buffers: copy the labels as they are
        return {}
    futures = {k: pool.submit(kern.host_pack_codes, h, n_groups, staged[k]) for k, h in todo.items()}
    out = {}
    for k in todo:
        futures[k].result()
        count_h2d(staged[k])
        d = staged[k].to(device, non_blocking=True)
        out[k] = kern.widen_codes(d) if sdt == torch.int16 else d
    return out


def _local_device_chunks(array: ChunkedArray, by, ctx: parallel.ParallelContext, need_global_index: bool = False,
                         n_groups: int | None = None):
    """Device tensors `(values[A, n_k, 1], codes[n_k])` for the chunks this rank owns, plus the
    global position of each chunk's first element along the reduced axis (only arg-reductions need
    positions that are global across ranks: that costs one small all-gather of shard lengths)."""
    device = default_device()
    # Device-resident chunks that need no copy or conversion are prepared ONCE per (array, labels) pair: the views, the
    # C-ABI chunk table and the offsets are reused by every later call (a map-reduce loop over the same dask-like
    # collection).  Chunk CONTENTS may change between calls; swapping a chunk tensor for another one changes the key.
    world = ctx.comm.world_size if ctx.comm.enabled else 1
    key = (id(by), ctx.sharding, ctx.comm.rank if ctx.comm.enabled else 0, world, str(device),
           tuple(map(id, array.chunks)), tuple(map(id, by.chunks)) if is_chunked(by) else None)  # swapped chunks = new key
    cached = getattr(array, "_prepared", None)
    if cached is not None and not need_global_index and cached[0] == key and cached[1] is by:
        return cached[2]
    copied = False
    mine = ctx.my_chunks(array.nchunks)
    sizes = array.chunk_sizes
    start = 0
    if need_global_index and ctx.sharding == "local" and ctx.comm.enabled:
        totals = ctx.comm.all_gather_object(sum(sizes))
        start = sum(totals[: ctx.comm.rank])
    offsets = np.concatenate([[0], np.cumsum(sizes)])[:-1]
    by_chunks = by.rechunk_like(array).chunks if is_chunked(by) else None
    lead = tuple(array.shape[:-1])
    A = math.prod(lead)
    first = array.chunks[0]
    kdt = kernel_dtype(first.dtype if is_torch(first) else to_device(first[..., :0], device).dtype)
    chunks, index_base = [], []
    labels = {k: (by_chunks[k] if by_chunks is not None else by[..., int(offsets[k]):int(offsets[k]) + sizes[k]]) for k in mine}
    values = {}
    # Slabs whose reduced (last) axis is OUTERMOST in memory — a (rows, columns) block handed over as block.T, which is
    # what moving the grouped dim to the end gives (flox/core.py:1014-1018) — are consumed in place as values[1, n, B]
    # (B = the kept columns: BASELINE config 4's layout), like the in-memory path does (aggregate_b200.make_accumulators).
    # decided from ALL chunks of the collection (not only this rank's): the accumulator layout must be the same on every rank
    time_major = bool(lead) and all(
        is_torch(c) and c.device.type == device.type and c.dtype == kdt and not c.is_contiguous()
        and c.movedim(-1, 0).is_contiguous() for c in array.chunks)
    B = A if time_major else 1
    if time_major:
        A = 1
    for k in mine:  # all value copies are enqueued first: the link is busy while the worker threads pack labels
        v = array.chunks[k]
        if time_major:
            if v.device != device:
                copied = True
                v = v.to(device)
            v = v.movedim(-1, 0)
        elif not (is_torch(v) and v.device == device and v.dtype == kdt and v.is_contiguous()):  # else: already in place
            copied = True
            v = to_device(v, device)
            if v.dtype != kdt:
                v = v.to(kdt)
        values[k] = v
    packed = _stage_packed_codes(labels, n_groups, device, ctx.comm.world_size if ctx.comm.enabled else 1) if n_groups else {}
    for k in mine:
        v = values[k]
        c = packed.get(k, labels[k])
        if not (is_torch(c) and c.device == device and c.dtype in _CODE_DTYPES and c.is_contiguous()):
            copied = True
            c = to_device(c, device)
            if c.dtype not in _CODE_DTYPES:
                c = c.to(torch.int64)
        chunks.append((v.reshape(1, -1, B) if time_major else v.view(A, -1, 1), c.view(-1)))
        index_base.append(int(offsets[k]) + start)
    prepared = PreparedChunks(chunks)
    prepared.static_labels = bool(getattr(by, "static_labels", False))
    out = (prepared, index_base, lead, (A, B), kdt, device)
    if not copied and not packed and not need_global_index:
        try:
            array._prepared = (key, by, out)
        except AttributeError:
            pass
    return out


def chunked_groupby_agg(array: ChunkedArray, by, agg: Aggregation, expected_groups: pd.RangeIndex, axis, fill_value,
                        method: str, reindex: ReindexStrategy, engine: str, sort: bool = True):
    """GPU counterpart of flox.dask.dask_groupby_agg (flox/dask.py:325-571).  `by` holds integer
    codes against `expected_groups` (a RangeIndex) and is 1-D along the chunked axis."""
    if by.ndim != 1:
        raise NotImplementedError("chunked inputs are grouped by 1-D labels along the chunked (last) axis")
    ctx = parallel.current()
    G = len(expected_groups)
    positional = _is_arg_reduction(agg) or agg.name in ("first", "last", "nanfirst", "nanlast")
    chunks, index_base, lead, (A, B), kdt, device = _local_device_chunks(array, by, ctx, need_global_index=positional,
                                                                         n_groups=G)
    sparse = ctx.output == "sharded" and method in ("cohorts", "blockwise")
    acc = GroupedAccumulators(chunks, A, B, G, kdt, device, index_base=index_base, comm=ctx.comm, sparse_merge=sparse)
    like = array.chunks[0]
    array_np_dtype = array.dtype
    shape = lead + (G,)

    def typed(t, dt):
        return from_device(t.reshape(shape), like, dt)

    if (method == "map-reduce" and agg.name in ("mean", "nanmean") and agg.min_count == 0 and len(agg.chunk) == 2
            and np.dtype(agg.dtype["final"]).kind == "f"):
        # mean = sum / count (aggregations.py:323-336): ONE epilogue launch does what the reference's finalize does with
        # the two combined intermediates — the sum rounded to its intermediate dtype, divided in f64, rounded to the
        # final dtype; absent groups are 0 / 0 = NaN like the reference's reindexed intermediates
        funcs, fills, dts = (agg.name,), (0,), (agg.dtype["final"],)
        do_finalize = False
    elif method == "map-reduce" and not positional:
        funcs, fills, dts = agg.chunk, agg.fill_value["intermediate"], agg.dtype["intermediate"]
        do_finalize = True
    else:  # blockwise / cohorts / arg-reductions: straight to the numpy-path value
        funcs, fills, dts = agg.numpy, agg.fill_value["numpy"], agg.dtype["numpy"]
        do_finalize = False
    ddof = agg.finalize_kwargs.get("ddof", 0) if agg.finalize_kwargs else 0

    inter, previous = [], None
    for red, fv, dt in zip(funcs, fills, dts):
        if red == "nanlen" and previous == "nanlen":
            inter.append(inter[-1])
        elif is_var_chunk_reduction(red):
            base, pkw = red.func, red.keywords
            skipna = bool(pkw.get("skipna", False))
            out_dt = np.dtype(dt) if dt is not None else (array_np_dtype if array_np_dtype.kind == "f" else np.dtype("float64"))
            if base is var_chunk:
                inter.append(MultiArray(tuple(typed(t, out_dt) for t in var_triple(acc, skipna, out_dt))))
            else:
                name = ("nan" if skipna else "") + ("std" if pkw.get("std", False) else "var")
                inter.append(typed(evaluate(acc, name, out_dtype=out_dt, fill_value=np.nan, ddof=ddof), out_dt))
        elif callable(red):
            raise NotImplementedError("custom chunk callables are supported for in-memory arrays only")
        else:
            out_dt = np.dtype(dt) if dt is not None else array_np_dtype
            inter.append(typed(evaluate(acc, red, out_dtype=out_dt, fill_value=fv), out_dt))
        previous = red

    if not do_finalize:
        agg.finalize = None
    results = {"groups": expected_groups, "intermediates": inter}  # stays a RangeIndex: no O(G) host work
    final = _finalize_results(results, agg, axis, expected_groups, reindex=ReindexStrategy(blockwise=True))
    return final[agg.name], (final["groups"],)
