"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch).

flox's map-reduce has exactly one exchange step: the tree combine of per-chunk intermediates
(`_tree_reduce(combine=_simple_combine)`, flox/dask.py:440-460; the scheduler moves python dicts).
Here every rank accumulates ITS chunks into `[A, B, G]` wide accumulators and the combine is an
all-reduce of those accumulators — sum for sums / counts, max / min for extremum keys and indices.
The payload is tiny next to the inputs (G x 24 B at config 2), so the collective is latency-bound
and NCCL's own algorithm choice (NVLS in-switch reduction when available) is what we want; there
is no compute to overlap it with inside a chunk kernel.

`Comm` wraps a process group; `LOCAL` (no group) turns every collective into a no-op so the same
code runs single-GPU.
"""
from __future__ import annotations

import contextlib
import contextvars
from dataclasses import dataclass

import torch
import torch.distributed as dist


class Comm:
    """Collectives on accumulator tensors.  `group=None` + `world_size==1` -> no-ops."""

    def __init__(self, group=None, enabled: bool | None = None):
        self.group = group
        if enabled is None:
            enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.enabled = enabled
        self.collectives = 0

    @property
    def world_size(self) -> int:
        return dist.get_world_size(self.group) if self.enabled else 1

    @property
    def rank(self) -> int:
        return dist.get_rank(self.group) if self.enabled else 0

    def _reduce(self, t: torch.Tensor, op) -> torch.Tensor:
        if self.enabled and t.numel():
            dist.all_reduce(t, op=op, group=self.group)
            self.collectives += 1
        return t

    def sum_(self, *tensors):
        for t in tensors:
            self._reduce(t, dist.ReduceOp.SUM)

    def max_(self, t):
        return self._reduce(t, dist.ReduceOp.MAX)

    def min_(self, t):
        return self._reduce(t, dist.ReduceOp.MIN)

    def prod_(self, t):
        return self._reduce(t, dist.ReduceOp.PRODUCT)

    def peer(self, nbytes: int, device: torch.device):
        """The one-launch NVLink peer-memory exchange for an accumulator set of `nbytes`, or None (-> NCCL): see
        flox_b200/exchange.py.  Every rank must ask with the same `nbytes` in the same order."""
        if not self.enabled:
            return None
        from .exchange import peer_exchange

        return peer_exchange(self.group, nbytes, device)

    def all_gather_pairs(self, pair: torch.Tensor) -> list:
        """[[lo, hi] of every rank] from a 2-element int64 device tensor (one tiny collective + one host read)."""
        if not self.enabled:
            return [pair.tolist()]
        out = torch.empty((self.world_size, 2), dtype=torch.int64, device=pair.device)
        dist.all_gather_into_tensor(out, pair.reshape(1, 2).contiguous(), group=self.group)
        self.collectives += 1
        return out.tolist()

    def all_gather_object(self, obj):
        if not self.enabled:
            return [obj]
        out = [None] * self.world_size
        dist.all_gather_object(out, obj, group=self.group)
        return out


LOCAL = Comm(enabled=False)


def bind_to_gpu_numa(device_index: int | None = None) -> list[int] | None:
    """Pin this process (and the label-pack worker threads it will start) to the CPU cores that are NUMA-local to its
    GPU, so that pinned staging buffers are first-touched on that node and host->device copies do not cross the
    socket interconnect.  One process per GPU on a two-socket box otherwise lands its buffers wherever the launcher
    happened to run it: round 1 measured 28 GB/s per GPU at 8 GPUs, half of what one GPU gets alone.
    Call it BEFORE allocating pinned memory.  Returns the CPU list, or None if NVML / affinity are unavailable
    (FLOXB_NUMA=0 disables it)."""
    import os

    if os.environ.get("FLOXB_NUMA", "1") == "0" or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        import pynvml

        pynvml.nvmlInit()
        idx = torch.cuda.current_device() if device_index is None else int(device_index)
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis and all(v.strip().isdigit() for v in vis.split(",")) and idx < len(vis.split(",")):
            idx = int(vis.split(",")[idx])
        handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
        local = {w * 64 + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = sorted(local & allowed)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:  # noqa: BLE001 - no NVML, no permission, exotic topology: stay unpinned
        return None


@dataclass
class ParallelContext:
    """How chunked inputs map onto ranks.

    sharding = "replicated": every rank was handed the same ChunkedArray and reduces the chunks
                             `assignment(rank)` gives it (contiguous blocks by default, so sorted
                             labels stay sorted per rank: the cohorts layout of SURVEY §8e);
               "local":      the ChunkedArray a rank holds IS its shard.
    """

    comm: Comm = LOCAL
    sharding: str = "replicated"
    # "replicated": every rank ends with the full result (accumulators all-reduced).
    # "sharded":    method="cohorts" / "blockwise" only — the result stays sharded BY GROUP like the reference's cohorts
    #               output stays chunked by cohort (flox/dask.py:469-518): a rank's result is complete for the groups its
    #               rows touch (boundary groups that straddle ranks are merged and reported by every rank that holds them)
    #               and holds the fill value elsewhere.  Only the straddling groups cross NVLink.
    output: str = "replicated"

    def my_chunks(self, nchunks: int) -> range:
        if self.sharding == "local" or not self.comm.enabled:
            return range(nchunks)
        w, r = self.comm.world_size, self.comm.rank
        base, extra = divmod(nchunks, w)
        lo = r * base + min(r, extra)
        return range(lo, lo + base + (1 if r < extra else 0))


# The active context is per THREAD (and per asyncio task): dask's threaded scheduler calls the per-chunk driver from many
# threads at once (the reference's race regression: tests/test_core.py:1827-1851), and one thread entering
# `distributed()` must not switch another thread's calls to multi-rank mode.
_DEFAULT = ParallelContext()
_current: contextvars.ContextVar = contextvars.ContextVar("flox_b200_parallel_context", default=_DEFAULT)


def current() -> ParallelContext:
    return _current.get()


@contextlib.contextmanager
def distributed(group=None, sharding: str = "replicated", output: str = "replicated"):
    """Run `groupby_reduce` on ChunkedArrays across the ranks of `group` (default: WORLD)."""
    if sharding not in ("replicated", "local"):
        raise ValueError("sharding must be 'replicated' or 'local'")
    if output not in ("replicated", "sharded"):
        raise ValueError("output must be 'replicated' or 'sharded'")
    ctx = ParallelContext(Comm(group), sharding, output)
    token = _current.set(ctx)
    try:
        yield ctx
    finally:
        _current.reset(token)
