"""One-launch cross-GPU merge of small accumulator sets over NVLink peer memory (csrc/exchange.cu).

flox's map-reduce combine (`_tree_reduce(combine=_simple_combine)`, flox/dask.py:440-460) becomes, per pass over the
data, ONE kernel that barriers over peer memory and sums / maxes / mins the ranks' accumulators in rank order — instead
of two NCCL all-reduces (sums, counts) whose fixed cost is comparable to an 8-GPU share of the chunk kernel.

A `PeerExchange` owns this rank's slab (header + two buffers, CUDA IPC-exported) and the mappings of every peer's slab.
The reduction kernels accumulate straight into views of the current buffer (`views`), `merge` launches the exchange and
returns the merged tensors in ordinary torch memory.  Everything is stream-ordered; nothing synchronises.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist

from . import _lib as L

_OPS = {"sum": (L.XCHG_SUM_F64, L.XCHG_SUM_I64), "max": (None, L.XCHG_MAX_I64), "min": (None, L.XCHG_MIN_I64),
        "prod": (L.XCHG_PROD_F64, L.XCHG_PROD_I64)}


class _RawCuda:
    """`__cuda_array_interface__` view of raw device memory, for torch.as_tensor."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerExchange:
    def __init__(self, group, buffer_bytes: int, device: torch.device):
        self.group, self.device = group, device
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.buffer_bytes = (int(buffer_bytes) + 255) // 256 * 256
        self.epoch = 0
        self._pending = None  # views handed out for the next collective
        lib = L.load()
        total = L.XCHG_HEADER_BYTES + 2 * self.buffer_bytes
        ptr, handle = C.c_void_p(), (C.c_ubyte * 64)()
        rc = lib.floxb_ipc_alloc(total, C.byref(ptr), handle)
        mine = bytes(handle) if rc == 0 else None
        handles = [None] * self.world
        dist.all_gather_object(handles, mine, group=group)
        self.own = ptr.value if rc == 0 else None
        self.bases, self._opened = [], []
        ok = all(h is not None for h in handles)
        if ok:
            for r, h in enumerate(handles):
                if r == self.rank:
                    self.bases.append(self.own)
                    continue
                p = C.c_void_p()
                if lib.floxb_ipc_open((C.c_ubyte * 64).from_buffer_copy(h), C.byref(p)) != 0:
                    ok = False
                    break
                self._opened.append(p.value)
                self.bases.append(p.value)
        flags = [None] * self.world
        dist.all_gather_object(flags, bool(ok), group=group)
        self.ok = all(flags)
        if not self.ok:
            self.close()
            return
        self._bases_c = (C.c_void_p * self.world)(*self.bases)
        self._buffers = [torch.as_tensor(_RawCuda(self.own + L.XCHG_HEADER_BYTES + i * self.buffer_bytes, self.buffer_bytes),
                                         device=device) for i in range(2)]
        torch.cuda.synchronize(device)
        dist.barrier(group=group)  # every rank has mapped every slab before the first collective touches one

    # ---- one collective = views(...) -> kernels fill them -> merge(...) ------------------------------------------------
    def fits(self, nbytes: int) -> bool:
        return self.ok and nbytes <= self.buffer_bytes

    def views(self, specs):
        """Tensors inside the buffer of the NEXT collective, laid out back to back: specs = [(torch dtype, numel), ...]
        (8-byte dtypes).  The kernels accumulate into them; pass them to `merge` in the same order."""
        # (views handed out but never merged belong to a call that raised on every rank alike: simply reuse the buffer)
        buf = self._buffers[(self.epoch + 1) % 2]
        out, off = [], 0
        for dt, n in specs:
            assert dt in (torch.float64, torch.int64)
            out.append(buf[off:off + 8 * n].view(dt))
            off += 8 * n
        assert off <= self.buffer_bytes
        self._pending = out
        return out

    def merge(self, tensors, ops):
        """Launch the exchange for the views handed out by `views` and return the merged tensors (local memory)."""
        pend = self._pending
        assert pend is not None and len(tensors) == len(pend) and all(a is b for a, b in zip(tensors, pend))
        self._pending = None
        self.epoch += 1
        nseg = len(tensors)
        assert nseg <= L.XCHG_MAX_SEGMENTS
        outs = [torch.empty_like(t) for t in tensors]
        words = (C.c_int64 * nseg)(*[t.numel() for t in tensors])
        opc = (C.c_int * nseg)(*[_OPS[op][0 if t.dtype == torch.float64 else 1] for t, op in zip(tensors, ops)])
        outp = (C.c_void_p * nseg)(*[o.data_ptr() for o in outs])
        off = L.XCHG_HEADER_BYTES + (self.epoch % 2) * self.buffer_bytes
        rc = L.load().floxb_xchg_allreduce(self._bases_c, self.world, self.rank, self.epoch, off, nseg, words, opc, outp,
                                           torch.cuda.current_stream().cuda_stream)
        L.check(rc, "floxb_xchg_allreduce")
        from .kernels import get_kernels

        k = get_kernels()
        k.launches += 1
        k.kernel_launches += 1  # bench.py reports the library's launches per step
        return outs

    def status(self) -> int:
        """0, or the epoch of the first collective whose barrier timed out (synchronises the device)."""
        v = C.c_uint64(0)
        L.check(L.load().floxb_xchg_status(self.own, C.byref(v)), "floxb_xchg_status")
        return int(v.value)

    def close(self):
        lib = L.load()
        for p in self._opened:
            lib.floxb_ipc_close(p)
        self._opened = []
        if self.own:
            lib.floxb_ipc_free(self.own)
            self.own = None
        self.ok = False


_exchanges: dict = {}


def peer_exchange(group, nbytes: int, device: torch.device):
    """The process group's PeerExchange if this collective can use one, else None (-> NCCL).  Creation is itself
    collective: every rank reaches it in the same call because accumulator sizes are the same on all ranks."""
    if os.environ.get("FLOXB_PEER_XCHG", "1") == "0" or device.type != "cuda":
        return None
    cap = int(os.environ.get("FLOXB_PEER_XCHG_BYTES", str(4 << 20)))
    if nbytes > cap:
        return None
    key = group if group is not None else dist.group.WORLD  # the ProcessGroup object itself: a new group is a new key
    ex = _exchanges.get(key, False)
    if ex is False:
        try:
            ok_world = dist.get_world_size(group) <= L.XCHG_MAX_RANKS and dist.get_backend(group) == "nccl"
        except Exception:  # noqa: BLE001
            ok_world = False
        ex = PeerExchange(group, cap, device) if ok_world else None
        if ex is not None and not ex.ok:
            ex = None
        _exchanges[key] = ex
    return ex if (ex is not None and ex.fits(nbytes)) else None


def close_all():
    for ex in _exchanges.values():
        if ex:
            ex.close()
    _exchanges.clear()
