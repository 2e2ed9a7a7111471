"""Kernel layer: torch tensors in, C-ABI calls out.

`CudaKernels` is the only backend the package ships; every method is a thin, stream-ordered
wrapper over one `floxb_*` entry point of include/flox_b200.h (torch supplies device memory
and the current stream — plumbing only).  There is no CPU implementation here: without the
built library and a CUDA device every call raises.  `set_backend` exists so that the test
suite can check the HOST logic on a CPU-only machine with an emulator that lives under
tests/ (never imported by the package).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Sequence

import torch

from . import _lib as L

_TORCH2CODE = {torch.float32: L.F32, torch.float64: L.F64, torch.int32: L.I32, torch.int64: L.I64}


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _TORCH2CODE[dt]
    except KeyError:
        raise TypeError(f"flox_b200 kernels take float32/float64/int32/int64, got {dt}") from None


def acc_dtype(value_dtype: torch.dtype) -> torch.dtype:
    """Wide accumulator dtype for a value dtype (f64 for floats, i64 for ints)."""
    return torch.float64 if value_dtype.is_floating_point else torch.int64


Chunks = Sequence[tuple[torch.Tensor, torch.Tensor]]  # (values[A, n, B], codes[n])


class PreparedChunks(list):
    """A chunk list whose C-ABI descriptor table (`floxb_chunk_t[]`) is built once and reused: a map-reduce loop hands the
    same device-resident chunks to the kernels every step, and filling 100 ctypes structs per call costs more host time
    than an 8-GPU share of the kernel runs (bench.py: host_enqueue_ms_per_step)."""

    def __init__(self, chunks):
        super().__init__(chunks)
        self.table = None  # (ctypes array, value dtype, code dtype, A, B), filled by CudaKernels._chunk_array
        self.memo: dict = {}  # facts derived from the LABELS (e.g. which groups straddle ranks), see label_versions
        self.static_labels = False  # the caller's promise that the labels stay as they are on EVERY rank (ChunkedArray.static_labels)

    def label_versions(self) -> tuple:
        """torch's per-tensor version counters of the label chunks: unchanged versions on unchanged tensor objects mean
        unchanged label contents, so label-derived facts can be reused (flox memoises its cohort detection the same way,
        flox/cache.py)."""
        return tuple(getattr(c, "_version", -1) for _, c in self)


class CudaKernels:
    name = "cuda"

    def __init__(self):
        self.launches = 0  # number of C-ABI kernel entry calls issued (bench reports it)
        # testing / profiling knob: 1 = force the shared-memory path, 2 = force the L2-atomic path
        self.force_path = int(os.environ.get("FLOXB_FORCE_PATH", "0"))
        # FLOXB_WORKSPACE=0 keeps the f64 level of the f32 shared-memory kernel in shared memory (profiling knob)
        self.use_workspace = os.environ.get("FLOXB_WORKSPACE", "1") != "0"
        # CUDA kernel launches issued by the library (mirrors the launch logic of csrc/*.cu:
        # optional fill kernel + one launch per chunk table (<= floxb_max_chunks_per_launch) on the shared-memory path, or one
        # per non-empty chunk on the L2-atomic path).  bench.py reports it as `gpu_launches`.
        self.kernel_launches = 0
        self._limits: dict = {}

    def _scan_launches(self, chunks, A, B, G, init, family, vdt, force_path, has_smem_path=True):
        nbins = A * B * G
        n = 1 if (init and nbins > 0) else 0
        if nbins > 0:
            nonempty = sum(1 for _, c in chunks if c.numel() > 0)
            key = (vdt, family)
            if key not in self._limits:
                self._limits[key] = self.smem_group_limit(vdt, family)
            smem = has_smem_path and A == 1 and B == 1 and G <= self._limits[key] and force_path != 2
            if smem:
                per = self.max_chunks_per_launch()
                for lo in range(0, len(chunks), per):
                    if any(c.numel() > 0 for _, c in chunks[lo:lo + per]):
                        n += 1
            else:
                n += nonempty
        self.kernel_launches += n

    # -- helpers ---------------------------------------------------------------------------
    @staticmethod
    def _stream() -> int:
        return torch.cuda.current_stream().cuda_stream

    @staticmethod
    def _ptr(t: torch.Tensor | None) -> int:
        return 0 if t is None else t.data_ptr()

    @staticmethod
    def _require_cuda(*tensors):
        for t in tensors:
            if t is not None and not t.is_cuda:
                raise L.FloxB200Error("flox_b200 kernels need CUDA tensors (no CPU fallback)")

    def _chunk_array(self, chunks: Chunks, A: int, B: int):
        cached = getattr(chunks, "table", None)
        if cached is not None and cached[3] == A and cached[4] == B:
            return cached[0], cached[1], cached[2]
        arr, vdt, cdt = self._build_chunk_array(chunks, A, B)
        if isinstance(chunks, PreparedChunks):
            chunks.table = (arr, vdt, cdt, A, B)
        return arr, vdt, cdt

    def _build_chunk_array(self, chunks: Chunks, A: int, B: int):
        arr = (L.Chunk * len(chunks))()
        vdt = cdt = None
        for k, (v, c) in enumerate(chunks):
            self._require_cuda(v, c)
            if not v.is_contiguous() or not c.is_contiguous():
                raise ValueError("chunks must be contiguous")
            n = c.numel()
            if v.numel() != A * n * B:
                raise ValueError(f"values chunk has {v.numel()} elements, expected {A}*{n}*{B}")
            if vdt is None:
                vdt, cdt = v.dtype, c.dtype
            elif (v.dtype, c.dtype) != (vdt, cdt):
                raise TypeError("all chunks must share dtypes")
            arr[k].values, arr[k].codes, arr[k].n = v.data_ptr(), c.data_ptr(), n
        return arr, vdt, cdt

    # -- reductions ------------------------------------------------------------------------
    def moments(self, chunks: Chunks, A: int, B: int, G: int, op: int, acc, cnt, nnan, init=True, force_path=0):
        arr, vdt, cdt = self._chunk_array(chunks, A, B)
        self._require_cuda(acc, cnt, nnan)
        if vdt is None:
            vdt, cdt = (torch.float64 if acc.dtype == torch.float64 else torch.int64), torch.int64
        lib = L.load()
        ws, ws_bytes = None, 0
        if self.use_workspace and (force_path or self.force_path) != 2:
            ws_bytes = int(lib.floxb_moments_workspace_bytes(dtype_code(vdt), A, B, G))
            if ws_bytes:
                ws = torch.empty((ws_bytes + 7) // 8, dtype=torch.float64, device=acc.device)  # caching allocator: no cudaMalloc
        rc = lib.floxb_moments(arr, len(chunks), dtype_code(vdt), dtype_code(cdt), A, B, G, op, self._ptr(acc),
                               self._ptr(cnt), self._ptr(nnan), int(init), force_path or self.force_path,
                               self._ptr(ws), ws_bytes, self._stream())
        L.check(rc, "floxb_moments")
        self.launches += 1
        self._scan_launches(chunks, A, B, G, init, 0, vdt, force_path or self.force_path)

    def minmax(self, chunks: Chunks, A: int, B: int, G: int, is_max: bool, key, cnt, nnan, init=True, force_path=0,
               value_dtype=None):
        arr, vdt, cdt = self._chunk_array(chunks, A, B)
        self._require_cuda(key, cnt, nnan)
        if vdt is None:
            vdt, cdt = value_dtype or torch.float64, torch.int64
        rc = L.load().floxb_minmax(arr, len(chunks), dtype_code(vdt), dtype_code(cdt), A, B, G, int(is_max),
                                   self._ptr(key), self._ptr(cnt), self._ptr(nnan), int(init),
                                   force_path or self.force_path, self._stream())
        L.check(rc, "floxb_minmax")
        self.launches += 1
        self._scan_launches(chunks, A, B, G, init, 1, vdt, force_path or self.force_path)

    def group_index(self, codes, G: int, perm, goff, cursor):
        self._require_cuda(codes, perm, goff, cursor)
        rc = L.load().floxb_group_index(self._ptr(codes), dtype_code(codes.dtype), codes.numel(), G, self._ptr(perm),
                                        self._ptr(goff), self._ptr(cursor), self._stream())
        L.check(rc, "floxb_group_index")
        self.launches += 1
        self.kernel_launches += 1 + (2 if codes.numel() else 0)

    def reduce_cols(self, values, perm, goff, A: int, n: int, B: int, G: int, mode: int, acc, cnt, nnan, init=True,
                    group_major=False):
        self._require_cuda(values, perm, goff, acc, cnt, nnan)
        rc = L.load().floxb_reduce_cols(self._ptr(values), dtype_code(values.dtype), self._ptr(perm), self._ptr(goff), A, n, B,
                                        G, mode, self._ptr(acc), self._ptr(cnt), self._ptr(nnan), int(init), int(group_major),
                                        self._stream())
        L.check(rc, "floxb_reduce_cols")
        self.launches += 1
        self.kernel_launches += 1 if A * B * G else 0

    def sqdev(self, chunks: Chunks, A: int, B: int, G: int, means, skipna: bool, m2, init=True, value_dtype=None):
        arr, vdt, cdt = self._chunk_array(chunks, A, B)
        self._require_cuda(means, m2)
        if vdt is None:
            vdt, cdt = value_dtype or torch.float64, torch.int64
        rc = L.load().floxb_sqdev(arr, len(chunks), dtype_code(vdt), dtype_code(cdt), A, B, G, self._ptr(means),
                                  int(skipna), self._ptr(m2), int(init), self._stream())
        L.check(rc, "floxb_sqdev")
        self.launches += 1
        self._scan_launches(chunks, A, B, G, init, 0, torch.float64, 0)

    def group_means(self, value_dtype, acc, cnt, nnan, skipna: bool, means):
        self._require_cuda(acc, cnt, nnan, means)
        rc = L.load().floxb_group_means(dtype_code(value_dtype), self._ptr(acc), self._ptr(cnt), self._ptr(nnan),
                                        acc.numel(), int(skipna), self._ptr(means), self._stream())
        L.check(rc, "floxb_group_means")
        self.launches += 1
        self.kernel_launches += 1

    def argext(self, chunks: Chunks, A: int, B: int, G: int, key, nnan, skipna: bool, index_base: Sequence[int], idx,
               init=True, value_dtype=None):
        arr, vdt, cdt = self._chunk_array(chunks, A, B)
        self._require_cuda(key, nnan, idx)
        if vdt is None:
            vdt, cdt = value_dtype or torch.float64, torch.int64
        base = (C.c_int64 * max(len(chunks), 1))(*[int(b) for b in index_base])
        rc = L.load().floxb_argext(arr, len(chunks), dtype_code(vdt), dtype_code(cdt), A, B, G, self._ptr(key),
                                   self._ptr(nnan), int(skipna), base, self._ptr(idx), int(init), self._stream())
        L.check(rc, "floxb_argext")
        self.launches += 1
        self._scan_launches(chunks, A, B, G, init, 0, vdt, 0, has_smem_path=False)

    def firstlast(self, chunks: Chunks, A: int, B: int, G: int, is_last: bool, skipna: bool, index_base: Sequence[int], pos,
                  init=True, value_dtype=None):
        arr, vdt, cdt = self._chunk_array(chunks, A, B)
        self._require_cuda(pos)
        if vdt is None:
            vdt, cdt = value_dtype or torch.float64, torch.int64
        base = (C.c_int64 * max(len(chunks), 1))(*[int(b) for b in index_base])
        rc = L.load().floxb_firstlast(arr, len(chunks), dtype_code(vdt), dtype_code(cdt), A, B, G, int(is_last), int(skipna), base,
                                      self._ptr(pos), int(init), self._stream())
        L.check(rc, "floxb_firstlast")
        self.launches += 1
        self._scan_launches(chunks, A, B, G, init, 0, vdt, 0, has_smem_path=False)

    def take(self, chunks: Chunks, A: int, B: int, G: int, pos, index_base: Sequence[int], acc, init=True, value_dtype=None):
        arr, vdt, cdt = self._chunk_array(chunks, A, B)
        self._require_cuda(pos, acc)
        if vdt is None:
            vdt, cdt = value_dtype or torch.float64, torch.int64
        base = (C.c_int64 * max(len(chunks), 1))(*[int(b) for b in index_base])
        rc = L.load().floxb_take(arr, len(chunks), dtype_code(vdt), dtype_code(cdt), A, B, G, self._ptr(pos), base, self._ptr(acc),
                                 int(init), self._stream())
        L.check(rc, "floxb_take")
        self.launches += 1
        self._scan_launches(chunks, A, B, G, init, 0, vdt, 0, has_smem_path=False)

    # -- epilogues -------------------------------------------------------------------------
    def finalize(self, kind: int, value_dtype, acc, acc2, cnt, nnan, skipna: bool, fill, min_count: int, mc_fill,
                 ddof: int, out, group_major_ab: int = 0):
        self._require_cuda(acc, acc2, cnt, nnan, out)

        def split(v):
            if v is None:
                return float("nan"), 0
            fv = float(v)
            try:
                iv = int(v)
            except (ValueError, OverflowError):  # nan / inf
                iv = 0
            return fv, iv

        ff, fi = split(fill)
        mf, mi = split(mc_fill)
        rc = L.load().floxb_finalize(kind, dtype_code(value_dtype), dtype_code(out.dtype), self._ptr(acc), self._ptr(acc2),
                                     self._ptr(cnt), self._ptr(nnan), out.numel(), int(skipna), ff, fi, int(min_count), mf,
                                     mi, int(ddof), self._ptr(out), int(group_major_ab), self._stream())
        L.check(rc, "floxb_finalize")
        self.launches += 1
        self.kernel_launches += 1

    def combine(self, kind: int, stacked, out):
        self._require_cuda(stacked, out)
        nblk = stacked.shape[0]
        rc = L.load().floxb_combine(kind, dtype_code(stacked.dtype), self._ptr(stacked), nblk, out.numel(), self._ptr(out),
                                    self._stream())
        L.check(rc, "floxb_combine")
        self.launches += 1
        self.kernel_launches += 1

    def var_combine(self, m2, s, n, out_m2, out_s, out_n):
        self._require_cuda(m2, s, n, out_m2, out_s, out_n)
        rc = L.load().floxb_var_combine(self._ptr(m2), self._ptr(s), self._ptr(n), m2.shape[0], out_m2.numel(),
                                        self._ptr(out_m2), self._ptr(out_s), self._ptr(out_n), self._stream())
        L.check(rc, "floxb_var_combine")
        self.launches += 1
        self.kernel_launches += 1

    # -- label encoding --------------------------------------------------------------------
    def digitize(self, labels, edges, closed_right: bool, codes):
        self._require_cuda(labels, edges, codes)
        rc = L.load().floxb_digitize(self._ptr(labels), dtype_code(labels.dtype), labels.numel(), self._ptr(edges),
                                     edges.numel(), int(closed_right), self._ptr(codes), self._stream())
        L.check(rc, "floxb_digitize")
        self.launches += 1
        self.kernel_launches += 1

    def ravel_codes(self, c0, c1, size0: int, size1: int, out):
        self._require_cuda(c0, c1, out)
        rc = L.load().floxb_ravel_codes(self._ptr(c0), self._ptr(c1), c0.numel(), size0, size1, self._ptr(out),
                                        self._stream())
        L.check(rc, "floxb_ravel_codes")
        self.launches += 1
        self.kernel_launches += 1

    def label_minmax(self, labels) -> tuple[int, int]:
        """(min, max) of a non-empty integer label tensor (one 16-byte D2H read)."""
        self._require_cuda(labels)
        mm = torch.empty(2, dtype=torch.int64, device=labels.device)
        L.check(L.load().floxb_label_minmax(self._ptr(labels), dtype_code(labels.dtype), labels.numel(), self._ptr(mm), self._stream()),
                "floxb_label_minmax")
        self.launches += 1
        self.kernel_launches += 2
        lo, hi = mm.tolist()
        return int(lo), int(hi)

    def label_presence(self, labels, lo: int, span: int) -> torch.Tensor:
        """u8[span]: 1 where lo + k occurs in `labels`."""
        self._require_cuda(labels)
        table = torch.empty(span, dtype=torch.uint8, device=labels.device)
        L.check(L.load().floxb_label_presence(self._ptr(labels), dtype_code(labels.dtype), labels.numel(), lo, span, self._ptr(table),
                                              self._stream()), "floxb_label_presence")
        self.launches += 1
        self.kernel_launches += 1 if labels.numel() else 0
        return table

    def lut_codes(self, labels, lo: int, lut, codes) -> None:
        self._require_cuda(labels, lut, codes)
        L.check(L.load().floxb_lut_codes(self._ptr(labels), dtype_code(labels.dtype), labels.numel(), lo, lut.numel(), self._ptr(lut),
                                         self._ptr(codes), self._stream()), "floxb_lut_codes")
        self.launches += 1
        self.kernel_launches += 1 if labels.numel() else 0

    # -- sorted-segment family (csrc/segments.cu) ----------------------------------------------------
    def segment_valid(self, sorted_vals, starts, nvalid):
        self._require_cuda(sorted_vals, starts, nvalid)
        R, n = sorted_vals.shape
        L.check(L.load().floxb_segment_valid(self._ptr(sorted_vals), dtype_code(sorted_vals.dtype), R, n, self._ptr(starts),
                                             starts.numel() - 1, self._ptr(nvalid), self._stream()), "floxb_segment_valid")
        self.launches += 1
        self.kernel_launches += 1 if nvalid.numel() else 0

    def segment_quantile(self, sorted_vals, starts, nvalid, before, q, skipna: bool, fill: float, out):
        self._require_cuda(sorted_vals, starts, nvalid, before, q, out)
        R, n = sorted_vals.shape
        L.check(L.load().floxb_segment_quantile(self._ptr(sorted_vals), dtype_code(sorted_vals.dtype), R, n, self._ptr(starts),
                                                starts.numel() - 1, self._ptr(nvalid), self._ptr(before), self._ptr(q),
                                                q.numel(), int(skipna), float(fill), dtype_code(out.dtype), self._ptr(out),
                                                self._stream()), "floxb_segment_quantile")
        self.launches += 1
        self.kernel_launches += 1 if out.numel() else 0

    def segment_scan(self, kind: int, values, perm, sorted_codes, skipna: bool, out):
        self._require_cuda(values, perm, sorted_codes, out)
        R, n = values.shape
        lib = L.load()
        nbytes = int(lib.floxb_segment_scan_scratch_bytes(R, n))
        scratch = torch.empty((nbytes + 7) // 8, dtype=torch.int64, device=values.device)
        L.check(lib.floxb_segment_scan(kind, self._ptr(values), dtype_code(values.dtype), self._ptr(perm), self._ptr(sorted_codes),
                                       R, n, int(skipna), self._ptr(scratch), nbytes, self._ptr(out), self._stream()),
                "floxb_segment_scan")
        self.launches += 1
        self.kernel_launches += 3 if out.numel() else 0

    def smem_group_limit(self, value_dtype, family: int) -> int:
        return int(L.load().floxb_smem_group_limit(dtype_code(value_dtype), family))

    def max_chunks_per_launch(self) -> int:
        return int(L.load().floxb_max_chunks_per_launch())

    # -- label staging (csrc/staging.cu) ------------------------------------------------------
    @staticmethod
    def host_pack_codes(codes_np, n_groups: int, out: torch.Tensor) -> None:
        """HOST function (no CUDA call; releases the GIL): range-check + narrow int32/int64 codes into `out`
        (a pinned int16 / int32 tensor).  Called from worker threads, one chunk each."""
        bits = 16 if out.dtype == torch.int16 else 32
        L.check(L.load().floxb_host_pack_codes(codes_np.ctypes.data, L.I64 if codes_np.dtype.itemsize == 8 else L.I32, codes_np.size, int(n_groups),
                                               out.data_ptr(), bits), "floxb_host_pack_codes")

    def widen_codes(self, codes16: torch.Tensor) -> torch.Tensor:
        out = torch.empty(codes16.shape, dtype=torch.int32, device=codes16.device)
        L.check(L.load().floxb_widen_codes(codes16.data_ptr(), codes16.numel(), out.data_ptr(), self._stream()), "floxb_widen_codes")
        self.launches += 1
        self.kernel_launches += 1 if codes16.numel() else 0
        return out


_backend = None


def get_kernels():
    """The active kernel backend.  CUDA unless a test installed an emulator via set_backend."""
    global _backend
    if _backend is None:
        if not torch.cuda.is_available():
            raise L.FloxB200Error("CUDA is not available and flox_b200 has no CPU fallback")
        L.load()
        _backend = CudaKernels()
    return _backend


def set_backend(backend) -> None:
    """TEST HOOK: install an object implementing the CudaKernels interface (None restores CUDA)."""
    global _backend
    _backend = backend


def default_device() -> torch.device:
    k = get_kernels()
    return torch.device("cuda", torch.cuda.current_device()) if k.name == "cuda" else torch.device("cpu")
