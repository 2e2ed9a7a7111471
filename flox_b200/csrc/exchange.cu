// flox_b200 — cross-GPU merge of the per-rank accumulators over NVLink peer memory (one box, one process per GPU).
//
// flox's map-reduce has one exchange step: the tree of `_simple_combine` nodes over the per-chunk intermediates
// (`_tree_reduce(combine=_simple_combine)`, flox/dask.py:440-460; dask_array_ops.py:29-121) — with dask the scheduler
// ships python dicts between workers.  Here every rank has already folded ITS chunks into one `[A, B, G]` accumulator
// set, so the combine is an all-reduce of at most a few hundred kilobytes: pure latency.  NCCL costs ~25 us of host time
// and ~15 us of device time per call and the step needs two of them (f64 sums, i64 counts); at 8 GPUs the chunk kernel
// itself is 0.3 ms.  This file replaces them by ONE launch:
//
//   * every rank keeps its accumulators in a cudaMalloc'ed SLAB that all peers map through CUDA IPC
//     (floxb_ipc_alloc / floxb_ipc_open);
//   * `k_xchg` first runs a flag barrier over peer memory (a release store of the collective's epoch into every peer's
//     flag row, an acquire spin on the local row), then every thread reads the SAME word from all W slabs — its own and
//     W-1 peers over NVLink — combines them in rank order (so all ranks get bit-identical results) and writes the merged
//     word to a local output buffer;
//   * slabs are double-buffered by the caller: collective s uses buffer s % 2.  A rank can only start refilling buffer
//     s % 2 for collective s + 2 after its own k_xchg(s + 1) finished, which waited for every peer's epoch s + 1 signal,
//     which each peer sends after it finished READING in k_xchg(s).  One barrier per collective is therefore enough.
//
// A peer that never arrives must not hang the GPU for good: the spin gives up after about a minute of clock64() and raises
// a sticky error word in the local slab header that the host reads back (floxb_xchg_status).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/flox_b200.h"

namespace floxb {

// Ranks of one job may legitimately be seconds apart (one is still loading or generating its chunks) — NCCL would simply
// wait.  The spin therefore waits about a minute of SM clock before it declares the peer dead.
constexpr long long kBarrierTimeoutCycles = 120000000000LL;  // ~63 s at 1.9 GHz
constexpr int kMaxPeers = FLOXB_XCHG_MAX_RANKS;
constexpr int kMaxSegs = FLOXB_XCHG_MAX_SEGMENTS;

struct XchgArgs {
  unsigned char* base[kMaxPeers];  // every rank's slab base (this rank's own mapping for `rank`)
  int world, rank;
  unsigned long long epoch;
  int64_t data_offset;             // byte offset of the buffer in use inside each slab (header excluded)
  int nseg;
  int64_t seg_words[kMaxSegs];     // 8-byte words per segment
  int seg_op[kMaxSegs];            // FLOXB_XCHG_*
  void* out[kMaxSegs];             // local destinations
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_peer(const unsigned long long* p) {
  unsigned long long v;  // peer memory: never through the non-coherent L1
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// slab header: flags[kMaxPeers] (epoch of the last collective each peer entered), then the sticky error word
constexpr int64_t kHeaderBytes = FLOXB_XCHG_HEADER_BYTES;

__global__ void __launch_bounds__(256)
k_xchg(const __grid_constant__ XchgArgs a) {
  unsigned long long* my_flags = reinterpret_cast<unsigned long long*>(a.base[a.rank]);
  unsigned long long* my_err = my_flags + kMaxPeers;
  // ---- barrier: tell every peer this rank's buffer is complete (the kernels that filled it precede this launch in
  // stream order), then wait until every peer said the same
  if (blockIdx.x == 0 && threadIdx.x < a.world) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<unsigned long long*>(a.base[threadIdx.x]) + a.rank, a.epoch);
  }
  __shared__ int s_ok;
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if (threadIdx.x < a.world) {
    const long long t0 = clock64();
    while (ld_acquire_sys(my_flags + threadIdx.x) < a.epoch) {
      if (clock64() - t0 > kBarrierTimeoutCycles) {  // a peer died: do not hang the device for good
        s_ok = 0;
        atomicExch(my_err, a.epoch);
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  if (!s_ok) return;
  // ---- merge: word j of the buffer from every rank, in rank order
  int64_t first = 0;
  for (int s = 0; s < a.nseg; ++s) {
    const int64_t nw = a.seg_words[s];
    const int op = a.seg_op[s];
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nw; j += (int64_t)gridDim.x * blockDim.x) {
      const int64_t off = a.data_offset + (first + j) * 8;
      unsigned long long w[kMaxPeers];
#pragma unroll
      for (int r = 0; r < kMaxPeers; ++r)
        if (r < a.world) w[r] = ld_peer(reinterpret_cast<const unsigned long long*>(a.base[r] + off));
      unsigned long long res = w[0];
#pragma unroll
      for (int r = 1; r < kMaxPeers; ++r) {
        if (r >= a.world) break;
        switch (op) {
          case FLOXB_XCHG_SUM_F64: res = (unsigned long long)__double_as_longlong(__longlong_as_double((long long)res) + __longlong_as_double((long long)w[r])); break;
          case FLOXB_XCHG_SUM_I64: res += w[r]; break;
          case FLOXB_XCHG_MAX_I64: res = (unsigned long long)max((long long)res, (long long)w[r]); break;
          case FLOXB_XCHG_MIN_I64: res = (unsigned long long)min((long long)res, (long long)w[r]); break;
          case FLOXB_XCHG_PROD_F64: res = (unsigned long long)__double_as_longlong(__longlong_as_double((long long)res) * __longlong_as_double((long long)w[r])); break;
          case FLOXB_XCHG_PROD_I64: res = (unsigned long long)((long long)res * (long long)w[r]); break;
        }
      }
      reinterpret_cast<unsigned long long*>(a.out[s])[j] = res;
    }
    first += nw;
  }
}

}  // namespace floxb

using namespace floxb;

extern "C" int floxb_ipc_alloc(int64_t bytes, void** dptr, void* handle64) {
  if (bytes <= 0 || !dptr || !handle64) return FLOXB_EINVAL;
  void* p = nullptr;
  if (cudaMalloc(&p, (size_t)bytes) != cudaSuccess) return FLOXB_ECUDA;
  if (cudaMemset(p, 0, (size_t)bytes) != cudaSuccess) return FLOXB_ECUDA;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), p) != cudaSuccess) {
    cudaFree(p);
    return FLOXB_ECUDA;
  }
  if (cudaDeviceSynchronize() != cudaSuccess) return FLOXB_ECUDA;
  *dptr = p;
  return FLOXB_OK;
}

extern "C" int floxb_ipc_open(const void* handle64, void** dptr) {
  if (!handle64 || !dptr) return FLOXB_EINVAL;
  cudaIpcMemHandle_t h;
  for (int i = 0; i < 64; ++i) reinterpret_cast<unsigned char*>(&h)[i] = reinterpret_cast<const unsigned char*>(handle64)[i];
  void* p = nullptr;
  if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) return FLOXB_ECUDA;
  *dptr = p;
  return FLOXB_OK;
}

extern "C" int floxb_ipc_close(void* dptr) {
  if (!dptr) return FLOXB_OK;
  return cudaIpcCloseMemHandle(dptr) == cudaSuccess ? FLOXB_OK : FLOXB_ECUDA;
}

extern "C" int floxb_ipc_free(void* dptr) {
  if (!dptr) return FLOXB_OK;
  return cudaFree(dptr) == cudaSuccess ? FLOXB_OK : FLOXB_ECUDA;
}

extern "C" int floxb_xchg_allreduce(void* const* slab_bases, int world, int rank, uint64_t epoch, int64_t data_offset,
                                    int n_segments, const int64_t* seg_words, const int* seg_ops, void* const* outs,
                                    void* stream) {
  if (!slab_bases || world < 1 || world > kMaxPeers || rank < 0 || rank >= world || n_segments < 0 ||
      n_segments > kMaxSegs || data_offset < kHeaderBytes || (data_offset & 7))
    return FLOXB_EINVAL;
  XchgArgs a{};
  a.world = world;
  a.rank = rank;
  a.epoch = epoch;
  a.data_offset = data_offset;
  a.nseg = n_segments;
  int64_t total = 0;
  for (int r = 0; r < world; ++r) {
    if (!slab_bases[r]) return FLOXB_EINVAL;
    a.base[r] = reinterpret_cast<unsigned char*>(slab_bases[r]);
  }
  for (int s = 0; s < n_segments; ++s) {
    if (seg_words[s] < 0 || (seg_words[s] > 0 && !outs[s])) return FLOXB_EINVAL;
    if (seg_ops[s] < FLOXB_XCHG_SUM_F64 || seg_ops[s] > FLOXB_XCHG_PROD_I64) return FLOXB_EINVAL;
    a.seg_words[s] = seg_words[s];
    a.seg_op[s] = seg_ops[s];
    a.out[s] = outs[s];
    total += seg_words[s];
  }
  int64_t blocks = (total + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 64) blocks = 64;  // latency-bound: a few CTAs per segment are plenty and keep the barrier cheap
  k_xchg<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  return cudaGetLastError() == cudaSuccess ? FLOXB_OK : FLOXB_ECUDA;
}

extern "C" int floxb_xchg_status(const void* slab_base, uint64_t* failed_epoch) {
  if (!slab_base || !failed_epoch) return FLOXB_EINVAL;
  unsigned long long v = 0;
  if (cudaMemcpy(&v, reinterpret_cast<const unsigned long long*>(slab_base) + kMaxPeers, 8, cudaMemcpyDeviceToHost) != cudaSuccess)
    return FLOXB_ECUDA;
  *failed_epoch = v;
  return FLOXB_OK;
}
