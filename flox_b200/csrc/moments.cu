// Grouped sum / product / sum-of-squares + non-NaN count + NaN count in ONE pass over
// values[A,n,B] and codes[n].   Replaces flox/aggregate_flox.py:133-192 (_np_grouped_op with
// np.add/multiply.reduceat), :195-217 (nan variants), :225-248 (sum_of_squares, nanlen) and the
// argsort + gathers of _prepare_for_flox (:9-23): a scatter reduction needs no sort.
//
// Two paths:
//  * smem   (A == B == 1, G <= limit): one persistent CTA per SM keeps PRIVATE bins in shared
//           memory: an accumulator slot + u32 non-NaN count + u32 NaN count (touched only by NaN
//           rows).  sm_100a has no native floating-point or 64-bit shared-memory add (ATOMS.ADD is
//           32-bit integer only), so the slot is updated with the compiler's LDS + ADD +
//           ATOMS.CAST.SPIN loop, and the kernel is bound by L1TEX shared-memory WAVEFRONTS, not
//           by HBM (ncu, profiles/r01_ncu_moments_f64cas_summary.json: l1tex 98% busy, 24
//           wavefronts per 32 rows with an f64 slot).  For f32 values the slot is therefore an
//           f32 PARTIAL sum (half the wavefronts of a 64-bit slot) that is drained into the f64
//           global accumulator (atomicExch + RED.ADD.F64) every kFlushEvery-th row of that bin,
//           so no f32 partial ever holds more than ~8 addends: worst-case relative error
//           7 * 2^-24 = 4.2e-7 of sum|x| for pathological rounding, ~1e-8 in practice because the
//           ~N/8 partials per group round independently — tighter than the reference's own f32
//           pairwise reduceat.  f64 / integer values keep a 64-bit slot (exact in f64 / i64).
//           Remaining bins are merged with one RED.ADD.F64 / RED.ADD.U64 per non-empty bin per
//           CTA at the end.
//  * global (anything else): every element issues RED.ADD.F64 + RED.ADD.U64 straight to the
//           L2-resident accumulators.
// HBM traffic is the algorithmic minimum: every value and code is read exactly once
// (ld.global.nc.L1::no_allocate, 128-bit), accumulators are written once.
#include "common.cuh"

namespace floxb {

template <typename AccT, int OP> __device__ __forceinline__ void acc_apply_shared(AccT* slot, AccT v) {
  if constexpr (OP == FLOXB_OP_PROD) atomic_mul(slot, v);
  else atomic_add(slot, v);
}

template <typename VT, int OP> __device__ __forceinline__ typename Acc<VT>::type lift(VT v) {
  using AccT = typename Acc<VT>::type;
  AccT w = (AccT)v;
  if constexpr (OP == FLOXB_OP_SUMSQ) return w * w;
  else return w;
}

// Shared-memory slot type per (value type, op): f32 partials for f32 sums / sums of squares.
constexpr unsigned kFlushEvery = 8;  // power of two
template <typename VT, int OP> struct Slot {
  using type = typename Acc<VT>::type;
  static constexpr bool partial = false;
};
template <> struct Slot<float, FLOXB_OP_SUM> { using type = float; static constexpr bool partial = true; };
template <> struct Slot<float, FLOXB_OP_SUMSQ> { using type = float; static constexpr bool partial = true; };

// ------------------------------------------------------------------------------------------
// shared-memory privatised bins
// ------------------------------------------------------------------------------------------
template <typename VT, typename CT, int OP, int THREADS, int ITERS>
__global__ void __launch_bounds__(THREADS, 1)
k_moments_smem(const __grid_constant__ ChunkTable tab, int64_t total_tiles, int G,
               typename Acc<VT>::type* __restrict__ gacc, unsigned long long* __restrict__ gcnt,
               unsigned long long* __restrict__ gnan) {
  using AccT = typename Acc<VT>::type;
  constexpr int TILE = THREADS * 4 * ITERS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using ST = typename Slot<VT, OP>::type;
  constexpr bool kPartial = Slot<VT, OP>::partial;
  ST* sacc = reinterpret_cast<ST*>(smem_raw);
  unsigned* scnt = reinterpret_cast<unsigned*>(sacc + G);
  unsigned* snan = scnt + G;  // allocated only for float VT

  const ST ident = (OP == FLOXB_OP_PROD) ? (ST)1 : (ST)0;
  for (int g = threadIdx.x; g < G; g += THREADS) {
    sacc[g] = ident;
    scnt[g] = 0u;
    if constexpr (is_float<VT>::value) snan[g] = 0u;
  }
  __syncthreads();

  auto accum = [&](VT v, CT c) {
    if ((unsigned long long)c < (unsigned long long)G) {
      if (is_nan(v)) {
        if constexpr (is_float<VT>::value) atomicAdd(&snan[(int)c], 1u);
      } else if constexpr (kPartial) {
        const float w = (OP == FLOXB_OP_SUMSQ) ? v * v : v;
        atomicAdd(&sacc[(int)c], w);
        const unsigned seen = atomicAdd(&scnt[(int)c], 1u);
        if ((seen & (kFlushEvery - 1)) == kFlushEvery - 1) {
          // drain this bin's f32 partial into the f64 accumulator; atomicExch hands the partial to
          // exactly one thread, concurrent adds land in the fresh zero
          const float part = atomicExch(&sacc[(int)c], 0.0f);
          atomic_add(&gacc[(int)c], (AccT)part);
        }
      } else {
        acc_apply_shared<AccT, OP>(&sacc[(int)c], lift<VT, OP>(v));
        atomicAdd(&scnt[(int)c], 1u);
      }
    }
  };

  for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int ch = find_chunk(tab, tile);
    const int64_t base = (tile - tab.tile_prefix[ch]) * TILE;
    const int64_t n = tab.n[ch];
    const VT* __restrict__ vals = reinterpret_cast<const VT*>(tab.values[ch]) + base;
    const CT* __restrict__ codes = reinterpret_cast<const CT*>(tab.codes[ch]) + base;
    const int64_t rows = (n - base < TILE) ? (n - base) : TILE;
    const bool aligned = ((reinterpret_cast<uintptr_t>(vals) | reinterpret_cast<uintptr_t>(codes)) & 15) == 0;
    if (rows == TILE && aligned) {
      Vec4<VT> v[ITERS];
      Vec4<CT> c[ITERS];
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const int off = (it * THREADS + threadIdx.x) * 4;
        v[it] = load4<VT>(vals + off);
        c[it] = load4<CT>(codes + off);
      }
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) accum(v[it].v[k], c[it].v[k]);
      }
    } else {
      for (int64_t i = threadIdx.x; i < rows; i += THREADS) accum(vals[i], codes[i]);
    }
  }
  __syncthreads();

  for (int g = threadIdx.x; g < G; g += THREADS) {
    const unsigned c = scnt[g];
    if (c) {
      if constexpr (OP == FLOXB_OP_PROD) atomic_mul(&gacc[g], (AccT)sacc[g]);
      else atomic_add(&gacc[g], (AccT)sacc[g]);
      atomicAdd(&gcnt[g], (unsigned long long)c);
    }
    if constexpr (is_float<VT>::value) {
      const unsigned m = snan[g];
      if (m) atomicAdd(&gnan[g], (unsigned long long)m);
    }
  }
}

// ------------------------------------------------------------------------------------------
// L2-atomic path: any A, B, G
// ------------------------------------------------------------------------------------------
template <typename VT, typename CT, int OP>
__global__ void __launch_bounds__(256)
k_moments_global(const VT* __restrict__ vals, const CT* __restrict__ codes, int64_t A, int64_t n, int64_t B,
                 int64_t G, typename Acc<VT>::type* __restrict__ gacc, unsigned long long* __restrict__ gcnt,
                 unsigned long long* __restrict__ gnan) {
  using AccT = typename Acc<VT>::type;
  const int64_t row_len = n * B;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t a = blockIdx.y; a < A; a += gridDim.y) {
    const VT* __restrict__ row = vals + a * row_len;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < row_len; j += stride) {
      int64_t i, b;
      if (B == 1) { i = j; b = 0; } else { i = j / B; b = j - i * B; }
      const CT c = codes[i];
      if ((unsigned long long)c >= (unsigned long long)G) continue;
      const VT v = row[j];
      const int64_t bin = (a * B + b) * G + (int64_t)c;
      if (is_nan(v)) {
        atomicAdd(&gnan[bin], 1ull);
      } else {
        if constexpr (OP == FLOXB_OP_PROD) atomic_mul(&gacc[bin], lift<VT, OP>(v));
        else atomic_add(&gacc[bin], lift<VT, OP>(v));
        atomicAdd(&gcnt[bin], 1ull);
      }
    }
  }
}

// L2-atomic path specialised for B == 1: 4 rows per thread (128-bit loads), adjacent equal codes are
// merged in registers, and a warp whose 128 rows all carry ONE code (sorted labels: BASELINE config 5,
// method="cohorts") reduces with shuffles and issues a single set of atomics.  For sorted labels the
// atomic traffic drops by ~128x and the kernel is a pure HBM stream; for random labels it degrades to
// one atomic set per row like k_moments_global.
template <typename VT, typename CT, int OP>
__global__ void __launch_bounds__(256)
k_moments_global_runs(const VT* __restrict__ vals, const CT* __restrict__ codes, int64_t A, int64_t n, int64_t G,
                      typename Acc<VT>::type* __restrict__ gacc, unsigned long long* __restrict__ gcnt,
                      unsigned long long* __restrict__ gnan) {
  using AccT = typename Acc<VT>::type;
  const AccT ident = (OP == FLOXB_OP_PROD) ? (AccT)1 : (AccT)0;
  auto merge = [](AccT a, AccT b) -> AccT { if constexpr (OP == FLOXB_OP_PROD) return a * b; else return a + b; };
  const int64_t nvec = n / 4;
  for (int64_t a = blockIdx.y; a < A; a += gridDim.y) {
    const VT* __restrict__ row = vals + a * n;
    AccT* __restrict__ racc = gacc + a * G;
    unsigned long long* __restrict__ rcnt = gcnt + a * G;
    unsigned long long* __restrict__ rnan = gnan + a * G;
    auto flush = [&](long long code, AccT s, unsigned c, unsigned m) {
      if ((unsigned long long)code >= (unsigned long long)G) return;
      if (c) {
        if constexpr (OP == FLOXB_OP_PROD) atomic_mul(&racc[code], s); else atomic_add(&racc[code], s);
        atomicAdd(&rcnt[code], (unsigned long long)c);
      }
      if (m) atomicAdd(&rnan[code], (unsigned long long)m);
    };
    const bool aligned = ((reinterpret_cast<uintptr_t>(row) | reinterpret_cast<uintptr_t>(codes)) & 15) == 0;
    // whole warps only in the vector loop so that the shuffles below are full-mask
    const int64_t warp_span = 32;
    const int64_t nvec_full = aligned ? (nvec / warp_span) * warp_span : 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nvec_full; q += stride) {
      const Vec4<VT> v = load4<VT>(row + q * 4);
      const Vec4<CT> c = load4<CT>(codes + q * 4);
      long long cur = (long long)c.v[0];
      AccT s = ident;
      unsigned cn = 0, mn = 0;
      bool single = true;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const long long ck = (long long)c.v[k];
        if (ck != cur) {
          flush(cur, s, cn, mn);
          cur = ck; s = ident; cn = 0; mn = 0; single = false;
        }
        if (is_nan(v.v[k])) mn++; else { s = merge(s, lift<VT, OP>(v.v[k])); cn++; }
      }
      int uniform;
      __match_all_sync(0xffffffffu, cur, &uniform);
      const bool warp_one = __all_sync(0xffffffffu, single) && uniform;
      if (warp_one) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
          if constexpr (sizeof(AccT) == 8) {
            const long long o = __shfl_xor_sync(0xffffffffu, *reinterpret_cast<long long*>(&s), d);
            s = merge(s, *reinterpret_cast<const AccT*>(&o));
          }
          cn += __shfl_xor_sync(0xffffffffu, cn, d);
          mn += __shfl_xor_sync(0xffffffffu, mn, d);
        }
        if ((threadIdx.x & 31) == 0) flush(cur, s, cn, mn);
      } else {
        flush(cur, s, cn, mn);
      }
    }
    // scalar remainder (unaligned chunks, n % 128 tail)
    for (int64_t i = nvec_full * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const VT v = row[i];
      if (is_nan(v)) flush((long long)codes[i], ident, 0u, 1u);
      else flush((long long)codes[i], lift<VT, OP>(v), 1u, 0u);
    }
  }
}

template <typename AccT>
__global__ void k_fill_acc(AccT* acc, unsigned long long* cnt, unsigned long long* nnan, int64_t nbins, AccT ident) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbins; i += (int64_t)gridDim.x * blockDim.x) {
    acc[i] = ident;
    if (cnt) cnt[i] = 0ull;
    if (nnan) nnan[i] = 0ull;
  }
}

// ------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------
constexpr int kSmemThreads = 1024;
constexpr int kSmemIters = 2;
constexpr size_t kSmemBudget = 227 * 1024;

template <typename VT, int OP> size_t smem_bytes_moments(int64_t G) {
  return (size_t)G * (sizeof(typename Slot<VT, OP>::type) + 4 + (is_float<VT>::value ? 4 : 0));
}

template <typename VT, typename CT, int OP>
int launch_moments(const floxb_chunk_t* chunks, int64_t n_chunks, int64_t A, int64_t B, int64_t G, void* acc,
                   int64_t* cnt, int64_t* nnan, int init, int force_path, cudaStream_t st) {
  using AccT = typename Acc<VT>::type;
  const int64_t nbins = A * B * G;
  AccT* gacc = reinterpret_cast<AccT*>(acc);
  auto* gcnt = reinterpret_cast<unsigned long long*>(cnt);
  auto* gnan = reinterpret_cast<unsigned long long*>(nnan);
  if (init && nbins > 0) {
    const int blocks = (int)((nbins + 255) / 256 < 1184 ? (nbins + 255) / 256 : 1184);
    k_fill_acc<AccT><<<blocks, 256, 0, st>>>(gacc, gcnt, gnan, nbins, (OP == FLOXB_OP_PROD) ? (AccT)1 : (AccT)0);
  }
  if (nbins == 0) return FLOXB_OK;
  const size_t smem = smem_bytes_moments<VT, OP>(G);
  bool use_smem = (A == 1 && B == 1 && smem <= kSmemBudget);
  if (force_path == 1 && !use_smem) return FLOXB_EINVAL;
  if (force_path == 2) use_smem = false;

  if (use_smem) {
    constexpr int TILE = kSmemThreads * 4 * kSmemIters;
    auto kern = k_moments_smem<VT, CT, OP, kSmemThreads, kSmemIters>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget) != cudaSuccess)
      return FLOXB_ECUDA;
    for (int64_t c0 = 0; c0 < n_chunks; c0 += kMaxChunksPerLaunch) {
      ChunkTable tab;
      const int m = (int)((n_chunks - c0 < kMaxChunksPerLaunch) ? (n_chunks - c0) : kMaxChunksPerLaunch);
      tab.nchunks = m;
      tab.tile_prefix[0] = 0;
      for (int k = 0; k < m; ++k) {
        const floxb_chunk_t& ch = chunks[c0 + k];
        if (ch.n < 0 || (ch.n > 0 && (!ch.values || !ch.codes))) return FLOXB_EINVAL;
        tab.values[k] = ch.values;
        tab.codes[k] = ch.codes;
        tab.n[k] = ch.n;
        tab.tile_prefix[k + 1] = tab.tile_prefix[k] + (ch.n + TILE - 1) / TILE;
      }
      const int64_t tiles = tab.tile_prefix[m];
      if (tiles == 0) continue;
      // u32 per-CTA counters: keep rows per CTA below 2^32
      if (tiles / sm_count() * TILE >= (int64_t)1 << 32) return FLOXB_ERANGE;
      const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
      kern<<<grid, kSmemThreads, smem, st>>>(tab, tiles, (int)G, gacc, gcnt, gnan);
    }
  } else {
    for (int64_t k = 0; k < n_chunks; ++k) {
      const floxb_chunk_t& ch = chunks[k];
      if (ch.n < 0 || (ch.n > 0 && (!ch.values || !ch.codes))) return FLOXB_EINVAL;
      if (ch.n == 0) continue;
      const int64_t row_len = ch.n * B;
      int64_t bx = (row_len + 255) / 256;
      const int64_t cap = (int64_t)sm_count() * 16;
      if (bx > cap) bx = cap;
      const int by = (int)(A < 65535 ? A : 65535);
      if (bx * by > cap * 4) bx = (cap * 4 + by - 1) / by;
      if (bx < 1) bx = 1;
      dim3 grid((unsigned)bx, (unsigned)by);
      if (B == 1) {
        k_moments_global_runs<VT, CT, OP><<<grid, 256, 0, st>>>(reinterpret_cast<const VT*>(ch.values),
                                                              reinterpret_cast<const CT*>(ch.codes), A, ch.n, G, gacc,
                                                              gcnt, gnan);
      } else {
        k_moments_global<VT, CT, OP><<<grid, 256, 0, st>>>(reinterpret_cast<const VT*>(ch.values),
                                                         reinterpret_cast<const CT*>(ch.codes), A, ch.n, B, G, gacc,
                                                         gcnt, gnan);
      }
    }
  }
  return cudaGetLastError() == cudaSuccess ? FLOXB_OK : FLOXB_ECUDA;
}

template <typename VT, typename CT>
int dispatch_op(int op, const floxb_chunk_t* chunks, int64_t n_chunks, int64_t A, int64_t B, int64_t G, void* acc,
                int64_t* cnt, int64_t* nnan, int init, int force_path, cudaStream_t st) {
  switch (op) {
    case FLOXB_OP_SUM: return launch_moments<VT, CT, FLOXB_OP_SUM>(chunks, n_chunks, A, B, G, acc, cnt, nnan, init, force_path, st);
    case FLOXB_OP_PROD: return launch_moments<VT, CT, FLOXB_OP_PROD>(chunks, n_chunks, A, B, G, acc, cnt, nnan, init, force_path, st);
    case FLOXB_OP_SUMSQ: return launch_moments<VT, CT, FLOXB_OP_SUMSQ>(chunks, n_chunks, A, B, G, acc, cnt, nnan, init, force_path, st);
  }
  return FLOXB_EINVAL;
}

template <typename VT>
int dispatch_code(int code_dtype, int op, const floxb_chunk_t* chunks, int64_t n_chunks, int64_t A, int64_t B,
                  int64_t G, void* acc, int64_t* cnt, int64_t* nnan, int init, int force_path, cudaStream_t st) {
  if (code_dtype == FLOXB_I64) return dispatch_op<VT, int64_t>(op, chunks, n_chunks, A, B, G, acc, cnt, nnan, init, force_path, st);
  if (code_dtype == FLOXB_I32) return dispatch_op<VT, int32_t>(op, chunks, n_chunks, A, B, G, acc, cnt, nnan, init, force_path, st);
  return FLOXB_EDTYPE;
}

}  // namespace floxb

using namespace floxb;

extern "C" int floxb_moments(const floxb_chunk_t* chunks, int64_t n_chunks, int value_dtype, int code_dtype, int64_t A,
                             int64_t B, int64_t n_groups, int op, void* acc, int64_t* cnt, int64_t* nnan, int init,
                             int force_path, void* stream) {
  if (n_chunks < 0 || A < 0 || B < 0 || n_groups < 0 || (n_chunks > 0 && !chunks)) return FLOXB_EINVAL;
  if (A * B * n_groups > 0 && (!acc || !cnt || !nnan)) return FLOXB_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (value_dtype) {
    case FLOXB_F32: return dispatch_code<float>(code_dtype, op, chunks, n_chunks, A, B, n_groups, acc, cnt, nnan, init, force_path, st);
    case FLOXB_F64: return dispatch_code<double>(code_dtype, op, chunks, n_chunks, A, B, n_groups, acc, cnt, nnan, init, force_path, st);
    case FLOXB_I32: return dispatch_code<int32_t>(code_dtype, op, chunks, n_chunks, A, B, n_groups, acc, cnt, nnan, init, force_path, st);
    case FLOXB_I64: return dispatch_code<int64_t>(code_dtype, op, chunks, n_chunks, A, B, n_groups, acc, cnt, nnan, init, force_path, st);
  }
  return FLOXB_EDTYPE;
}

extern "C" int64_t floxb_smem_group_limit(int value_dtype, int family) {
  // family 0 reports the PROD / f64 worst case (8-byte slot); f32 sums use 4-byte slots and fit more
  const size_t per = (family == 0) ? (8 + 4 + ((value_dtype == FLOXB_F32 || value_dtype == FLOXB_F64) ? 4 : 0))
                                   : (((value_dtype == FLOXB_F32 || value_dtype == FLOXB_I32) ? 4 : 8) + 4 + 4);
  return (int64_t)(kSmemBudget / per);
}
