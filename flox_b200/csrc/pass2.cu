// Second-pass kernels that need a per-group value from a first pass:
//   * floxb_group_means + floxb_sqdev : var_chunk's sum of squared deviations about the group
//     mean (flox/aggregations.py:348-389) — a numerically centred two-pass variance, exactly
//     the reference's formulation (one-pass sum-of-squares would lose the tests' 1e12 offsets,
//     tests/test_core.py:2287-2319).
//   * floxb_argext : first position of the extremum per group (chunk_argreduce,
//     flox/core.py:157-211; numpy first-occurrence semantics pinned by tests/test_core.py:388-408).
#include <type_traits>

#include "common.cuh"

namespace floxb {

template <typename VT>
__global__ void k_group_means(const typename Acc<VT>::type* __restrict__ acc, const long long* __restrict__ cnt,
                              const long long* __restrict__ nnan, int64_t nbins, int skipna, double* __restrict__ means) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbins; i += (int64_t)gridDim.x * blockDim.x) {
    double m;
    if (!skipna && nnan[i] > 0) {
      m = __longlong_as_double(0x7ff8000000000000LL);
    } else if constexpr (is_float<VT>::value) {
      // the reference divides in the array dtype: array_sums / array_lens (aggregations.py:375-376)
      const VT s = (VT)acc[i];
      m = (double)(VT)((double)s / (double)cnt[i]);
    } else {
      m = (double)acc[i] / (double)cnt[i];
    }
    means[i] = m;
  }
}

// ---- squared deviations -----------------------------------------------------------------
template <typename VT, typename CT, int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
k_sqdev_smem(const __grid_constant__ ChunkTable tab, int G, const double* __restrict__ means,
             int skipna, double* __restrict__ gm2) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // f32 values: the group mean is f32-representable by construction (floxb_group_means rounds it
  // like the reference's f32 `array_sums / array_lens`), the deviation is formed exactly in f64,
  // and d^2 goes into an f32 PARTIAL that is drained into the f64 accumulator every 8th addend —
  // the same shared-memory-wavefront argument as k_moments_smem (moments.cu).  Other value types
  // keep f64 slots.
  constexpr bool kPartial = sizeof(VT) == 4 && is_float<VT>::value;
  using ST = typename std::conditional<kPartial, float, double>::type;
  ST* sm2 = reinterpret_cast<ST*>(smem_raw);
  ST* smean = sm2 + G;
  unsigned* sseen = reinterpret_cast<unsigned*>(smean + G);  // kPartial only
  for (int g = threadIdx.x; g < G; g += THREADS) {
    sm2[g] = (ST)0;
    smean[g] = (ST)means[g];
    if constexpr (kPartial) sseen[g] = 0u;
  }
  __syncthreads();
  auto accum = [&](VT v, CT c) {
    if ((unsigned long long)c < (unsigned long long)G) {
      if (is_nan(v) && skipna) return;
      const double d = (double)v - (double)smean[(int)c];
      if constexpr (kPartial) {
        atomicAdd(&sm2[(int)c], (float)(d * d));
        const unsigned seen = atomicAdd(&sseen[(int)c], 1u);
        if ((seen & 7u) == 7u) atomicAdd(&gm2[(int)c], (double)atomicExch(&sm2[(int)c], 0.0f));
      } else {
        atomicAdd(&sm2[(int)c], d * d);
      }
    }
  };
  walk_tiles<VT, CT, THREADS>(tab, accum, [] {});
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += THREADS) {
    const double s = (double)sm2[g];
    if (s != 0.0) atomicAdd(&gm2[g], s);  // NaN != 0 is true: NaN propagates
  }
}

template <typename VT, typename CT>
__global__ void __launch_bounds__(256)
k_sqdev_global(const VT* __restrict__ vals, const CT* __restrict__ codes, int64_t A, int64_t n, int64_t B, int64_t G,
               const double* __restrict__ means, int skipna, double* __restrict__ gm2) {
  const int64_t row_len = n * B;
  constexpr int U = 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t a = blockIdx.y; a < A; a += gridDim.y) {
    const VT* __restrict__ row = vals + a * row_len;
    // three phases per step so that U value loads, then U mean loads, are in flight together (see walk_flat)
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t i = j / B, b = j - (j / B) * B;
    const int64_t di = stride / B, db = stride - (stride / B) * B;
    while (j < row_len) {
      VT v[U];
      int64_t bin[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool ok = j < row_len;
        const CT c = ok ? codes[i] : (CT)-1;
        v[u] = ok ? row[j] : VT(0);
        const bool in = (unsigned long long)c < (unsigned long long)G;
        bin[u] = !in ? -1 : (B == 1) ? a * G + (int64_t)c : (int64_t)c * (A * B) + a * B + b;  // group-major when B > 1
        j += stride;
        i += di;
        b += db;
        if (b >= B) { b -= B; ++i; }
      }
      double m[U];
#pragma unroll
      for (int u = 0; u < U; ++u) m[u] = (bin[u] >= 0 && !(is_nan(v[u]) && skipna)) ? means[bin[u]] : 0.0;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (bin[u] < 0 || (is_nan(v[u]) && skipna)) continue;
        const double d = (double)v[u] - m[u];
        atomicAdd(&gm2[bin[u]], d * d);
      }
    }
  }
}

// Small power-of-two B (config 4): the thread keeps its column's running sum of squared deviations in a register
// while the code does not change and issues ONE RED per run (walk_rows_small_b; see k_moments_global_bruns).
template <typename VT, typename CT, int B>
__global__ void __launch_bounds__(256)
k_sqdev_global_bruns(const VT* __restrict__ vals, const CT* __restrict__ codes, int64_t A, int64_t n, int64_t G,
                     const double* __restrict__ means, int skipna, double* __restrict__ gm2) {
  const int col = threadIdx.x & (B - 1);
  for (int64_t a = blockIdx.y; a < A; a += gridDim.y) {
    const int64_t ab = a * B + col, AB = A * B;
    long long cur = -1;
    double s = 0.0, mean = 0.0;
    bool any = false;
    auto flush = [&] {
      if (any) atomicAdd(&gm2[(int64_t)cur * AB + ab], s);
      s = 0.0;
      any = false;
    };
    walk_rows_small_b<B, 8, 64>(vals + a * n * B, codes, n,
      [&](VT v, CT c, int64_t) {
        if ((long long)c != cur) {
          flush();
          cur = (long long)c;
          if ((unsigned long long)cur < (unsigned long long)G) mean = means[(int64_t)cur * AB + ab];
        }
        if ((unsigned long long)cur >= (unsigned long long)G || (is_nan(v) && skipna)) return;
        const double d = (double)v - mean;
        s += d * d;
        any = true;
      },
      [&] { flush(); cur = -1; });
  }
}

// ---- arg-extremum ------------------------------------------------------------------------
template <typename VT, typename CT>
__global__ void __launch_bounds__(256)
k_argext_global(const VT* __restrict__ vals, const CT* __restrict__ codes, int64_t A, int64_t n, int64_t B, int64_t G,
                const long long* __restrict__ key, const long long* __restrict__ nnan, int skipna, int64_t index_base,
                long long* __restrict__ gidx) {
  const int64_t row_len = n * B;
  for (int64_t a = blockIdx.y; a < A; a += gridDim.y) {
    walk_flat<4>(vals + a * row_len, codes, row_len, B, G, [&](VT v, CT c, int64_t i, int64_t b) {
      const int64_t bin = (B == 1) ? a * G + (int64_t)c : (int64_t)c * (A * B) + a * B + b;  // group-major when B > 1
      bool hit;
      if (!skipna && nnan[bin] > 0) hit = is_nan(v);          // NaN propagates: first NaN wins
      else hit = !is_nan(v) && (long long)Key<VT>::enc(v) == key[bin];
      if (hit) atomicMin(&gidx[bin], (long long)(index_base + i));
    });
  }
}

// ---- first / last -------------------------------------------------------------------------
// Position of the first (last) row of every group — of the first (last) NON-NaN row when skipna — as a global index
// along the reduced axis, then a second pass in which the one row whose position won writes its value.  Behind
// first / last / nanfirst / nanlast (flox/aggregations.py:634-651; numpy_groupies' "first"/"last", xrutils.nanfirst /
// nanlast flox/xrutils.py:386-412 for the combine).
template <typename VT, typename CT>
__global__ void __launch_bounds__(256)
k_firstlast_pos(const VT* __restrict__ vals, const CT* __restrict__ codes, int64_t A, int64_t n, int64_t B, int64_t G,
                int is_last, int skipna, int64_t index_base, long long* __restrict__ gpos) {
  const int64_t row_len = n * B;
  for (int64_t a = blockIdx.y; a < A; a += gridDim.y) {
    walk_flat<4>(vals + a * row_len, codes, row_len, B, G, [&](VT v, CT c, int64_t i, int64_t b) {
      if (skipna && is_nan(v)) return;
      const int64_t bin = (B == 1) ? a * G + (int64_t)c : (int64_t)c * (A * B) + a * B + b;
      if (is_last) atomicMax(&gpos[bin], (long long)(index_base + i));
      else atomicMin(&gpos[bin], (long long)(index_base + i));
    });
  }
}

template <typename VT, typename CT>
__global__ void __launch_bounds__(256)
k_take_pos(const VT* __restrict__ vals, const CT* __restrict__ codes, int64_t A, int64_t n, int64_t B, int64_t G,
           const long long* __restrict__ gpos, int64_t index_base, typename Acc<VT>::type* __restrict__ gacc) {
  const int64_t row_len = n * B;
  for (int64_t a = blockIdx.y; a < A; a += gridDim.y) {
    walk_flat<4>(vals + a * row_len, codes, row_len, B, G, [&](VT v, CT c, int64_t i, int64_t b) {
      const int64_t bin = (B == 1) ? a * G + (int64_t)c : (int64_t)c * (A * B) + a * B + b;
      if (gpos[bin] == (long long)(index_base + i)) gacc[bin] = (typename Acc<VT>::type)v;  // exactly one row per bin
    });
  }
}

__global__ void k_fill_i64(long long* p, int64_t n, long long v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void k_fill_f64(double* p, int64_t n, double v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

static inline dim3 generic_grid(int64_t row_len, int64_t A) {
  int64_t bx = (row_len + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (bx > cap) bx = cap;
  const int by = (int)(A < 65535 ? A : 65535);
  if (bx * by > cap * 4) bx = (cap * 4 + by - 1) / by;
  if (bx < 1) bx = 1;
  return dim3((unsigned)bx, (unsigned)by);
}
static inline int fill_blocks(int64_t n) { return (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184); }

constexpr int kSqThreads = 1024;

template <typename VT, typename CT>
int launch_sqdev(const floxb_chunk_t* chunks, int64_t n_chunks, int64_t A, int64_t B, int64_t G, const double* means,
                 int skipna, double* m2, int init, cudaStream_t st) {
  const int64_t nbins = A * B * G;
  if (init && nbins > 0) k_fill_f64<<<fill_blocks(nbins), 256, 0, st>>>(m2, nbins, 0.0);
  if (nbins == 0) return FLOXB_OK;
  const size_t smem = (size_t)G * ((sizeof(VT) == 4 && is_float<VT>::value) ? 12 : 16);
  if (A == 1 && B == 1 && smem <= 227 * 1024) {
    constexpr int TILE = kSqThreads * 4;
    auto kern = k_sqdev_smem<VT, CT, kSqThreads>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return FLOXB_ECUDA;
    for (int64_t c0 = 0; c0 < n_chunks; c0 += kMaxChunksPerLaunch) {
      ChunkTable tab;
      const int64_t tiles = fill_chunk_table(chunks, n_chunks, c0, TILE, &tab);
      if (tiles < 0) return (int)tiles;
      if (tiles == 0) continue;
      const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
      kern<<<grid, kSqThreads, smem, st>>>(tab, (int)G, means, skipna, m2);
    }
  } else {
    for (int64_t k = 0; k < n_chunks; ++k) {
      const floxb_chunk_t& ch = chunks[k];
      if (ch.n < 0 || (ch.n > 0 && (!ch.values || !ch.codes))) return FLOXB_EINVAL;
      if (ch.n == 0) continue;
      const VT* v = reinterpret_cast<const VT*>(ch.values);
      const CT* c = reinterpret_cast<const CT*>(ch.codes);
      if (B > 1 && B <= 32 && (B & (B - 1)) == 0) {
        const dim3 g2 = generic_grid((ch.n * B + 7) / 8, A);  // one warp per 64-step segment, 8 warps per block
        switch (B) {
          case 2: k_sqdev_global_bruns<VT, CT, 2><<<g2, 256, 0, st>>>(v, c, A, ch.n, G, means, skipna, m2); break;
          case 4: k_sqdev_global_bruns<VT, CT, 4><<<g2, 256, 0, st>>>(v, c, A, ch.n, G, means, skipna, m2); break;
          case 8: k_sqdev_global_bruns<VT, CT, 8><<<g2, 256, 0, st>>>(v, c, A, ch.n, G, means, skipna, m2); break;
          case 16: k_sqdev_global_bruns<VT, CT, 16><<<g2, 256, 0, st>>>(v, c, A, ch.n, G, means, skipna, m2); break;
          default: k_sqdev_global_bruns<VT, CT, 32><<<g2, 256, 0, st>>>(v, c, A, ch.n, G, means, skipna, m2); break;
        }
        continue;
      }
      k_sqdev_global<VT, CT><<<generic_grid(ch.n * B, A), 256, 0, st>>>(v, c, A, ch.n, B, G, means, skipna, m2);
    }
  }
  return cudaGetLastError() == cudaSuccess ? FLOXB_OK : FLOXB_ECUDA;
}

template <typename VT, typename CT>
int launch_argext(const floxb_chunk_t* chunks, int64_t n_chunks, int64_t A, int64_t B, int64_t G, const int64_t* key,
                  const int64_t* nnan, int skipna, const int64_t* index_base, int64_t* idx, int init, cudaStream_t st) {
  const int64_t nbins = A * B * G;
  if (init && nbins > 0) k_fill_i64<<<fill_blocks(nbins), 256, 0, st>>>(reinterpret_cast<long long*>(idx), nbins, INT64_MAX);
  if (nbins == 0) return FLOXB_OK;
  for (int64_t k = 0; k < n_chunks; ++k) {
    const floxb_chunk_t& ch = chunks[k];
    if (ch.n < 0 || (ch.n > 0 && (!ch.values || !ch.codes))) return FLOXB_EINVAL;
    if (ch.n == 0) continue;
    k_argext_global<VT, CT><<<generic_grid(ch.n * B, A), 256, 0, st>>>(
        reinterpret_cast<const VT*>(ch.values), reinterpret_cast<const CT*>(ch.codes), A, ch.n, B, G,
        reinterpret_cast<const long long*>(key), reinterpret_cast<const long long*>(nnan), skipna,
        index_base ? index_base[k] : 0, reinterpret_cast<long long*>(idx));
  }
  return cudaGetLastError() == cudaSuccess ? FLOXB_OK : FLOXB_ECUDA;
}

template <typename VT, typename CT>
int launch_firstlast(const floxb_chunk_t* chunks, int64_t n_chunks, int64_t A, int64_t B, int64_t G, int is_last, int skipna,
                     const int64_t* index_base, int64_t* pos, int init, cudaStream_t st) {
  const int64_t nbins = A * B * G;
  if (init && nbins > 0)
    k_fill_i64<<<fill_blocks(nbins), 256, 0, st>>>(reinterpret_cast<long long*>(pos), nbins, is_last ? INT64_MIN : INT64_MAX);
  if (nbins == 0) return FLOXB_OK;
  for (int64_t k = 0; k < n_chunks; ++k) {
    const floxb_chunk_t& ch = chunks[k];
    if (ch.n < 0 || (ch.n > 0 && (!ch.values || !ch.codes))) return FLOXB_EINVAL;
    if (ch.n == 0) continue;
    k_firstlast_pos<VT, CT><<<generic_grid(ch.n * B, A), 256, 0, st>>>(
        reinterpret_cast<const VT*>(ch.values), reinterpret_cast<const CT*>(ch.codes), A, ch.n, B, G, is_last, skipna,
        index_base ? index_base[k] : 0, reinterpret_cast<long long*>(pos));
  }
  return cudaGetLastError() == cudaSuccess ? FLOXB_OK : FLOXB_ECUDA;
}

template <typename VT, typename CT>
int launch_take(const floxb_chunk_t* chunks, int64_t n_chunks, int64_t A, int64_t B, int64_t G, const int64_t* pos,
                const int64_t* index_base, void* acc, int init, cudaStream_t st) {
  using AccT = typename Acc<VT>::type;
  const int64_t nbins = A * B * G;
  if (init && nbins > 0) {
    if constexpr (is_float<VT>::value) k_fill_f64<<<fill_blocks(nbins), 256, 0, st>>>(reinterpret_cast<double*>(acc), nbins, 0.0);
    else k_fill_i64<<<fill_blocks(nbins), 256, 0, st>>>(reinterpret_cast<long long*>(acc), nbins, 0LL);
  }
  if (nbins == 0) return FLOXB_OK;
  for (int64_t k = 0; k < n_chunks; ++k) {
    const floxb_chunk_t& ch = chunks[k];
    if (ch.n < 0 || (ch.n > 0 && (!ch.values || !ch.codes))) return FLOXB_EINVAL;
    if (ch.n == 0) continue;
    k_take_pos<VT, CT><<<generic_grid(ch.n * B, A), 256, 0, st>>>(
        reinterpret_cast<const VT*>(ch.values), reinterpret_cast<const CT*>(ch.codes), A, ch.n, B, G,
        reinterpret_cast<const long long*>(pos), index_base ? index_base[k] : 0, reinterpret_cast<AccT*>(acc));
  }
  return cudaGetLastError() == cudaSuccess ? FLOXB_OK : FLOXB_ECUDA;
}

}  // namespace floxb

using namespace floxb;

#define FLOXB_DISPATCH_VT_CT(FN, ...)                                                          \
  do {                                                                                         \
    if (code_dtype == FLOXB_I64) {                                                             \
      switch (value_dtype) {                                                                   \
        case FLOXB_F32: return FN<float, int64_t>(__VA_ARGS__);                                \
        case FLOXB_F64: return FN<double, int64_t>(__VA_ARGS__);                               \
        case FLOXB_I32: return FN<int32_t, int64_t>(__VA_ARGS__);                              \
        case FLOXB_I64: return FN<int64_t, int64_t>(__VA_ARGS__);                              \
      }                                                                                        \
    } else if (code_dtype == FLOXB_I32) {                                                      \
      switch (value_dtype) {                                                                   \
        case FLOXB_F32: return FN<float, int32_t>(__VA_ARGS__);                                \
        case FLOXB_F64: return FN<double, int32_t>(__VA_ARGS__);                               \
        case FLOXB_I32: return FN<int32_t, int32_t>(__VA_ARGS__);                              \
        case FLOXB_I64: return FN<int64_t, int32_t>(__VA_ARGS__);                              \
      }                                                                                        \
    }                                                                                          \
    return FLOXB_EDTYPE;                                                                       \
  } while (0)

extern "C" int floxb_sqdev(const floxb_chunk_t* chunks, int64_t n_chunks, int value_dtype, int code_dtype, int64_t A,
                           int64_t B, int64_t n_groups, const double* means, int skipna, double* acc_m2, int init,
                           void* stream) {
  if (n_chunks < 0 || A < 0 || B < 0 || n_groups < 0 || (n_chunks > 0 && !chunks)) return FLOXB_EINVAL;
  if (A * B * n_groups > 0 && (!means || !acc_m2)) return FLOXB_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  FLOXB_DISPATCH_VT_CT(launch_sqdev, chunks, n_chunks, A, B, n_groups, means, skipna, acc_m2, init, st);
}

extern "C" int floxb_argext(const floxb_chunk_t* chunks, int64_t n_chunks, int value_dtype, int code_dtype, int64_t A,
                            int64_t B, int64_t n_groups, const int64_t* acc_key, const int64_t* nnan, int skipna,
                            const int64_t* index_base, int64_t* acc_idx, int init, void* stream) {
  if (n_chunks < 0 || A < 0 || B < 0 || n_groups < 0 || (n_chunks > 0 && !chunks)) return FLOXB_EINVAL;
  if (A * B * n_groups > 0 && (!acc_key || !nnan || !acc_idx)) return FLOXB_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  FLOXB_DISPATCH_VT_CT(launch_argext, chunks, n_chunks, A, B, n_groups, acc_key, nnan, skipna, index_base, acc_idx, init, st);
}

extern "C" int floxb_group_means(int value_dtype, const void* acc, const int64_t* cnt, const int64_t* nnan, int64_t n_bins,
                                 int skipna, double* means, void* stream) {
  if (n_bins < 0) return FLOXB_EINVAL;
  if (n_bins == 0) return FLOXB_OK;
  if (!acc || !cnt || !nnan || !means) return FLOXB_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int blocks = fill_blocks(n_bins);
  const long long* c = reinterpret_cast<const long long*>(cnt);
  const long long* m = reinterpret_cast<const long long*>(nnan);
  switch (value_dtype) {
    case FLOXB_F32: k_group_means<float><<<blocks, 256, 0, st>>>(reinterpret_cast<const double*>(acc), c, m, n_bins, skipna, means); break;
    case FLOXB_F64: k_group_means<double><<<blocks, 256, 0, st>>>(reinterpret_cast<const double*>(acc), c, m, n_bins, skipna, means); break;
    case FLOXB_I32: k_group_means<int32_t><<<blocks, 256, 0, st>>>(reinterpret_cast<const long long*>(acc), c, m, n_bins, skipna, means); break;
    case FLOXB_I64: k_group_means<int64_t><<<blocks, 256, 0, st>>>(reinterpret_cast<const long long*>(acc), c, m, n_bins, skipna, means); break;
    default: return FLOXB_EDTYPE;
  }
  return cudaGetLastError() == cudaSuccess ? FLOXB_OK : FLOXB_ECUDA;
}

extern "C" int floxb_firstlast(const floxb_chunk_t* chunks, int64_t n_chunks, int value_dtype, int code_dtype, int64_t A,
                               int64_t B, int64_t n_groups, int is_last, int skipna, const int64_t* index_base,
                               int64_t* acc_pos, int init, void* stream) {
  if (n_chunks < 0 || A < 0 || B < 0 || n_groups < 0 || (n_chunks > 0 && !chunks)) return FLOXB_EINVAL;
  if (A * B * n_groups > 0 && !acc_pos) return FLOXB_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  FLOXB_DISPATCH_VT_CT(launch_firstlast, chunks, n_chunks, A, B, n_groups, is_last, skipna, index_base, acc_pos, init, st);
}

extern "C" int floxb_take(const floxb_chunk_t* chunks, int64_t n_chunks, int value_dtype, int code_dtype, int64_t A, int64_t B,
                          int64_t n_groups, const int64_t* pos, const int64_t* index_base, void* acc, int init, void* stream) {
  if (n_chunks < 0 || A < 0 || B < 0 || n_groups < 0 || (n_chunks > 0 && !chunks)) return FLOXB_EINVAL;
  if (A * B * n_groups > 0 && (!pos || !acc)) return FLOXB_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  FLOXB_DISPATCH_VT_CT(launch_take, chunks, n_chunks, A, B, n_groups, pos, index_base, acc, init, st);
}
