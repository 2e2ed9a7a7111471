// Grouped max / min of the non-NaN values (as order-preserving integer keys) + counts, one pass.
// Replaces aggregate_flox.max/min (np.maximum/minimum.reduceat, flox/aggregate_flox.py:214-217)
// and the NaN->identity substitution of nanmax/nanmin (flox/aggregate_flox.py:195-207).
//
// f32 / i32 values use the NATIVE 32-bit shared-memory ATOMS.MAX/MIN.S32 on an order-preserving
// key; f64 / i64 use a 64-bit CAS loop.  NaN rows never touch the key bin: they bump a NaN
// counter, and floxb_finalize decides between propagate (max/min) and skip (nanmax/nanmin).
#include "common.cuh"

namespace floxb {

template <typename K, bool IS_MAX> __device__ __forceinline__ void key_update(K* slot, K k) {
  if constexpr (IS_MAX) atomicMax(slot, k); else atomicMin(slot, k);
}

// RSHIFT: 2^RSHIFT slot copies per group selected by the low lane bits (see k_moments_smem): with 32 copies the
// lanes of a warp never collide on an address or a bank, which matters for small G.
template <typename VT, typename CT, bool IS_MAX, int THREADS, int RSHIFT>
__global__ void __launch_bounds__(THREADS, 1)
k_minmax_smem(const __grid_constant__ ChunkTable tab, int G, long long* __restrict__ gkey,
              unsigned long long* __restrict__ gcnt, unsigned long long* __restrict__ gnan) {
  using K = typename Key<VT>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int R = 1 << RSHIFT;
  const int S = G << RSHIFT;
  K* skey = reinterpret_cast<K*>(smem_raw);
  unsigned* scnt = reinterpret_cast<unsigned*>(skey + S);
  unsigned* snan = scnt + S;
  const int copy = threadIdx.x & (R - 1);
  const K ident = IS_MAX ? Key<VT>::lowest : Key<VT>::highest;
  for (int s = threadIdx.x; s < S; s += THREADS) {
    skey[s] = ident;
    scnt[s] = 0u;
    snan[s] = 0u;
  }
  __syncthreads();

  auto accum = [&](VT v, CT c) {
    if ((unsigned long long)c < (unsigned long long)G) {
      const int s = ((int)c << RSHIFT) | copy;
      if (is_nan(v)) {
        atomicAdd(&snan[s], 1u);
      } else {
        key_update<K, IS_MAX>(&skey[s], Key<VT>::enc(v));
        atomicAdd(&scnt[s], 1u);
      }
    }
  };

  walk_tiles<VT, CT, THREADS>(tab, accum, [] {});
  __syncthreads();

  for (int g = threadIdx.x; g < G; g += THREADS) {
    K best = ident;
    unsigned long long c = 0ull, m = 0ull;
#pragma unroll 4
    for (int r = 0; r < R; ++r) {
      const int s = (g << RSHIFT) | ((r + threadIdx.x) & (R - 1));
      const K k = skey[s];
      best = IS_MAX ? (k > best ? k : best) : (k < best ? k : best);
      c += scnt[s];
      m += snan[s];
    }
    if (c) {
      key_update<long long, IS_MAX>(&gkey[g], (long long)best);
      atomicAdd(&gcnt[g], c);
    }
    if (m) atomicAdd(&gnan[g], m);
  }
}

template <typename VT, typename CT, bool IS_MAX>
__global__ void __launch_bounds__(256)
k_minmax_global(const VT* __restrict__ vals, const CT* __restrict__ codes, int64_t A, int64_t n, int64_t B, int64_t G,
                long long* __restrict__ gkey, unsigned long long* __restrict__ gcnt,
                unsigned long long* __restrict__ gnan) {
  const int64_t row_len = n * B;
  for (int64_t a = blockIdx.y; a < A; a += gridDim.y) {
    walk_flat<4>(vals + a * row_len, codes, row_len, B, G, [&](VT v, CT c, int64_t, int64_t b) {
      const int64_t bin = (B == 1) ? a * G + (int64_t)c : (int64_t)c * (A * B) + a * B + b;  // group-major when B > 1
      if (is_nan(v)) {
        atomicAdd(&gnan[bin], 1ull);
      } else {
        key_update<long long, IS_MAX>(&gkey[bin], (long long)Key<VT>::enc(v));
        atomicAdd(&gcnt[bin], 1ull);
      }
    });
  }
}

__global__ void k_fill_key(long long* key, unsigned long long* cnt, unsigned long long* nnan, int64_t nbins,
                           long long ident) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbins; i += (int64_t)gridDim.x * blockDim.x) {
    key[i] = ident;
    if (cnt) cnt[i] = 0ull;
    if (nnan) nnan[i] = 0ull;
  }
}

constexpr int kMMThreads = 1024;
constexpr size_t kMMSmemBudget = 227 * 1024;

template <typename VT, typename CT, bool IS_MAX>
int launch_minmax(const floxb_chunk_t* chunks, int64_t n_chunks, int64_t A, int64_t B, int64_t G, int64_t* acc_key,
                  int64_t* cnt, int64_t* nnan, int init, int force_path, cudaStream_t st) {
  using K = typename Key<VT>::type;
  const int64_t nbins = A * B * G;
  auto* gkey = reinterpret_cast<long long*>(acc_key);
  auto* gcnt = reinterpret_cast<unsigned long long*>(cnt);
  auto* gnan = reinterpret_cast<unsigned long long*>(nnan);
  if (init && nbins > 0) {
    const int blocks = (int)((nbins + 255) / 256 < 1184 ? (nbins + 255) / 256 : 1184);
    k_fill_key<<<blocks, 256, 0, st>>>(gkey, gcnt, gnan, nbins, IS_MAX ? INT64_MIN : INT64_MAX);
  }
  if (nbins == 0) return FLOXB_OK;
  const size_t smem = (size_t)G * (sizeof(K) + 8);
  bool use_smem = (A == 1 && B == 1 && smem <= kMMSmemBudget);
  if (force_path == 1 && !use_smem) return FLOXB_EINVAL;
  if (force_path == 2) use_smem = false;
  if (use_smem) {
    constexpr int TILE = kMMThreads * 4;
    using KernT = void (*)(const ChunkTable, int, long long*, unsigned long long*, unsigned long long*);
    KernT kern = k_minmax_smem<VT, CT, IS_MAX, kMMThreads, 0>;
    size_t smem_used = smem;
    if (smem * 32 <= kMMSmemBudget) { kern = k_minmax_smem<VT, CT, IS_MAX, kMMThreads, 5>; smem_used = smem * 32; }
    else if (smem * 8 <= kMMSmemBudget) { kern = k_minmax_smem<VT, CT, IS_MAX, kMMThreads, 3>; smem_used = smem * 8; }
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMMSmemBudget) != cudaSuccess)
      return FLOXB_ECUDA;
    for (int64_t c0 = 0; c0 < n_chunks; c0 += kMaxChunksPerLaunch) {
      ChunkTable tab;
      const int64_t tiles = fill_chunk_table(chunks, n_chunks, c0, TILE, &tab);
      if (tiles < 0) return (int)tiles;
      if (tiles == 0) continue;
      if (tiles / sm_count() * TILE >= (int64_t)1 << 32) return FLOXB_ERANGE;
      const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
      kern<<<grid, kMMThreads, smem_used, st>>>(tab, (int)G, gkey, gcnt, gnan);
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
      k_minmax_global<VT, CT, IS_MAX><<<grid, 256, 0, st>>>(reinterpret_cast<const VT*>(ch.values),
                                                          reinterpret_cast<const CT*>(ch.codes), A, ch.n, B, G, gkey,
                                                          gcnt, gnan);
    }
  }
  return cudaGetLastError() == cudaSuccess ? FLOXB_OK : FLOXB_ECUDA;
}

template <typename VT>
int mm_dispatch(int code_dtype, int is_max, const floxb_chunk_t* chunks, int64_t n_chunks, int64_t A, int64_t B,
                int64_t G, int64_t* key, int64_t* cnt, int64_t* nnan, int init, int force_path, cudaStream_t st) {
  if (code_dtype == FLOXB_I64) {
    return is_max ? launch_minmax<VT, int64_t, true>(chunks, n_chunks, A, B, G, key, cnt, nnan, init, force_path, st)
                  : launch_minmax<VT, int64_t, false>(chunks, n_chunks, A, B, G, key, cnt, nnan, init, force_path, st);
  }
  if (code_dtype == FLOXB_I32) {
    return is_max ? launch_minmax<VT, int32_t, true>(chunks, n_chunks, A, B, G, key, cnt, nnan, init, force_path, st)
                  : launch_minmax<VT, int32_t, false>(chunks, n_chunks, A, B, G, key, cnt, nnan, init, force_path, st);
  }
  return FLOXB_EDTYPE;
}

}  // namespace floxb

using namespace floxb;

extern "C" int floxb_minmax(const floxb_chunk_t* chunks, int64_t n_chunks, int value_dtype, int code_dtype, int64_t A,
                            int64_t B, int64_t n_groups, int is_max, int64_t* acc_key, int64_t* cnt, int64_t* nnan,
                            int init, int force_path, void* stream) {
  if (n_chunks < 0 || A < 0 || B < 0 || n_groups < 0 || (n_chunks > 0 && !chunks)) return FLOXB_EINVAL;
  if (A * B * n_groups > 0 && (!acc_key || !cnt || !nnan)) return FLOXB_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (value_dtype) {
    case FLOXB_F32: return mm_dispatch<float>(code_dtype, is_max, chunks, n_chunks, A, B, n_groups, acc_key, cnt, nnan, init, force_path, st);
    case FLOXB_F64: return mm_dispatch<double>(code_dtype, is_max, chunks, n_chunks, A, B, n_groups, acc_key, cnt, nnan, init, force_path, st);
    case FLOXB_I32: return mm_dispatch<int32_t>(code_dtype, is_max, chunks, n_chunks, A, B, n_groups, acc_key, cnt, nnan, init, force_path, st);
    case FLOXB_I64: return mm_dispatch<int64_t>(code_dtype, is_max, chunks, n_chunks, A, B, n_groups, acc_key, cnt, nnan, init, force_path, st);
  }
  return FLOXB_EDTYPE;
}
