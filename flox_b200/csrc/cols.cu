// "Column-owner" kernels for chunks with kept trailing dims: values[A, n, B] with B > 1, labels along n.
// This is BASELINE config 3's native layout (time, lat, lon) grouped by time.month: the reference
// transposes the reduced axis to the end and copies (flox/core.py:1014-1018, _move_reduce_dims_to_end +
// reshape); here the array is consumed as it lies in memory.
//
// A thread owns VEC adjacent columns b..b+VEC-1 (one 16-byte load per row).  Rows are visited GROUP BY
// GROUP through a (code-sorted) row permutation built once per label vector by floxb_group_index, so
// the running result of the current group lives in REGISTERS: no atomics, no shared memory, any G.
// Each (a, b, g) bin is written exactly once per launch by its owner thread.
// HBM traffic: every value read once (4 B / element for f32), labels read once to build the
// permutation; accumulators written once.
#include <type_traits>

#include "common.cuh"

namespace floxb {

// ---- group index: counting sort of row numbers by code -------------------------------------------
template <typename CT>
__global__ void __launch_bounds__(256)
k_gi_count(const CT* __restrict__ codes, int64_t n, int64_t G, unsigned long long* __restrict__ goff) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const CT c = codes[i];
    if ((unsigned long long)c < (unsigned long long)G) atomicAdd(&goff[(int64_t)c + 1], 1ull);
  }
}

// single-block exclusive scan in place over goff[0..G] (goff[0] == 0 on entry), copy to cursor
__global__ void __launch_bounds__(1024)
k_gi_scan(unsigned long long* __restrict__ goff, unsigned long long* __restrict__ cursor, int64_t G) {
  __shared__ unsigned long long warp_tot[32];
  __shared__ unsigned long long carry;
  if (threadIdx.x == 0) carry = 0ull;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t base = 1; base <= G; base += 1024) {
    const int64_t i = base + threadIdx.x;
    unsigned long long v = (i <= G) ? goff[i] : 0ull;
    unsigned long long x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
      unsigned long long t = warp_tot[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xffffffffu, t, d);
        if (lane >= d) t += y;
      }
      warp_tot[lane] = t;
    }
    __syncthreads();
    const unsigned long long incl = x + (warp ? warp_tot[warp - 1] : 0ull) + carry;
    if (i <= G) {
      goff[i] = incl;                 // inclusive count up to group i-1  == start of group i
      if (i < G) cursor[i] = incl;
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) cursor[0] = 0ull;
}

template <typename CT>
__global__ void __launch_bounds__(256)
k_gi_scatter(const CT* __restrict__ codes, int64_t n, int64_t G, unsigned long long* __restrict__ cursor,
             long long* __restrict__ perm) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const CT c = codes[i];
    if ((unsigned long long)c < (unsigned long long)G) perm[atomicAdd(&cursor[(int64_t)c], 1ull)] = i;
  }
}

// ---- column-owner reductions ------------------------------------------------------------------------
template <typename VT, int VEC> struct ColVec { VT v[VEC]; };

template <typename VT, int VEC> __device__ __forceinline__ ColVec<VT, VEC> load_cols(const VT* p) {
  ColVec<VT, VEC> r;
  if constexpr (VEC == 4 && sizeof(VT) == 4) {
    const int4 q = ldg_stream(reinterpret_cast<const int4*>(p));
    const int w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) r.v[k] = *reinterpret_cast<const VT*>(&w[k]);
  } else if constexpr (VEC == 2 && sizeof(VT) == 8) {
    const longlong2 q = ldg_stream(reinterpret_cast<const longlong2*>(p));
    const long long w[2] = {q.x, q.y};
#pragma unroll
    for (int k = 0; k < 2; ++k) r.v[k] = *reinterpret_cast<const VT*>(&w[k]);
  } else if constexpr (VEC == 2 && sizeof(VT) == 4) {
    const int2 q = __ldg(reinterpret_cast<const int2*>(p));
    r.v[0] = *reinterpret_cast<const VT*>(&q.x);
    r.v[1] = *reinterpret_cast<const VT*>(&q.y);
  } else {
#pragma unroll
    for (int k = 0; k < VEC; ++k) r.v[k] = __ldg(p + k);
  }
  return r;
}

// MODE: 0 sum, 1 prod, 2 sumsq, 3 max, 4 min  (accumulator: wide sum, or order-preserving key)
template <typename VT, int MODE> struct ColAcc {
  using AccT = typename Acc<VT>::type;
  using T = typename std::conditional<(MODE >= 3), long long, AccT>::type;
  static __device__ __forceinline__ T ident() {
    if constexpr (MODE == 1) return (T)1;
    else if constexpr (MODE == 3) return INT64_MIN;
    else if constexpr (MODE == 4) return INT64_MAX;
    else return (T)0;
  }
  static __device__ __forceinline__ T step(T a, VT v) {
    if constexpr (MODE == 0) return a + (AccT)v;
    else if constexpr (MODE == 1) return a * (AccT)v;
    else if constexpr (MODE == 2) return a + (AccT)v * (AccT)v;
    else if constexpr (MODE == 3) { const long long k = (long long)Key<VT>::enc(v); return k > a ? k : a; }
    else { const long long k = (long long)Key<VT>::enc(v); return k < a ? k : a; }
  }
  static __device__ __forceinline__ T merge(T a, T b) {
    if constexpr (MODE == 0 || MODE == 2) return a + b;
    else if constexpr (MODE == 1) return a * b;
    else if constexpr (MODE == 3) return a > b ? a : b;
    else return a < b ? a : b;
  }
};

// blockIdx.z owns a RANGE of groups [z*gpt, (z+1)*gpt): the launcher sizes the ranges so that the grid is many
// waves deep (a single wave of column tiles left a 1.7-wave tail: ncu r01, 64% of HBM).
// group_major != 0: accumulators are laid out [G][A*B] (columns fastest) so that a thread's VEC results are one
// full 32-byte sector and a warp writes 1 KB contiguously; the standard [A*B][G] layout scatters 8-byte writes
// G*8 bytes apart (ncu r01: 1.17 GB written for 0.3 GB of accumulators).  floxb_finalize reads either layout.
template <typename VT, int MODE, int VEC, int UNROLL>
__global__ void __launch_bounds__(128)
k_cols(const VT* __restrict__ vals, const long long* __restrict__ perm, const unsigned long long* __restrict__ goff,
       int64_t n, int64_t B, int64_t G, int64_t gpt, typename ColAcc<VT, MODE>::T* __restrict__ gacc,
       unsigned long long* __restrict__ gcnt, unsigned long long* __restrict__ gnan, int accumulate, int group_major) {
  using CA = ColAcc<VT, MODE>;
  using T = typename CA::T;
  const int64_t col = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (col >= B) return;
  const int64_t a = blockIdx.y;
  const int64_t AB = (int64_t)gridDim.y * B;
  const VT* __restrict__ base = vals + a * n * B + col;
  const int ncol = (B - col < VEC) ? (int)(B - col) : VEC;  // only when B % VEC != 0 (scalar VEC == 1 otherwise)
  const int64_t g0 = (int64_t)blockIdx.z * gpt;
  const int64_t g1 = (g0 + gpt < G) ? g0 + gpt : G;
  for (int64_t g = g0; g < g1; ++g) {
    T acc[VEC];
    unsigned nan[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) { acc[k] = CA::ident(); nan[k] = 0u; }
    const int64_t lo = (int64_t)goff[g], hi = (int64_t)goff[g + 1];
    int64_t r = lo;
    for (; r + UNROLL <= hi; r += UNROLL) {
      ColVec<VT, VEC> x[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) x[u] = load_cols<VT, VEC>(base + perm[r + u] * B);
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          const VT v = x[u].v[k];
          if (is_nan(v)) nan[k]++; else acc[k] = CA::step(acc[k], v);
        }
      }
    }
    for (; r < hi; ++r) {
      const ColVec<VT, VEC> x = load_cols<VT, VEC>(base + perm[r] * B);
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const VT v = x.v[k];
        if (is_nan(v)) nan[k]++; else acc[k] = CA::step(acc[k], v);
      }
    }
    const unsigned len = (unsigned)(hi - lo);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      if (k < ncol) {
        const int64_t ab = a * B + col + k;
        const int64_t bin = group_major ? g * AB + ab : ab * G + g;
        if (accumulate) {
          gacc[bin] = CA::merge(gacc[bin], acc[k]);
          gcnt[bin] += len - nan[k];
          gnan[bin] += nan[k];
        } else {
          gacc[bin] = acc[k];
          gcnt[bin] = len - nan[k];
          gnan[bin] = nan[k];
        }
      }
    }
  }
}

// Row-batched layout values[A, n] (B == 1, A >> 1: e.g. (lat*lon, time), what flox's own reshape hands its
// engines, flox/core.py:313-315).  One WARP per row: for each group the lanes stride over that group's rows
// of the permutation (runs of consecutive time steps stay coalesced), accumulate in registers and finish
// with a shuffle reduction; lane 0 writes the (row, group) bin.  No atomics, no shared memory.
template <typename VT, int MODE>
__global__ void __launch_bounds__(256)
k_rows(const VT* __restrict__ vals, const long long* __restrict__ perm, const unsigned long long* __restrict__ goff,
       int64_t A, int64_t n, int64_t G, typename ColAcc<VT, MODE>::T* __restrict__ gacc,
       unsigned long long* __restrict__ gcnt, unsigned long long* __restrict__ gnan, int accumulate) {
  using CA = ColAcc<VT, MODE>;
  using T = typename CA::T;
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t a = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); a < A; a += warps) {
    const VT* __restrict__ row = vals + a * n;
    for (int64_t g = 0; g < G; ++g) {
      T acc = CA::ident();
      unsigned cnt = 0u, nan = 0u;
      const int64_t lo = (int64_t)goff[g], hi = (int64_t)goff[g + 1];
      int64_t k = lo + lane;
      for (; k + 96 < hi; k += 128) {  // 4 independent loads in flight per lane
        VT x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = __ldg(row + perm[k + 32 * u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (is_nan(x[u])) nan++; else { acc = CA::step(acc, x[u]); cnt++; }
        }
      }
      for (; k < hi; k += 32) {
        const VT x = __ldg(row + perm[k]);
        if (is_nan(x)) nan++; else { acc = CA::step(acc, x); cnt++; }
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        const long long o = __shfl_xor_sync(0xffffffffu, *reinterpret_cast<long long*>(&acc), d);
        acc = CA::merge(acc, *reinterpret_cast<const T*>(&o));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
        nan += __shfl_xor_sync(0xffffffffu, nan, d);
      }
      if (lane == 0) {
        const int64_t bin = a * G + g;
        if (accumulate) {
          gacc[bin] = CA::merge(gacc[bin], acc);
          gcnt[bin] += cnt;
          gnan[bin] += nan;
        } else {
          gacc[bin] = acc;
          gcnt[bin] = cnt;
          gnan[bin] = nan;
        }
      }
    }
  }
}

// Row-batched layout with FEW groups (G <= 32: months, seasons, hours): the row is read sequentially and fully
// coalesced instead of through the permutation (k_rows pays an 8-byte permutation read per 4-byte value and is
// L1TEX-bound at ~30% of the HBM roof).  The block first rebuilds the per-row-position code table (u8) in shared
// memory from the permutation; then a warp takes a row, every lane keeps PRIVATE bins [G + 1][32 lanes] in shared
// memory (slot (g, lane) lives in bank pair `lane`: no conflicts, no atomics; bin G swallows rows without a
// group), and the 32 lane-partials of a group are summed by lane g with a skewed, conflict-free read.
constexpr int kRowsBinsMaxG = 32;
constexpr int kRowsBinsWarps = 8;

template <typename VT, int MODE>
__global__ void __launch_bounds__(kRowsBinsWarps * 32)
k_rows_bins(const VT* __restrict__ vals, const long long* __restrict__ perm, const unsigned long long* __restrict__ goff,
            int64_t A, int64_t n, int G, typename ColAcc<VT, MODE>::T* __restrict__ gacc,
            unsigned long long* __restrict__ gcnt, unsigned long long* __restrict__ gnan, int accumulate) {
  using CA = ColAcc<VT, MODE>;
  using T = typename CA::T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int G1 = G + 1;
  T* bins = reinterpret_cast<T*>(smem_raw) + (size_t)warp * G1 * 32;                              // [G1][32]
  unsigned* nans = reinterpret_cast<unsigned*>(smem_raw + (size_t)kRowsBinsWarps * G1 * 32 * sizeof(T)) + (size_t)warp * G1 * 32;
  unsigned char* code = smem_raw + (size_t)kRowsBinsWarps * G1 * 32 * (sizeof(T) + 4);            // [n]
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) code[i] = (unsigned char)G;
  __syncthreads();
  for (int g = 0; g < G; ++g)
    for (int64_t k = (int64_t)goff[g] + threadIdx.x; k < (int64_t)goff[g + 1]; k += blockDim.x) code[perm[k]] = (unsigned char)g;
  __syncthreads();
  const unsigned len = lane < G ? (unsigned)(goff[lane + 1] - goff[lane]) : 0u;  // rows of group `lane`

  const int64_t warps = (int64_t)gridDim.x * kRowsBinsWarps;
  for (int64_t a = (int64_t)blockIdx.x * kRowsBinsWarps + warp; a < A; a += warps) {
    const VT* __restrict__ row = vals + a * n;
    for (int g = 0; g < G1; ++g) {
      bins[g * 32 + lane] = CA::ident();
      nans[g * 32 + lane] = 0u;
    }
    __syncwarp();
    int64_t k = lane;
    // Adjacent row positions usually carry the same code (time-like labels: 30 consecutive days per month), so each
    // lane takes PAIRS (one 8- or 16-byte load) and folds both elements into its bin with a single LDS + STS: the
    // scalar loop was L1TEX-bound (81% busy, 8 shared wavefronts per 32 elements, 48% of the HBM roof).
    if ((n & 1) == 0 && (reinterpret_cast<uintptr_t>(vals) % (2 * sizeof(VT))) == 0) {
      const int64_t np = n >> 1;
      auto fold = [&](const ColVec<VT, 2>& x, int64_t p) {
        const unsigned cc = *reinterpret_cast<const unsigned short*>(code + 2 * p);
        const int s0 = (int)(cc & 0xffu) * 32 + lane, s1 = (int)(cc >> 8) * 32 + lane;
        const bool n0 = is_nan(x.v[0]), n1 = is_nan(x.v[1]);
        if (s0 == s1 && !n0 && !n1) {
          bins[s0] = CA::step(CA::step(bins[s0], x.v[0]), x.v[1]);
        } else {
          if (n0) nans[s0]++; else bins[s0] = CA::step(bins[s0], x.v[0]);
          if (n1) nans[s1]++; else bins[s1] = CA::step(bins[s1], x.v[1]);
        }
      };
      constexpr int UP = 8;
      int64_t p = lane;
      for (; p + 32 * (UP - 1) < np; p += 32 * UP) {
        ColVec<VT, 2> x[UP];
#pragma unroll
        for (int u = 0; u < UP; ++u) x[u] = load_cols<VT, 2>(row + 2 * (p + 32 * u));
#pragma unroll
        for (int u = 0; u < UP; ++u) fold(x[u], p + 32 * u);
      }
      for (; p < np; p += 32) fold(load_cols<VT, 2>(row + 2 * p), p);
      k = n;  // nothing left for the scalar loops
    }
    // 16 independent coalesced loads in flight per lane: 4 left the SM at ~16 KB in flight and 40% of the HBM roof
    constexpr int U = 16;
    for (; k + 32 * (U - 1) < n; k += 32 * U) {
      VT x[U];
#pragma unroll
      for (int u = 0; u < U; ++u) x[u] = __ldg(row + k + 32 * u);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int s = (int)code[k + 32 * u] * 32 + lane;
        if (is_nan(x[u])) nans[s]++; else bins[s] = CA::step(bins[s], x[u]);
      }
    }
    for (; k + 96 < n; k += 128) {
      VT x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) x[u] = __ldg(row + k + 32 * u);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int s = (int)code[k + 32 * u] * 32 + lane;
        if (is_nan(x[u])) nans[s]++; else bins[s] = CA::step(bins[s], x[u]);
      }
    }
    for (; k < n; k += 32) {
      const VT x = __ldg(row + k);
      const int s = (int)code[k] * 32 + lane;
      if (is_nan(x)) nans[s]++; else bins[s] = CA::step(bins[s], x);
    }
    __syncwarp();
    if (lane < G) {  // lane g folds the 32 lane-partials of group g; the skew keeps the 32 readers in 32 banks
      T acc = CA::ident();
      unsigned nan = 0u;
#pragma unroll 8
      for (int j = 0; j < 32; ++j) {
        const int l = (j + lane) & 31;
        acc = CA::merge(acc, bins[lane * 32 + l]);
        nan += nans[lane * 32 + l];
      }
      const int64_t bin = a * G + lane;
      if (accumulate) {
        gacc[bin] = CA::merge(gacc[bin], acc);
        gcnt[bin] += len - nan;
        gnan[bin] += nan;
      } else {
        gacc[bin] = acc;
        gcnt[bin] = len - nan;
        gnan[bin] = nan;
      }
    }
    __syncwarp();
  }
}

template <typename VT, int MODE>
int launch_cols(const void* values, const int64_t* perm, const int64_t* goff, int64_t A, int64_t n, int64_t B, int64_t G,
                void* acc, int64_t* cnt, int64_t* nnan, int init, int group_major, cudaStream_t st) {
  using T = typename ColAcc<VT, MODE>::T;
  if (A * B * G == 0) return FLOXB_OK;
  const VT* v = reinterpret_cast<const VT*>(values);
  auto* p = reinterpret_cast<const long long*>(perm);
  auto* o = reinterpret_cast<const unsigned long long*>(goff);
  T* ga = reinterpret_cast<T*>(acc);
  auto* gc = reinterpret_cast<unsigned long long*>(cnt);
  auto* gn = reinterpret_cast<unsigned long long*>(nnan);
  if (B == 1) {  // row-batched layout: a warp per row
    int64_t blocks = (A + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 32;
    if (blocks > cap) blocks = cap;
    if (group_major) return FLOXB_EINVAL;  // the warp-per-row kernels write the standard [A][G] layout
    const size_t bins_smem = (size_t)kRowsBinsWarps * (G + 1) * 32 * (sizeof(T) + 4) + (size_t)n;
    if (G <= kRowsBinsMaxG && bins_smem <= 100 * 1024) {
      auto kern = k_rows_bins<VT, MODE>;
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess) return FLOXB_ECUDA;
      int64_t nb = (A + kRowsBinsWarps - 1) / kRowsBinsWarps;
      const int64_t per_sm = (int64_t)(200 * 1024 / bins_smem) > 0 ? (int64_t)(200 * 1024 / bins_smem) : 1;
      const int64_t capb = (int64_t)sm_count() * (per_sm < 8 ? per_sm : 8);
      if (nb > capb) nb = capb;
      kern<<<(unsigned)nb, kRowsBinsWarps * 32, bins_smem, st>>>(v, p, o, A, n, (int)G, ga, gc, gn, init ? 0 : 1);
      return cudaGetLastError() == cudaSuccess ? FLOXB_OK : FLOXB_ECUDA;
    }
    k_rows<VT, MODE><<<(unsigned)blocks, 256, 0, st>>>(v, p, o, A, n, G, ga, gc, gn, init ? 0 : 1);
    return cudaGetLastError() == cudaSuccess ? FLOXB_OK : FLOXB_ECUDA;
  }
  if (A > 65535) return FLOXB_ERANGE;
  constexpr int VEC = 16 / sizeof(VT);
  const bool vec_ok = (B % VEC == 0) && ((reinterpret_cast<uintptr_t>(values) & 15) == 0);
  const int64_t threads = vec_ok ? B / VEC : B;
  const int64_t bxy = ((threads + 127) / 128) * A;
  // aim for >= ~16 waves of 8 resident 128-thread blocks per SM, but keep >= ~64 rows per (thread, group range)
  const int64_t want_blocks = (int64_t)sm_count() * 8 * 16;
  int64_t zsplit = (want_blocks + bxy - 1) / bxy;
  if (zsplit > G) zsplit = G;
  const int64_t max_by_rows = n / 64 > 0 ? n / 64 : 1;
  if (zsplit > max_by_rows) zsplit = max_by_rows;
  if (zsplit > 65535) zsplit = 65535;
  if (zsplit < 1) zsplit = 1;
  const int64_t gpt = (G + zsplit - 1) / zsplit;
  const unsigned gz = (unsigned)((G + gpt - 1) / gpt);
  if (vec_ok) {
    dim3 grid((unsigned)((threads + 127) / 128), (unsigned)A, gz);
    k_cols<VT, MODE, VEC, 8><<<grid, 128, 0, st>>>(v, p, o, n, B, G, gpt, ga, gc, gn, init ? 0 : 1, group_major);
  } else {
    dim3 grid((unsigned)((B + 127) / 128), (unsigned)A, gz);
    k_cols<VT, MODE, 1, 8><<<grid, 128, 0, st>>>(v, p, o, n, B, G, gpt, ga, gc, gn, init ? 0 : 1, group_major);
  }
  return cudaGetLastError() == cudaSuccess ? FLOXB_OK : FLOXB_ECUDA;
}

template <typename VT>
int cols_mode(int mode, const void* values, const int64_t* perm, const int64_t* goff, int64_t A, int64_t n, int64_t B,
              int64_t G, void* acc, int64_t* cnt, int64_t* nnan, int init, int group_major, cudaStream_t st) {
  switch (mode) {
    case 0: return launch_cols<VT, 0>(values, perm, goff, A, n, B, G, acc, cnt, nnan, init, group_major, st);
    case 1: return launch_cols<VT, 1>(values, perm, goff, A, n, B, G, acc, cnt, nnan, init, group_major, st);
    case 2: return launch_cols<VT, 2>(values, perm, goff, A, n, B, G, acc, cnt, nnan, init, group_major, st);
    case 3: return launch_cols<VT, 3>(values, perm, goff, A, n, B, G, acc, cnt, nnan, init, group_major, st);
    case 4: return launch_cols<VT, 4>(values, perm, goff, A, n, B, G, acc, cnt, nnan, init, group_major, st);
  }
  return FLOXB_EINVAL;
}

}  // namespace floxb

using namespace floxb;

extern "C" int floxb_group_index(const void* codes, int code_dtype, int64_t n, int64_t n_groups, int64_t* perm,
                                 int64_t* group_offsets, int64_t* cursor, void* stream) {
  if (n < 0 || n_groups < 0) return FLOXB_EINVAL;
  if (!group_offsets || (n_groups > 0 && !cursor) || (n > 0 && (!codes || !perm))) return FLOXB_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  auto* goff = reinterpret_cast<unsigned long long*>(group_offsets);
  auto* cur = reinterpret_cast<unsigned long long*>(cursor);
  cudaMemsetAsync(goff, 0, (size_t)(n_groups + 1) * 8, st);
  if (n_groups == 0) return cudaGetLastError() == cudaSuccess ? FLOXB_OK : FLOXB_ECUDA;
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (code_dtype == FLOXB_I64) {
    if (n > 0) k_gi_count<int64_t><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const int64_t*>(codes), n, n_groups, goff);
    k_gi_scan<<<1, 1024, 0, st>>>(goff, cur, n_groups);
    if (n > 0) k_gi_scatter<int64_t><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const int64_t*>(codes), n, n_groups, cur, reinterpret_cast<long long*>(perm));
  } else if (code_dtype == FLOXB_I32) {
    if (n > 0) k_gi_count<int32_t><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const int32_t*>(codes), n, n_groups, goff);
    k_gi_scan<<<1, 1024, 0, st>>>(goff, cur, n_groups);
    if (n > 0) k_gi_scatter<int32_t><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const int32_t*>(codes), n, n_groups, cur, reinterpret_cast<long long*>(perm));
  } else {
    return FLOXB_EDTYPE;
  }
  return cudaGetLastError() == cudaSuccess ? FLOXB_OK : FLOXB_ECUDA;
}

extern "C" int floxb_reduce_cols(const void* values, int value_dtype, const int64_t* perm, const int64_t* group_offsets,
                                 int64_t A, int64_t n, int64_t B, int64_t n_groups, int mode, void* acc, int64_t* cnt,
                                 int64_t* nnan, int init, int group_major, void* stream) {
  if (A < 0 || n < 0 || B < 0 || n_groups < 0) return FLOXB_EINVAL;
  if (A * B * n_groups > 0 && (!acc || !cnt || !nnan || !group_offsets || (n > 0 && (!values || !perm)))) return FLOXB_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (value_dtype) {
    case FLOXB_F32: return cols_mode<float>(mode, values, perm, group_offsets, A, n, B, n_groups, acc, cnt, nnan, init, group_major, st);
    case FLOXB_F64: return cols_mode<double>(mode, values, perm, group_offsets, A, n, B, n_groups, acc, cnt, nnan, init, group_major, st);
    case FLOXB_I32: return cols_mode<int32_t>(mode, values, perm, group_offsets, A, n, B, n_groups, acc, cnt, nnan, init, group_major, st);
    case FLOXB_I64: return cols_mode<int64_t>(mode, values, perm, group_offsets, A, n, B, n_groups, acc, cnt, nnan, init, group_major, st);
  }
  return FLOXB_EDTYPE;
}
