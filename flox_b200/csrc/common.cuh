// flox_b200 — shared device helpers for the grouped-reduction kernels (sm_100a).
//
// Data model (mirrors what flox.core.chunk_reduce hands an engine, reference
// flox/core.py:298-316,368-370): a chunk is `values[A, n, B]` (C-contiguous; A = kept
// leading dims, n = the reduced axis, B = kept trailing dims) plus integer group codes
// `codes[n]` shared by every (a, b).  Codes outside [0, G) are skipped — that is what the
// reference's "-1 -> sentinel bin -> strip" dance (flox/factorize.py:201-210,
// flox/core.py:381-383) amounts to.  Accumulators are laid out `[A, B, G]`.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/flox_b200.h"

namespace floxb {

constexpr int kMaxChunksPerLaunch = 128;

// Chunk table passed BY VALUE in kernel parameter space (no device allocation, no copy): 3 KB of the 4 KB
// parameter space, so BASELINE config 2's 100 chunks are ONE persistent launch.
struct ChunkTable {
  const void* values[kMaxChunksPerLaunch];
  const void* codes[kMaxChunksPerLaunch];
  int64_t n[kMaxChunksPerLaunch];  // reduced length of each chunk
  int nchunks;
};

template <typename T> struct is_float { static constexpr bool value = false; };
template <> struct is_float<float> { static constexpr bool value = true; };
template <> struct is_float<double> { static constexpr bool value = true; };

// Wide accumulator per value type: floats accumulate in f64 (the reference keeps f32 and
// relies on pairwise reduceat, flox/aggregations.py:962-968; f64 + one final rounding is at
// least as accurate), integers in int64 (flox/xrdtypes.py:175-185).
template <typename T> struct Acc { using type = double; };
template <> struct Acc<int32_t> { using type = long long; };
template <> struct Acc<int64_t> { using type = long long; };

template <typename T> __device__ __forceinline__ bool is_nan(T) { return false; }
template <> __device__ __forceinline__ bool is_nan<float>(float v) { return v != v; }
template <> __device__ __forceinline__ bool is_nan<double>(double v) { return v != v; }

// ---- streaming loads (read-once data: bypass L1 allocation) -----------------------------
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ int4 ldg_stream(const int4* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ longlong2 ldg_stream(const longlong2* p) {
  longlong2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.s64 {%0,%1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ double2 ldg_stream(const double2* p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

// Load 4 consecutive elements of T starting at a 16-byte-aligned (for 4-byte T) or
// 32-byte-aligned-pair (for 8-byte T: two 16-byte loads) address.
template <typename T> struct Vec4 { T v[4]; };

template <typename T> __device__ __forceinline__ Vec4<T> load4(const T* p);
template <> __device__ __forceinline__ Vec4<float> load4<float>(const float* p) {
  float4 r = ldg_stream(reinterpret_cast<const float4*>(p));
  return {{r.x, r.y, r.z, r.w}};
}
template <> __device__ __forceinline__ Vec4<int32_t> load4<int32_t>(const int32_t* p) {
  int4 r = ldg_stream(reinterpret_cast<const int4*>(p));
  return {{r.x, r.y, r.z, r.w}};
}
template <> __device__ __forceinline__ Vec4<int64_t> load4<int64_t>(const int64_t* p) {
  longlong2 a = ldg_stream(reinterpret_cast<const longlong2*>(p));
  longlong2 b = ldg_stream(reinterpret_cast<const longlong2*>(p) + 1);
  return {{(int64_t)a.x, (int64_t)a.y, (int64_t)b.x, (int64_t)b.y}};
}
template <> __device__ __forceinline__ Vec4<double> load4<double>(const double* p) {
  double2 a = ldg_stream(reinterpret_cast<const double2*>(p));
  double2 b = ldg_stream(reinterpret_cast<const double2*>(p) + 1);
  return {{a.x, a.y, b.x, b.y}};
}

// ---- order-preserving integer keys for float min/max ------------------------------------
// Signed-integer order of the key equals the numeric order of the float, so the NATIVE
// 32-bit shared-memory ATOMS.MAX/MIN.S32 can be used for f32 (f64 uses a 64-bit CAS loop).
__device__ __forceinline__ int32_t f32_key(float f) {
  int32_t b = __float_as_int(f);
  return b ^ ((b >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float f32_unkey(int32_t k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }
__device__ __forceinline__ long long f64_key(double d) {
  long long b = __double_as_longlong(d);
  return b ^ ((b >> 63) & 0x7fffffffffffffffLL);
}
__device__ __forceinline__ double f64_unkey(long long k) {
  return __longlong_as_double(k ^ ((k >> 63) & 0x7fffffffffffffffLL));
}

// Key type / conversion per value type (ints are their own keys).
template <typename T> struct Key;
template <> struct Key<float> {
  using type = int32_t;
  static __device__ __forceinline__ type enc(float v) { return f32_key(v); }
  static __device__ __forceinline__ float dec(type k) { return f32_unkey(k); }
  static constexpr type lowest = INT32_MIN, highest = INT32_MAX;
};
template <> struct Key<int32_t> {
  using type = int32_t;
  static __device__ __forceinline__ type enc(int32_t v) { return v; }
  static __device__ __forceinline__ int32_t dec(type k) { return k; }
  static constexpr type lowest = INT32_MIN, highest = INT32_MAX;
};
template <> struct Key<double> {
  using type = long long;
  static __device__ __forceinline__ type enc(double v) { return f64_key(v); }
  static __device__ __forceinline__ double dec(type k) { return f64_unkey(k); }
  static constexpr type lowest = INT64_MIN, highest = INT64_MAX;
};
template <> struct Key<int64_t> {
  using type = long long;
  static __device__ __forceinline__ type enc(int64_t v) { return (long long)v; }
  static __device__ __forceinline__ int64_t dec(type k) { return (int64_t)k; }
  static constexpr type lowest = INT64_MIN, highest = INT64_MAX;
};

// ---- atomics ------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_mul(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(__longlong_as_double((long long)assumed) * v));
  } while (old != assumed);
}
__device__ __forceinline__ void atomic_mul(long long* addr, long long v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    old = atomicCAS(a, assumed, (unsigned long long)((long long)assumed * v));
  } while (old != assumed);
}
__device__ __forceinline__ void atomic_add(double* a, double v) { atomicAdd(a, v); }
__device__ __forceinline__ void atomic_add(long long* a, long long v) {
  atomicAdd(reinterpret_cast<unsigned long long*>(a), (unsigned long long)v);
}

// Persistent walk over the tiles (THREADS*4 rows each) of a chunk table.  (tools/kbench3.py: a three-buffer
// walk with per-tile chunk lookup and partial-tile branches cost 6-11 points of HBM roofline against this
// two-buffer loop.)  The hot loop only ever sees FULL, 16-byte-aligned tiles:
// CTA b takes every gridDim.x-th full tile of the concatenated chunks (the round-robin position carries over
// chunk boundaries, so the load stays balanced), with the 128-bit loads of its next tile in flight while
// `add_row(value, code)` runs on the current one.  Chunk tails (n % TILE rows) and unaligned chunks are
// handled afterwards by a scalar loop, spread over the CTAs.  `after_tile()` runs once per full tile.
// `add_tile(Vec4 values, Vec4 codes)` handles the 4 rows a thread owns in a full tile (so that a kernel can overlap the
// latencies of the four rows' atomics), `add_row(value, code)` the scalar leftovers.
template <typename VT, typename CT, int THREADS, typename AddTile, typename AddRow, typename AfterTile>
__device__ __forceinline__ void walk_tiles4(const ChunkTable& tab, AddTile add_tile, AddRow add_row, AfterTile after_tile) {
  constexpr int TILE = THREADS * 4;
  const int nch = tab.nchunks;
  auto full_tiles = [&](int k) -> int64_t {
    const bool aligned = ((reinterpret_cast<uintptr_t>(tab.values[k]) | reinterpret_cast<uintptr_t>(tab.codes[k])) & 15) == 0;
    return aligned ? tab.n[k] / TILE : 0;
  };
  int ch = 0;
  int64_t local = blockIdx.x, ft = nch > 0 ? full_tiles(0) : 0;
  auto settle = [&] {
    while (ch < nch && local >= ft) {
      local -= ft;
      if (++ch < nch) ft = full_tiles(ch);
    }
  };
  settle();
  Vec4<VT> v0, v1;
  Vec4<CT> c0, c1;
  bool have = ch < nch;
  if (have) {
    v0 = load4<VT>(reinterpret_cast<const VT*>(tab.values[ch]) + local * TILE + threadIdx.x * 4);
    c0 = load4<CT>(reinterpret_cast<const CT*>(tab.codes[ch]) + local * TILE + threadIdx.x * 4);
  }
  while (have) {
    local += gridDim.x;
    settle();
    have = ch < nch;
    if (have) {
      v1 = load4<VT>(reinterpret_cast<const VT*>(tab.values[ch]) + local * TILE + threadIdx.x * 4);
      c1 = load4<CT>(reinterpret_cast<const CT*>(tab.codes[ch]) + local * TILE + threadIdx.x * 4);
    }
    add_tile(v0, c0);
    after_tile();
    v0 = v1;
    c0 = c1;
  }
  // scalar leftovers: piece j of chunk k goes to CTA (k + j) mod gridDim.x
  for (int k = 0; k < nch; ++k) {
    const VT* vals = reinterpret_cast<const VT*>(tab.values[k]);
    const CT* codes = reinterpret_cast<const CT*>(tab.codes[k]);
    const int64_t start = full_tiles(k) * TILE, n = tab.n[k];
    int64_t j = ((int64_t)blockIdx.x - k) % (int64_t)gridDim.x;
    if (j < 0) j += gridDim.x;
    for (int64_t base = start + j * TILE; base < n; base += (int64_t)gridDim.x * TILE) {
      const int64_t end = (n - base < TILE) ? n : base + TILE;
      for (int64_t i = base + threadIdx.x; i < end; i += THREADS) add_row(vals[i], codes[i]);
    }
  }
}

template <typename VT, typename CT, int THREADS, typename AddRow, typename AfterTile>
__device__ __forceinline__ void walk_tiles(const ChunkTable& tab, AddRow add_row, AfterTile after_tile) {
  walk_tiles4<VT, CT, THREADS>(tab, [&](const Vec4<VT>& v, const Vec4<CT>& c) {
#pragma unroll
    for (int k = 0; k < 4; ++k) add_row(v.v[k], c.v[k]);
  }, add_row, after_tile);
}

// Grid-stride walk over the flat elements j = i*B + b of one `values[n, B]` slab for the L2-atomic kernels, U
// elements per thread per step: all U (code, value) loads are issued before the first `apply(value, code, i, b)`,
// so U loads are in flight per thread.  (ncu r01 on config 4, one element per step: nothing above 41% busy, 29
// cycles of long-scoreboard stall per issue — pure load latency.)  (i, b) advance incrementally: one 64-bit
// division per thread, not per element.  Elements whose code is outside [0, G) are skipped.
template <int U, typename VT, typename CT, typename Apply>
__device__ __forceinline__ void walk_flat(const VT* __restrict__ row, const CT* __restrict__ codes, int64_t row_len, int64_t B,
                                          int64_t G, Apply apply) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t i = j / B, b = j - (j / B) * B;
  const int64_t di = stride / B, db = stride - (stride / B) * B;
  while (j < row_len) {
    VT v[U];
    CT c[U];
    int64_t ii[U], bb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool ok = j < row_len;
      ii[u] = i;
      bb[u] = b;
      c[u] = ok ? codes[i] : (CT)-1;
      v[u] = ok ? row[j] : VT(0);
      j += stride;
      i += di;
      b += db;
      if (b >= B) { b -= B; ++i; }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if ((unsigned long long)c[u] < (unsigned long long)G) apply(v[u], c[u], ii[u], bb[u]);
  }
}

// Walk over `values[n, B]` for a SMALL power-of-two B (2..32) so that every thread stays in ONE column and sees the rows
// in order: lane = (row within the warp's row group) * B + column, a warp reads 32 consecutive elements (one fully
// coalesced request) per step and owns SEG_STEPS consecutive steps (a segment of SEG_STEPS * 32/B consecutive rows), U steps
// in flight.  `elem(v, code, row)` runs for every element in row order per thread, `segment_end()` after the last element
// of each segment.  With sorted or blocked labels (flox's method="blockwise": rechunk.py:158-223 puts every group in one
// block) a thread sees long runs of one code and can keep the running result in registers: one flush per RUN, not per
// element.  Rows whose code is outside [0, G) are passed through (the caller's run logic skips them).
template <int B, int U, int SEG_STEPS, typename VT, typename CT, typename Elem, typename SegEnd>
__device__ __forceinline__ void walk_rows_small_b(const VT* __restrict__ vals, const CT* __restrict__ codes, int64_t n,
                                                  Elem elem, SegEnd segment_end) {
  static_assert(B >= 1 && B <= 32 && (B & (B - 1)) == 0, "B must be a power of two <= 32");
  static_assert(SEG_STEPS % U == 0, "segment steps must be a multiple of the unroll");
  constexpr int RW = 32 / B;  // rows per warp step
  const int lane = threadIdx.x & 31;
  const int rw = lane / B;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  constexpr int64_t SEG_ROWS = (int64_t)SEG_STEPS * RW;
  for (int64_t seg0 = warp * SEG_ROWS; seg0 < n; seg0 += warps * SEG_ROWS) {
    const int64_t seg1 = seg0 + SEG_ROWS < n ? seg0 + SEG_ROWS : n;
    for (int64_t r0 = seg0 + rw; r0 < seg1; r0 += (int64_t)U * RW) {
      VT v[U];
      CT c[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t r = r0 + (int64_t)u * RW;
        const bool ok = r < seg1;
        c[u] = ok ? codes[r] : (CT)-1;
        v[u] = ok ? vals[r * B + (lane & (B - 1))] : VT(0);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t r = r0 + (int64_t)u * RW;
        if (r < seg1) elem(v[u], c[u], r);
      }
    }
    segment_end();
  }
}

// Host side: fill `tab` with chunks [c0, c0 + m) and return the number of tiles (of `tile` rows) they hold, or a
// negative FLOXB_E* code.
inline int64_t fill_chunk_table(const floxb_chunk_t* chunks, int64_t n_chunks, int64_t c0, int64_t tile, ChunkTable* tab) {
  const int m = (int)((n_chunks - c0 < kMaxChunksPerLaunch) ? (n_chunks - c0) : kMaxChunksPerLaunch);
  tab->nchunks = m;
  int64_t tiles = 0;
  for (int k = 0; k < m; ++k) {
    const floxb_chunk_t& ch = chunks[c0 + k];
    if (ch.n < 0 || (ch.n > 0 && (!ch.values || !ch.codes))) return FLOXB_EINVAL;
    tab->values[k] = ch.values;
    tab->codes[k] = ch.codes;
    tab->n[k] = ch.n;
    tiles += (ch.n + tile - 1) / tile;
  }
  return tiles;
}

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace floxb
