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
//           wavefronts per 32 rows with an f64 slot; a 32-bit slot costs half).  f32 sums
//           therefore use the TWO-LEVEL kernel k_moments_smem_f32 below: a 32-bit f32 partial per
//           slot in front of an f64 accumulator, both in shared memory.  f64 / integer values and
//           products keep a single 64-bit slot (exact in f64 / i64).  Bins are merged into the
//           global accumulators with one RED.ADD.F64 / RED.ADD.U64 per non-empty bin per CTA.
//  * global (anything else): every element issues RED.ADD.F64 + RED.ADD.U64 straight to the
//           L2-resident accumulators.
// HBM traffic is the algorithmic minimum: every value and code is read exactly once
// (ld.global.nc.L1::no_allocate, 128-bit), accumulators are written once.
#include <cstdlib>
#include <type_traits>

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


// ------------------------------------------------------------------------------------------
// shared-memory privatised bins
// ------------------------------------------------------------------------------------------
// RSHIFT: every group owns 2^RSHIFT slot copies, selected by the low lane bits.  With 32 copies the
// lanes of a warp never share a slot and slot (c*32 + lane) sits in bank `lane`: no CAS retries and
// no bank conflicts for small G (at G = 100 the un-replicated CAS loop is 2.6x slower than at
// G = 1e4).  The launcher picks the largest replication that fits shared memory.
template <typename VT, typename CT, int OP, int THREADS, int RSHIFT>
__global__ void __launch_bounds__(THREADS, 1)
k_moments_smem(const __grid_constant__ ChunkTable tab, int G,
               typename Acc<VT>::type* __restrict__ gacc, unsigned long long* __restrict__ gcnt,
               unsigned long long* __restrict__ gnan) {
  using AccT = typename Acc<VT>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using ST = AccT;  // one exact 64-bit slot (f64 / i64) per group copy
  constexpr int R = 1 << RSHIFT;
  const int S = G << RSHIFT;  // slots
  ST* sacc = reinterpret_cast<ST*>(smem_raw);
  unsigned* scnt = reinterpret_cast<unsigned*>(sacc + S);
  unsigned* snan = scnt + S;  // allocated only for float VT
  const int copy = threadIdx.x & (R - 1);

  const ST ident = (OP == FLOXB_OP_PROD) ? (ST)1 : (ST)0;
  for (int s = threadIdx.x; s < S; s += THREADS) {
    sacc[s] = ident;
    scnt[s] = 0u;
    if constexpr (is_float<VT>::value) snan[s] = 0u;
  }
  __syncthreads();

  auto accum = [&](VT v, CT c) {
    if ((unsigned long long)c < (unsigned long long)G) {
      const int s = ((int)c << RSHIFT) | copy;
      if (is_nan(v)) {
        if constexpr (is_float<VT>::value) atomicAdd(&snan[s], 1u);
      } else {
        acc_apply_shared<AccT, OP>(&sacc[s], lift<VT, OP>(v));
        atomicAdd(&scnt[s], 1u);
      }
    }
  };

  walk_tiles<VT, CT, THREADS>(tab, accum, [] {});
  __syncthreads();

  for (int g = threadIdx.x; g < G; g += THREADS) {
    AccT tot = (OP == FLOXB_OP_PROD) ? (AccT)1 : (AccT)0;
    unsigned long long c = 0ull, m = 0ull;
#pragma unroll 4
    for (int r = 0; r < R; ++r) {
      const int s = (g << RSHIFT) | ((r + threadIdx.x) & (R - 1));  // rotate: conflict-free across the warp
      const unsigned cs = scnt[s];
      if (cs) {
        if constexpr (OP == FLOXB_OP_PROD) tot *= (AccT)sacc[s]; else tot += (AccT)sacc[s];
        c += cs;
      }
      if constexpr (is_float<VT>::value) m += snan[s];
    }
    if (c) {
      if constexpr (OP == FLOXB_OP_PROD) atomic_mul(&gacc[g], tot);
      else atomic_add(&gacc[g], tot);
      atomicAdd(&gcnt[g], c);
    }
    if (m) atomicAdd(&gnan[g], m);
  }
}

// ------------------------------------------------------------------------------------------
// two-level shared-memory bins for f32 sums (the headline kernel)
// ------------------------------------------------------------------------------------------
// Level 1: f32 partial per slot, updated with the 32-bit CAS loop (LDS + FADD + ATOMS.CAST.SPIN: ~10
//          wavefronts per 32 random rows instead of ~21 for a 64-bit slot) + a native u32 counter
//          (ATOMS.ADD; the returned old count costs nothing extra — tools/kbench.py v24 vs v9).
// Level 2: f64 accumulator per slot.  Level 1 moves into level 2
//          (a) in a barrier-free SWEEP: every kSweepTiles tiles each thread atomicExch-es its own slots
//              (conflict-free) and adds them to level 2; for evenly spread groups a partial then holds
//              ~kSweepTiles * TILE / G addends (26 at G = 1e4);
//          (b) immediately, by the thread whose counter value shows that the slot just completed another
//              kDrainEvery adds (atomicExch of the partial + one 64-bit CAS), so hot or bursty groups
//              (skewed or sorted labels) never keep more than kDrainEvery (+ racing) addends in f32.
// Worst case |error| <= kDrainEvery * 2^-24 * sum|x| per group; measured ~1e-9..1e-8 relative, because the
// many short partials round independently (tests/test_gpu_parity.py::test_f32_two_level_accuracy_under_skew)
// — tighter than the reference's own f32 pairwise reduceat (1-2e-7).
// NaN rows only set a per-group flag bit (nnan becomes "number of CTAs that saw a NaN in the group":
// > 0 iff the group has NaNs, which is all floxb_finalize consults).
// The tile loop is software-pipelined (loads of tile t+1 in flight during the atomics of tile t) and walks
// the chunk table monotonically (no per-tile binary search): tools/kbench.py v22 vs v9.
constexpr int kSweepTiles = 64;  // upper bound, tiles of THREADS*4 rows: ~26 adds per slot between sweeps at G = 1e4
constexpr unsigned kDrainEvery = 32;  // <= this many (+ racing) addends per f32 partial, whatever the skew

// WS: level 2 lives in a caller-provided, CTA-private slice of GLOBAL memory (L2-resident: 148 x G x 8 B)
// instead of shared memory.  It is only touched by the sweeps (coalesced RED.ADD.F64 / RED.ADD.U32) and the rare
// hot-slot drains, and it shrinks the kernel's shared-memory footprint from 16 to 8 bytes per slot — which
// matters because the L1 that backs the streaming loads is whatever shared memory leaves over
// (tools/kbench.py: the same loop runs 26% slower with 200 KB than with 120 KB of dynamic shared memory).
template <typename CT, int OP, int THREADS, int RSHIFT, bool WS>
__global__ void __launch_bounds__(THREADS, 1)
k_moments_smem_f32(const __grid_constant__ ChunkTable tab, int G, double* __restrict__ gacc,
                   unsigned long long* __restrict__ gcnt, unsigned long long* __restrict__ gnan,
                   double* __restrict__ workspace) {
  constexpr int R = 1 << RSHIFT;
  const int S = G << RSHIFT;
  const int NW = (G + 31) >> 5;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // level 2 (sum f64 + adds before the last sweep u32): workspace slice [S f64][S u32] or shared memory
  const size_t ws_stride = (size_t)S * 8 + (((size_t)S * 4 + 7) & ~(size_t)7);  // keep every slice 8-byte aligned
  double* sacc = WS ? reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(workspace) + (size_t)blockIdx.x * ws_stride)
                    : reinterpret_cast<double*>(smem_raw);
  unsigned* stot = WS ? reinterpret_cast<unsigned*>(sacc + S) : reinterpret_cast<unsigned*>(smem_raw + (size_t)S * 8);
  float* spart = reinterpret_cast<float*>(smem_raw + (WS ? 0 : (size_t)S * 12));  // [S] level 1 partial
  unsigned* scnt = reinterpret_cast<unsigned*>(spart + S);   // [S] adds since the last sweep
  unsigned* snanbits = scnt + S;                              // [NW] group-has-NaN flags
  const int copy = threadIdx.x & (R - 1);
  for (int s = threadIdx.x; s < S; s += THREADS) {
    sacc[s] = 0.0;
    stot[s] = 0u;
    spart[s] = 0.0f;
    scnt[s] = 0u;
  }
  for (int w = threadIdx.x; w < NW; w += THREADS) snanbits[w] = 0u;
  __syncthreads();

  auto add_row = [&](float v, CT c) {
    if ((unsigned long long)c >= (unsigned long long)G) return;
    if (v != v) {
      const unsigned bit = 1u << ((int)c & 31);
      if (!(snanbits[(int)c >> 5] & bit)) atomicOr(&snanbits[(int)c >> 5], bit);
      return;
    }
    const int sl = ((int)c << RSHIFT) | copy;
    const float w = (OP == FLOXB_OP_SUMSQ) ? v * v : v;
    atomicAdd(&spart[sl], w);
    const unsigned seen = atomicAdd(&scnt[sl], 1u);
    if (seen >= kDrainEvery - 1) {
      // hot slot (kDrainEvery adds since the last sweep): hand partial and counter to level 2 now.  atomicExch
      // gives each to exactly one thread; concurrent adds land in the fresh zeros.  Never taken for evenly
      // spread groups (~10 adds per slot between sweeps), every kDrainEvery-th add for a hot group.
      const float part = atomicExch(&spart[sl], 0.0f);
      atomicAdd(&sacc[sl], (double)part);
      atomicAdd(&stot[sl], atomicExch(&scnt[sl], 0u));
    }
  };

  // Barrier-free sweep: every thread hands ITS slots to level 2 with atomicExch every kSweepTiles of its own
  // tiles.  A row racing with the sweep either lands before the exchange (moved now) or in the fresh zero
  // (moved by the next sweep); its count may travel one sweep later than its value, which only matters to the
  // drain heuristic.  (tools/kbench3.py: __syncthreads-fenced sweeps cost 7 points of roofline, these 2.)
  // sweep interval: ~24 adds per slot between sweeps for evenly spread groups (64 tiles at S = 1e4; fewer for the
  // replicated small-G layouts, whose slots would otherwise fill up and take the drain path every 32 adds)
  const int sweep_tiles = max(4, min(kSweepTiles, (int)(((int64_t)S * 24) / (THREADS * 4))));
  int since_sweep = 0;
  walk_tiles<float, CT, THREADS>(tab, add_row, [&] {
    if (++since_sweep == sweep_tiles) {
      since_sweep = 0;
      for (int sl = threadIdx.x; sl < S; sl += THREADS) {
        const unsigned c = atomicExch(&scnt[sl], 0u);
        if (c) {
          const float p = atomicExch(&spart[sl], 0.0f);
          atomicAdd(&sacc[sl], (double)p);
          atomicAdd(&stot[sl], c);
        }
      }
    }
  });
  __syncthreads();

  for (int g = threadIdx.x; g < G; g += THREADS) {
    double tot = 0.0;
    unsigned long long c = 0ull;
#pragma unroll 4
    for (int r = 0; r < R; ++r) {
      const int s = (g << RSHIFT) | ((r + threadIdx.x) & (R - 1));  // rotated: conflict-free across the warp
      const unsigned cs = scnt[s] + (WS ? __ldcg(&stot[s]) : stot[s]);  // level 2 was written by L2 atomics: bypass L1
      if (cs) {
        tot += (WS ? __ldcg(&sacc[s]) : sacc[s]) + (double)spart[s];
        c += cs;
      }
    }
    if (c) {
      atomicAdd(&gacc[g], tot);
      atomicAdd(&gcnt[g], c);
    }
    if (snanbits[g >> 5] & (1u << (g & 31))) atomicAdd(&gnan[g], 1ull);
  }
}

// ------------------------------------------------------------------------------------------
// fixed-point level 1 for f32 sums: TWO NATIVE integer atomics per row, no CAS loop
// ------------------------------------------------------------------------------------------
// The two-level kernel above is bound by shared-memory wavefronts, not by HBM: an f32 add to shared memory is an
// LDS + ATOMS.CAST.SPIN loop (a CAS costs two wavefronts per conflict way) next to the native counter add
// (profiles/r01_ncu_moments_twolevel_leanwalk_summary.json: 16.3 wavefronts per 32 rows, L1TEX 96% busy).  The only
// native shared-memory add on sm_100a is the 32-bit integer ATOMS.ADD, so this kernel makes level 1 an INTEGER:
//
//   slot = { lo : u32,  hi : u32 = [count:13 | sum bits 32..50:19] }        (8 bytes per slot, as before)
//
// A row adds the 64-bit fixed-point number  a' = 2^37 +- m * 2^sh  (m = the 24-bit significand, sh = the row's
// exponent relative to a per-CTA base e_lo, 0 <= sh <= 13; the 2^37 bias keeps every addend non-negative and is
// removed again through the count) with
//      old = ATOMS.ADD(lo, a'_lo);   ATOMS.ADD(hi, 2^19 + a'_hi + carry(old, a'_lo))
// Both are native and RETURNING: the first return value yields the carry, the second the slot's add count, which
// triggers the hot-slot drain.  Every row contributes its full amount to whichever sweep picks up each half, and a
// carry is produced by exactly the add that wrapped `lo`, so the 64-bit total is exact under any interleaving of
// adds, sweeps and drains (a wrap and its carry may land in different sweeps: the two swept values are then off by
// -+2^32 units and cancel in level 2).  The conversion is EXACT or not taken: a value whose bits do not fit the
// window (|v| >= 2^(e_lo+14-127), or low significand bits below 2^(e_lo-150) set; infinities) is added to the
// global f64 accumulator directly.  e_lo comes from the largest exponent in a 1024-value sample of the launch's data
// (+1 binade of headroom), so for data within ~4 decades of its maximum (uniform, normal, physical fields) almost
// every row takes the two-atomic path: 2 x ~3.5 conflict wavefronts per 32 rows instead of 16.
// Field budget: 24 (significand) + 1 (bias) + 13 (window) + 13 (adds) = 51 sum bits; a slot is drained when its count
// reaches kFxDrain and at most THREADS*4 more adds can be in flight, so count <= 1024 + 4096 < 2^13.
// Level 2, sweeps, NaN flags and the final merge are those of k_moments_smem_f32; sums are exact integers until they are
// rounded ONCE per sweep into the f64 level 2 (|value| < 2^51: exact conversion).
constexpr int kFxWindow = 13;                       // accepted shifts 0..13
constexpr unsigned long long kFxBias = 1ull << (24 + kFxWindow);
constexpr int kFxCountShift = 19;                   // hi = [count:13 | sum:19]
constexpr unsigned kFxDrain = 1024;

__device__ __forceinline__ long long fx_value(unsigned h, unsigned l) {
  return (long long)(((unsigned long long)(h & ((1u << kFxCountShift) - 1u)) << 32) | (unsigned long long)l) -
         (long long)(h >> kFxCountShift) * (long long)kFxBias;
}

template <typename CT, int THREADS, int RSHIFT, bool WS>
__global__ void __launch_bounds__(THREADS, 1)
k_moments_smem_fx(const __grid_constant__ ChunkTable tab, int G, double* __restrict__ gacc,
                  unsigned long long* __restrict__ gcnt, unsigned long long* __restrict__ gnan,
                  double* __restrict__ workspace) {
  constexpr int R = 1 << RSHIFT;
  const int S = G << RSHIFT;
  const int NW = (G + 31) >> 5;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const size_t ws_stride = (size_t)S * 8 + (((size_t)S * 4 + 7) & ~(size_t)7);
  double* sacc = WS ? reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(workspace) + (size_t)blockIdx.x * ws_stride)
                    : reinterpret_cast<double*>(smem_raw);
  unsigned* stot = WS ? reinterpret_cast<unsigned*>(sacc + S) : reinterpret_cast<unsigned*>(smem_raw + (size_t)S * 8);
  unsigned* slo = reinterpret_cast<unsigned*>(smem_raw + (WS ? 0 : (size_t)S * 12));  // [S]
  unsigned* shi = slo + S;                                                             // [S]
  unsigned* snanbits = shi + S;                                                        // [NW]
  int& s_emax = *reinterpret_cast<int*>(snanbits + NW);  // dynamic too: static + dynamic must fit the 227 KB opt-in
  const int copy = threadIdx.x & (R - 1);
  if (threadIdx.x == 0) s_emax = 0;
  for (int s = threadIdx.x; s < S; s += THREADS) {
    sacc[s] = 0.0;
    stot[s] = 0u;
    slo[s] = 0u;
    shi[s] = 0u;
  }
  for (int w = threadIdx.x; w < NW; w += THREADS) snanbits[w] = 0u;
  __syncthreads();
  {  // exponent base: the largest finite exponent among THREADS values sampled across the launch's chunks
    int e = 0;
    const unsigned h = ((unsigned)blockIdx.x * THREADS + threadIdx.x) * 2654435761u;
    int k = (int)(h % (unsigned)tab.nchunks);
    for (int t = 0; t < 4 && tab.n[k] == 0; ++t) k = (k + 1) % tab.nchunks;
    if (tab.n[k] > 0) {
      const unsigned long long r = ((unsigned long long)h * 0x9E3779B97F4A7C15ull) >> 20;
      const unsigned b = __float_as_uint(reinterpret_cast<const float*>(tab.values[k])[r % (unsigned long long)tab.n[k]]);
      e = (int)((b >> 23) & 0xffu);
      if (e == 0xff) e = 0;
    }
    e = __reduce_max_sync(0xffffffffu, e);
    if ((threadIdx.x & 31) == 0 && e > 0) atomicMax(&s_emax, e);
  }
  __syncthreads();
  const int e_lo = max(1, min(s_emax + 1, 254) - kFxWindow);
  const double scale = __longlong_as_double((long long)(e_lo - 150 + 1023) << 52);  // 2^(e_lo - 150): one fixed-point unit
  // float -> fixed point through the f64 "magic constant": C = 1.5 * 2^52 units.  d = v + C rounds v to a whole number of
  // units and leaves that integer in d's low significand bits; (d - C) == v says that nothing was rounded away (v is an
  // exact multiple of the unit), and the biased integer a' = bits(d) - bits(C) + 2^37 must fall in [0, 2^38) (|v| inside
  // the window).  NaN fails the comparison, infinities and huge values the range test: they take the slow path.
  const unsigned long long cbits = ((unsigned long long)(e_lo + 925) << 52) | (1ull << 51);
  const double C = __longlong_as_double((long long)cbits);
  const unsigned long long kc = cbits - kFxBias;

  auto to_level2 = [&](int sl, unsigned h, unsigned l) {
    atomicAdd(&sacc[sl], (double)fx_value(h, l) * scale);
    atomicAdd(&stot[sl], h >> kFxCountShift);
  };
  auto drain = [&](int sl) {
    const unsigned h = atomicExch(&shi[sl], 0u);
    if (h) to_level2(sl, h, atomicExch(&slo[sl], 0u));
  };
  auto slow_row = [&](float v, int c) {  // outside the exact window (rare): NaN -> flag, else straight to the global f64 bins
    if (v != v) {
      const unsigned bit = 1u << (c & 31);
      if (!(snanbits[c >> 5] & bit)) atomicOr(&snanbits[c >> 5], bit);
    } else {
      atomicAdd(&gacc[c], (double)v);
      atomicAdd(&gcnt[c], 1ull);
    }
  };
  // state: 0 = no group, 1 = two native atomics, 2 = slow path
  auto convert = [&](float v, CT c, unsigned& lo, unsigned& hi) -> int {
    const double vd = (double)v;
    const double d = vd + C;
    const unsigned long long ap = (unsigned long long)__double_as_longlong(d) - kc;
    lo = (unsigned)ap;
    hi = (unsigned)(ap >> 32);
    const bool fast = ((d - C) == vd) && hi < (1u << (24 + kFxWindow + 1 - 32));
    return ((unsigned long long)c < (unsigned long long)G) ? (fast ? 1 : 2) : 0;
  };
  auto add_row = [&](float v, CT c) {
    unsigned lo, hi;
    const int st = convert(v, c, lo, hi);
    if (st == 1) {
      const int sl = ((int)c << RSHIFT) | copy;
      const unsigned old = atomicAdd(&slo[sl], lo);
      const unsigned oh = atomicAdd(&shi[sl], (1u << kFxCountShift) + hi + (unsigned)((old + lo) < lo));
      if ((oh >> kFxCountShift) >= kFxDrain - 1) drain(sl);
    } else if (st == 2) {
      slow_row(v, (int)c);
    }
  };
  // The four rows of a tile.  ncu on the first version of this kernel: 84% of the issue slots busy, 68 instructions per
  // row — the kernel had become ISSUE-bound (generic->shared address conversion in front of every atomic, a
  // BSSY/BRA/BSYNC triple around each, predicates spilled into a mask register).  So: when all 128 rows of the warp
  // are ordinary (valid code, value inside the exact window — practically always), a straight-line path issues the
  // eight native atomics through 32-bit shared-memory addresses with no branch in between; anything else takes the
  // row-by-row path above.
  const unsigned slo_s = (unsigned)__cvta_generic_to_shared(slo);
  const unsigned hi_off = (unsigned)S * 4u;
  const unsigned copy4 = (unsigned)copy * 4u;
  auto atoms_add = [](unsigned addr, unsigned v) -> unsigned {
    unsigned old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
    return old;
  };
  auto add_tile = [&](const Vec4<float>& v, const Vec4<CT>& c) {
    unsigned lo[4], hi[4];
    bool all_fast = true;
#pragma unroll
    for (int k = 0; k < 4; ++k) all_fast &= convert(v.v[k], c.v[k], lo[k], hi[k]) == 1;
    if (__all_sync(0xffffffffu, all_fast)) {
      unsigned la[4], old[4], oh[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) la[k] = slo_s + ((unsigned)c.v[k] << (2 + RSHIFT)) + copy4;
#pragma unroll
      for (int k = 0; k < 4; ++k) old[k] = atoms_add(la[k], lo[k]);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        oh[k] = atoms_add(la[k] + hi_off, (1u << kFxCountShift) + hi[k] + (unsigned)((old[k] + lo[k]) < lo[k]));
      if ((oh[0] | oh[1] | oh[2] | oh[3]) >= ((kFxDrain - 1) << kFxCountShift)) {  // rare: some slot may be full
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if ((oh[k] >> kFxCountShift) >= kFxDrain - 1) drain(((int)c.v[k] << RSHIFT) | copy);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) add_row(v.v[k], c.v[k]);
    }
  };

  const int sweep_tiles = max(4, min(kSweepTiles, (int)(((int64_t)S * 24) / (THREADS * 4))));
  int since_sweep = 0;
  walk_tiles4<float, CT, THREADS>(tab, add_tile, add_row, [&] {
    if (++since_sweep == sweep_tiles) {
      since_sweep = 0;
      for (int sl = threadIdx.x; sl < S; sl += THREADS) {
        const unsigned h = atomicExch(&shi[sl], 0u);
        if (h) to_level2(sl, h, atomicExch(&slo[sl], 0u));
      }
    }
  });
  __syncthreads();

  for (int g = threadIdx.x; g < G; g += THREADS) {
    double tot = 0.0;
    unsigned long long c = 0ull;
#pragma unroll 4
    for (int r = 0; r < R; ++r) {
      const int s = (g << RSHIFT) | ((r + threadIdx.x) & (R - 1));
      const unsigned h = shi[s];
      const unsigned c2 = WS ? __ldcg(&stot[s]) : stot[s];
      if (h | c2) {
        tot += (WS ? __ldcg(&sacc[s]) : sacc[s]) + (double)fx_value(h, slo[s]) * scale;
        c += (unsigned long long)(h >> kFxCountShift) + c2;
      }
    }
    if (c) {
      atomicAdd(&gacc[g], tot);
      atomicAdd(&gcnt[g], c);
    }
    if (snanbits[g >> 5] & (1u << (g & 31))) atomicAdd(&gnan[g], 1ull);
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
  for (int64_t a = blockIdx.y; a < A; a += gridDim.y) {
    walk_flat<4>(vals + a * row_len, codes, row_len, B, G, [&](VT v, CT c, int64_t, int64_t b) {
      // kept trailing dims: accumulators are GROUP-MAJOR [G][A*B], so the B values of one row land in one
      // contiguous run (tools/kbench2.py: 2.6x the L2 atomic rate of the [A*B][G] layout at G = 1e5)
      const int64_t bin = (B == 1) ? a * G + (int64_t)c : (int64_t)c * (A * B) + a * B + b;
      if (is_nan(v)) {
        atomicAdd(&gnan[bin], 1ull);
      } else {
        if constexpr (OP == FLOXB_OP_PROD) atomic_mul(&gacc[bin], lift<VT, OP>(v));
        else atomic_add(&gacc[bin], lift<VT, OP>(v));
        atomicAdd(&gcnt[bin], 1ull);
      }
    });
  }
}

// L2-atomic path specialised for B == 1: 4 rows per thread (128-bit loads), adjacent equal codes are
// merged in registers, then RUNS of lanes that end in the same code are merged with a segmented shuffle
// reduction and only the first lane of a run issues atomics.  For sorted labels (BASELINE config 5,
// method="cohorts") a warp's 128 rows hold one or two groups: the atomic traffic drops by ~100x and the
// kernel is an HBM stream; for random labels it degrades to one atomic set per row like k_moments_global.
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
    // two vectors per thread per step, both loaded before either is merged: twice the loads in flight
    auto merge_vec = [&](const Vec4<VT>& v, const Vec4<CT>& c) {
        long long cur = (long long)c.v[0];
        AccT s = ident;
        unsigned cn = 0, mn = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const long long ck = (long long)c.v[k];
          if (ck != cur) {
            flush(cur, s, cn, mn);
            cur = ck; s = ident; cn = 0; mn = 0;
          }
          if (is_nan(v.v[k])) mn++; else { s = merge(s, lift<VT, OP>(v.v[k])); cn++; }
        }
        // Segmented warp reduction over RUNS of lanes that end in the same group: one flush per run instead of
        // one per lane.  With sorted labels a warp's 128 rows hold 1-2 groups, so it issues 1-2 sets of REDs; the
        // earlier "whole warp uniform, else every lane flushes" left a quarter of the warps sending 32 REDs each
        // to the same two addresses, and the kernel swung between 1.65 and 4.1 ms on same-address L2 serialisation.
        // Random labels: every lane is its own run, nothing changes.
        const unsigned lane = threadIdx.x & 31u;
        const long long kprev = __shfl_up_sync(0xffffffffu, cur, 1);
        const unsigned heads = __ballot_sync(0xffffffffu, lane == 0u || kprev != cur);
        const unsigned after = (heads >> lane) >> 1;  // head flags of lanes lane+1, lane+2, ...
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          AccT s2;
          if constexpr (sizeof(AccT) == 8) {
            const long long o = __shfl_down_sync(0xffffffffu, *reinterpret_cast<long long*>(&s), d);
            s2 = *reinterpret_cast<const AccT*>(&o);
          }
          const unsigned cn2 = __shfl_down_sync(0xffffffffu, cn, d), mn2 = __shfl_down_sync(0xffffffffu, mn, d);
          // lane+d belongs to this lane's run iff no run starts in (lane, lane+d]
          if (lane + d < 32u && (after & ((1u << d) - 1u)) == 0u) {
            s = merge(s, s2);
            cn += cn2;
            mn += mn2;
          }
        }
        if ((heads >> lane) & 1u) flush(cur, s, cn, mn);
    };
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nvec_full; q += 2 * stride) {
      const bool two = q + stride < nvec_full;  // warp-uniform: nvec_full and stride are multiples of 32
      const Vec4<VT> v0 = load4<VT>(row + q * 4);
      const Vec4<CT> c0 = load4<CT>(codes + q * 4);
      Vec4<VT> v1;
      Vec4<CT> c1;
      if (two) {
        v1 = load4<VT>(row + (q + stride) * 4);
        c1 = load4<CT>(codes + (q + stride) * 4);
      }
      merge_vec(v0, c0);
      if (two) merge_vec(v1, c1);
    }
    // scalar remainder (unaligned chunks, n % 128 tail)
    for (int64_t i = nvec_full * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const VT v = row[i];
      if (is_nan(v)) flush((long long)codes[i], ident, 0u, 1u);
      else flush((long long)codes[i], lift<VT, OP>(v), 1u, 0u);
    }
  }
}

// L2-atomic path for a SMALL power-of-two number of kept trailing columns (BASELINE config 4: f64 (1e8, 8), G = 1e5,
// method="blockwise"): every thread keeps one column's running (sum, count, NaN count) in registers while the code does
// not change (walk_rows_small_b) and issues one RED set per RUN.  Blockwise / sorted labels: ~1 RED set per
// thread per 64-step segment instead of one per element, the kernel becomes an HBM stream; random labels: one set per
// element, like k_moments_global.  Accumulators are group-major [G][A*B] like every B > 1 kernel.
// ROWCNT: the labels are shared by every (a, b), so the number of ROWS of a group is counted once — by the thread of
// column 0 of a == 0, into that column's count slot — instead of once per element, and the per-(a, b) counts of non-NaN
// values are rebuilt afterwards as rows - nnan (k_cnt_from_rows).  With random labels (no runs to merge) that is 8 sum
// REDs + 1 row RED + the NaN REDs per row of 8 instead of 8 + 8 + NaN: the kernel is bound by the L2 atomic rate.
template <typename VT, typename CT, int OP, int B, bool ROWCNT>
__global__ void __launch_bounds__(256)
k_moments_global_bruns(const VT* __restrict__ vals, const CT* __restrict__ codes, int64_t A, int64_t n, int64_t G,
                       typename Acc<VT>::type* __restrict__ gacc, unsigned long long* __restrict__ gcnt,
                       unsigned long long* __restrict__ gnan) {
  using AccT = typename Acc<VT>::type;
  const AccT ident = (OP == FLOXB_OP_PROD) ? (AccT)1 : (AccT)0;
  const int col = threadIdx.x & (B - 1);
  for (int64_t a = blockIdx.y; a < A; a += gridDim.y) {
    const int64_t ab = a * B + col, AB = A * B;
    long long cur = -1;
    AccT s = ident;
    unsigned cn = 0, mn = 0;
    auto flush = [&] {
      if ((unsigned long long)cur < (unsigned long long)G) {
        const int64_t bin = (int64_t)cur * AB + ab;
        if (cn) {
          if constexpr (OP == FLOXB_OP_PROD) atomic_mul(&gacc[bin], s); else atomic_add(&gacc[bin], s);
          if constexpr (!ROWCNT) atomicAdd(&gcnt[bin], (unsigned long long)cn);
        }
        if constexpr (ROWCNT) {
          if (ab == 0 && (cn + mn)) atomicAdd(&gcnt[bin], (unsigned long long)(cn + mn));  // rows of the group in this run
        }
        if (mn) atomicAdd(&gnan[bin], (unsigned long long)mn);
      }
      s = ident; cn = 0; mn = 0;
    };
    walk_rows_small_b<B, 8, 64>(vals + a * n * B, codes, n,
      [&](VT v, CT c, int64_t) {
        if ((long long)c != cur) { flush(); cur = (long long)c; }
        if (is_nan(v)) ++mn;
        else {
          if constexpr (OP == FLOXB_OP_PROD) s *= lift<VT, OP>(v); else s += lift<VT, OP>(v);
          ++cn;
        }
      },
      [&] { flush(); cur = -1; });
  }
}

// rows <-> counts for the ROWCNT kernels (group-major bins [G][AB]): `to_rows` turns column 0's count into the group's row
// count before accumulating on top of earlier calls (rows = cnt + nnan), the second form rebuilds every count from it.
__global__ void k_cnt_to_rows(unsigned long long* __restrict__ cnt, const unsigned long long* __restrict__ nnan, int64_t G,
                              int64_t AB) {
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < G; g += (int64_t)gridDim.x * blockDim.x)
    cnt[g * AB] += nnan[g * AB];
}
__global__ void k_cnt_from_rows(unsigned long long* __restrict__ cnt, const unsigned long long* __restrict__ nnan, int64_t G,
                                int64_t AB) {
  const int64_t total = G * AB;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = i / AB, ab = i - g * AB;
    if (ab != 0) cnt[i] = cnt[g * AB] - nnan[i];   // cnt[g * AB] still holds the rows: column 0 is rewritten below
  }
}
__global__ void k_cnt_from_rows_col0(unsigned long long* __restrict__ cnt, const unsigned long long* __restrict__ nnan, int64_t G,
                                     int64_t AB) {
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < G; g += (int64_t)gridDim.x * blockDim.x)
    cnt[g * AB] -= nnan[g * AB];
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
constexpr size_t kSmemBudget = 227 * 1024;

template <typename VT, int OP> size_t smem_bytes_moments(int64_t G) {
  return (size_t)G * (sizeof(typename Acc<VT>::type) + 4 + (is_float<VT>::value ? 4 : 0));
}

template <typename VT, typename CT, int OP>
int launch_moments(const floxb_chunk_t* chunks, int64_t n_chunks, int64_t A, int64_t B, int64_t G, void* acc,
                   int64_t* cnt, int64_t* nnan, int init, int force_path, void* workspace, int64_t workspace_bytes,
                   cudaStream_t st) {
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
  constexpr bool kTwoLevel = std::is_same<VT, float>::value && OP != FLOXB_OP_PROD;
  const bool use_ws = kTwoLevel && workspace != nullptr && workspace_bytes >= (int64_t)sm_count() * (G * 12 + 8);
  const size_t smem = kTwoLevel ? (size_t)G * (use_ws ? 8 : 20) + (size_t)((G + 31) / 32) * 4 : smem_bytes_moments<VT, OP>(G);
  bool use_smem = (A == 1 && B == 1 && smem + (kTwoLevel ? 16 : 0) <= kSmemBudget);
  if ((force_path == 1 || force_path == 3) && !use_smem) return FLOXB_EINVAL;
  if (force_path == 2) use_smem = false;

  if (use_smem) {
    constexpr int TILE = kSmemThreads * 4;
    using KernT = void (*)(const ChunkTable, int, AccT*, unsigned long long*, unsigned long long*);
    using KernWsT = void (*)(const ChunkTable, int, double*, unsigned long long*, unsigned long long*, double*);
    KernT kern = nullptr;
    KernWsT kern_ws = nullptr;
    size_t smem_used = smem;
    if constexpr (kTwoLevel) {
      auto bytes = [&](int r) { return (size_t)G * r * 20 + (size_t)((G + 31) / 32) * 4; };
      // f32 sums: fixed-point level 1 (two native atomics per row); force_path 3 = the CAS-loop level 1 (sums of
      // squares always use it: a squared f32 has a 48-bit significand)
      constexpr bool kFx = OP == FLOXB_OP_SUM;
      const bool fx = kFx && force_path != 3;
      if (bytes(32) + 16 <= kSmemBudget) {
        kern_ws = k_moments_smem_f32<CT, OP, kSmemThreads, 5, false>;
        if constexpr (kFx) if (fx) kern_ws = k_moments_smem_fx<CT, kSmemThreads, 5, false>;
        smem_used = bytes(32);
      } else if (bytes(8) + 16 <= kSmemBudget) {
        kern_ws = k_moments_smem_f32<CT, OP, kSmemThreads, 3, false>;
        if constexpr (kFx) if (fx) kern_ws = k_moments_smem_fx<CT, kSmemThreads, 3, false>;
        smem_used = bytes(8);
      } else if (use_ws) {
        kern_ws = k_moments_smem_f32<CT, OP, kSmemThreads, 0, true>;
        if constexpr (kFx) if (fx) kern_ws = k_moments_smem_fx<CT, kSmemThreads, 0, true>;
      } else {
        kern_ws = k_moments_smem_f32<CT, OP, kSmemThreads, 0, false>;
        if constexpr (kFx) if (fx) kern_ws = k_moments_smem_fx<CT, kSmemThreads, 0, false>;
      }
    } else {
      kern = k_moments_smem<VT, CT, OP, kSmemThreads, 0>;
      if (smem * 32 <= kSmemBudget) { kern = k_moments_smem<VT, CT, OP, kSmemThreads, 5>; smem_used = smem * 32; }
      else if (smem * 8 <= kSmemBudget) { kern = k_moments_smem<VT, CT, OP, kSmemThreads, 3>; smem_used = smem * 8; }
    }
    if (cudaFuncSetAttribute(kern_ws ? (const void*)kern_ws : (const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)kSmemBudget) != cudaSuccess)
      return FLOXB_ECUDA;
    for (int64_t c0 = 0; c0 < n_chunks; c0 += kMaxChunksPerLaunch) {
      ChunkTable tab;
      const int64_t tiles = fill_chunk_table(chunks, n_chunks, c0, TILE, &tab);
      if (tiles < 0) return (int)tiles;
      if (tiles == 0) continue;
      // u32 per-CTA counters: keep rows per CTA below 2^32
      if (tiles / sm_count() * TILE >= (int64_t)1 << 32) return FLOXB_ERANGE;
      const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
      if (kern_ws) {
        if constexpr (kTwoLevel)
          kern_ws<<<grid, kSmemThreads, smem_used + 16, st>>>(tab, (int)G, gacc, gcnt, gnan, reinterpret_cast<double*>(workspace));
      } else {
        kern<<<grid, kSmemThreads, smem_used, st>>>(tab, (int)G, gacc, gcnt, gnan);
      }
    }
  } else {
    const bool bruns = B > 1 && B <= 32 && (B & (B - 1)) == 0 && force_path != 4;
    const bool rowcnt = bruns && force_path != 5;  // 5 = per-element counts (testing / profiling)
    const int rc_blocks = (int)((G + 255) / 256 < 1184 ? (G + 255) / 256 : 1184);
    if (rowcnt && !init) k_cnt_to_rows<<<rc_blocks, 256, 0, st>>>(gcnt, gnan, G, A * B);
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
      } else if (bruns) {
        // small power-of-two B: run-merging kernel, one warp per segment of 64 steps
        const int64_t segs = (ch.n * B + 64 * 32 - 1) / (64 * 32);
        int64_t gx = (segs + 7) / 8;  // 8 warps per block
        if (gx > cap) gx = cap;
        if (gx * by > cap * 4) gx = (cap * 4 + by - 1) / by;
        if (gx < 1) gx = 1;
        dim3 g2((unsigned)gx, (unsigned)by);
        const VT* v = reinterpret_cast<const VT*>(ch.values);
        const CT* c = reinterpret_cast<const CT*>(ch.codes);
#define FLOXB_BRUNS(BB)                                                                                                  \
  do {                                                                                                                   \
    if (rowcnt) k_moments_global_bruns<VT, CT, OP, BB, true><<<g2, 256, 0, st>>>(v, c, A, ch.n, G, gacc, gcnt, gnan);     \
    else k_moments_global_bruns<VT, CT, OP, BB, false><<<g2, 256, 0, st>>>(v, c, A, ch.n, G, gacc, gcnt, gnan);           \
  } while (0)
        switch (B) {
          case 2: FLOXB_BRUNS(2); break;
          case 4: FLOXB_BRUNS(4); break;
          case 8: FLOXB_BRUNS(8); break;
          case 16: FLOXB_BRUNS(16); break;
          default: FLOXB_BRUNS(32); break;
        }
#undef FLOXB_BRUNS
      } else {
        k_moments_global<VT, CT, OP><<<grid, 256, 0, st>>>(reinterpret_cast<const VT*>(ch.values),
                                                         reinterpret_cast<const CT*>(ch.codes), A, ch.n, B, G, gacc,
                                                         gcnt, gnan);
      }
    }
    if (rowcnt) {
      const int64_t total = G * A * B;
      const int fb = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
      k_cnt_from_rows<<<fb, 256, 0, st>>>(gcnt, gnan, G, A * B);
      k_cnt_from_rows_col0<<<rc_blocks, 256, 0, st>>>(gcnt, gnan, G, A * B);
    }
  }
  return cudaGetLastError() == cudaSuccess ? FLOXB_OK : FLOXB_ECUDA;
}

template <typename VT, typename CT>
int dispatch_op(int op, const floxb_chunk_t* chunks, int64_t n_chunks, int64_t A, int64_t B, int64_t G, void* acc,
                int64_t* cnt, int64_t* nnan, int init, int force_path, void* workspace, int64_t workspace_bytes,
                cudaStream_t st) {
  switch (op) {
    case FLOXB_OP_SUM: return launch_moments<VT, CT, FLOXB_OP_SUM>(chunks, n_chunks, A, B, G, acc, cnt, nnan, init, force_path, workspace, workspace_bytes, st);
    case FLOXB_OP_PROD: return launch_moments<VT, CT, FLOXB_OP_PROD>(chunks, n_chunks, A, B, G, acc, cnt, nnan, init, force_path, workspace, workspace_bytes, st);
    case FLOXB_OP_SUMSQ: return launch_moments<VT, CT, FLOXB_OP_SUMSQ>(chunks, n_chunks, A, B, G, acc, cnt, nnan, init, force_path, workspace, workspace_bytes, st);
  }
  return FLOXB_EINVAL;
}

template <typename VT>
int dispatch_code(int code_dtype, int op, const floxb_chunk_t* chunks, int64_t n_chunks, int64_t A, int64_t B,
                  int64_t G, void* acc, int64_t* cnt, int64_t* nnan, int init, int force_path, void* workspace,
                  int64_t workspace_bytes, cudaStream_t st) {
  if (code_dtype == FLOXB_I64) return dispatch_op<VT, int64_t>(op, chunks, n_chunks, A, B, G, acc, cnt, nnan, init, force_path, workspace, workspace_bytes, st);
  if (code_dtype == FLOXB_I32) return dispatch_op<VT, int32_t>(op, chunks, n_chunks, A, B, G, acc, cnt, nnan, init, force_path, workspace, workspace_bytes, st);
  return FLOXB_EDTYPE;
}

}  // namespace floxb

using namespace floxb;

extern "C" int floxb_moments(const floxb_chunk_t* chunks, int64_t n_chunks, int value_dtype, int code_dtype, int64_t A,
                             int64_t B, int64_t n_groups, int op, void* acc, int64_t* cnt, int64_t* nnan, int init,
                             int force_path, void* workspace, int64_t workspace_bytes, void* stream) {
  if (n_chunks < 0 || A < 0 || B < 0 || n_groups < 0 || (n_chunks > 0 && !chunks)) return FLOXB_EINVAL;
  if (A * B * n_groups > 0 && (!acc || !cnt || !nnan)) return FLOXB_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (value_dtype) {
    case FLOXB_F32: return dispatch_code<float>(code_dtype, op, chunks, n_chunks, A, B, n_groups, acc, cnt, nnan, init, force_path, workspace, workspace_bytes, st);
    case FLOXB_F64: return dispatch_code<double>(code_dtype, op, chunks, n_chunks, A, B, n_groups, acc, cnt, nnan, init, force_path, workspace, workspace_bytes, st);
    case FLOXB_I32: return dispatch_code<int32_t>(code_dtype, op, chunks, n_chunks, A, B, n_groups, acc, cnt, nnan, init, force_path, workspace, workspace_bytes, st);
    case FLOXB_I64: return dispatch_code<int64_t>(code_dtype, op, chunks, n_chunks, A, B, n_groups, acc, cnt, nnan, init, force_path, workspace, workspace_bytes, st);
  }
  return FLOXB_EDTYPE;
}

extern "C" int64_t floxb_moments_workspace_bytes(int value_dtype, int64_t A, int64_t B, int64_t n_groups) {
  // only the two-level f32 kernel with one slot copy per group (G too large for replicated slots) uses it
  if (value_dtype != FLOXB_F32 || A != 1 || B != 1) return 0;
  if ((size_t)n_groups * 8 * 20 + 64 <= kSmemBudget) return 0;          // replicated slots: level 2 stays in smem
  if ((size_t)n_groups * 8 + (size_t)((n_groups + 31) / 32) * 4 > kSmemBudget) return 0;  // L2-atomic path
  return (int64_t)sm_count() * (n_groups * 12 + 8);
}

extern "C" int floxb_max_chunks_per_launch(void) { return floxb::kMaxChunksPerLaunch; }

extern "C" int64_t floxb_smem_group_limit(int value_dtype, int family) {
  // family 0: f32 sums use the two-level kernel (20 B / group), everything else one 64-bit slot
  const size_t per = (family == 0) ? (value_dtype == FLOXB_F32 ? 9 : (8 + 4 + (value_dtype == FLOXB_F64 ? 4 : 0)))
                                   : (((value_dtype == FLOXB_F32 || value_dtype == FLOXB_I32) ? 4 : 8) + 4 + 4);
  return (int64_t)(kSmemBudget / per);
}
