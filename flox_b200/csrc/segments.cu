// flox_b200 — the sorted-segment family: grouped quantile / median.
//
// Reference: flox/aggregate_flox.py:50-130 (quantile_), :218-221 (quantile / nanquantile / median / nanmedian): the rows
// are partitioned by (group, value) — NaNs behind everything within their group —, the linear-method quantile of the
// VALID values of each group is interpolated between two order statistics with numpy's `_lerp` (:26-47).
// Here the (group, value) ordering is produced by the caller (a radix sort of composite keys on the device,
// flox_b200/aggregate_b200.py); these kernels take the sorted rows and do the per-(row, group) work:
//
//   floxb_segment_valid     nvalid[r, g] = number of non-NaN values of group g in row r (a binary search for the first
//                           NaN of the segment: NaNs sort last)
//   floxb_segment_quantile  out[iq, r, g] from the order statistics.  Bit-for-bit the reference's arithmetic: its virtual
//                           index q * (n_valid - 1) carries the running count of valid values of the EARLIER groups of the
//                           row, added in float64 (aggregate_flox.py:85-90), which decides the last bits of the
//                           interpolation weight; products and sums below are individually rounded (__dmul_rn /
//                           __dadd_rn: no FMA contraction), the difference hi - lo is taken in the value dtype.
#include <math_constants.h>

#include "common.cuh"

namespace floxb {

template <typename VT>
__global__ void __launch_bounds__(256)
k_segment_valid(const VT* __restrict__ sorted, int64_t R, int64_t n, const long long* __restrict__ starts, int64_t G,
                long long* __restrict__ nvalid) {
  const int64_t total = R * G;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / G, g = t - r * G;
    const VT* row = sorted + r * n;
    long long lo = starts[g], hi = starts[g + 1];
    const long long s0 = lo;
    while (lo < hi) {  // first index in [lo, hi) holding a NaN
      const long long mid = (lo + hi) >> 1;
      if (is_nan(row[mid])) hi = mid; else lo = mid + 1;
    }
    nvalid[t] = lo - s0;
  }
}

template <typename VT, typename OT>
__global__ void __launch_bounds__(256)
k_segment_quantile(const VT* __restrict__ sorted, int64_t R, int64_t n, const long long* __restrict__ starts, int64_t G,
                   const long long* __restrict__ nvalid, const long long* __restrict__ before, const double* __restrict__ qs,
                   int nq, int skipna, OT fill, OT* __restrict__ out) {
  const int64_t total = R * G;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / G, g = t - r * G;
    const long long s0 = starts[g], len = starts[g + 1] - s0;
    const long long nv = nvalid[t], bef = before[t];
    const VT* seg = sorted + r * n + s0;
    for (int iq = 0; iq < nq; ++iq) {
      OT res;
      if (len == 0) {
        res = fill;                                  // the group does not occur: fill_value (aggregate_flox.py:174-190)
      } else if (nv == 0 || (!skipna && nv != len)) {
        res = (OT)CUDART_NAN;                        // all NaN, or a NaN under a NaN-propagating quantile (:126-129)
      } else {
        double virt = __dmul_rn(qs[iq], (double)(nv - 1));
        if (bef != 0) virt = __dadd_rn(virt, (double)bef);   // offset of the earlier groups' valid values (:88-90)
        const double fl = floor(virt), ce = ceil(virt);
        const double gamma = __dsub_rn(virt, fl);
        const VT a = seg[(long long)fl - bef], b = seg[(long long)ce - bef];
        const VT d = (VT)(b - a);                    // in the value dtype, like np.subtract(b, a)
        double v = __dadd_rn((double)a, __dmul_rn((double)d, gamma));
        if (gamma >= 0.5) v = __dsub_rn((double)b, __dmul_rn((double)d, __dsub_rn(1.0, gamma)));
        res = (OT)v;
      }
      out[((int64_t)iq * R + r) * G + g] = res;
    }
  }
}

static inline int seg_blocks(int64_t total) {
  int64_t b = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace floxb

using namespace floxb;

extern "C" int floxb_segment_valid(const void* sorted, int value_dtype, int64_t n_rows, int64_t n, const int64_t* starts,
                                   int64_t n_groups, int64_t* nvalid, void* stream) {
  if (n_rows < 0 || n < 0 || n_groups < 0) return FLOXB_EINVAL;
  if (n_rows * n_groups == 0) return FLOXB_OK;
  if (!sorted || !starts || !nvalid) return FLOXB_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int blocks = seg_blocks(n_rows * n_groups);
  const long long* s = reinterpret_cast<const long long*>(starts);
  long long* nv = reinterpret_cast<long long*>(nvalid);
  switch (value_dtype) {
    case FLOXB_F32: k_segment_valid<float><<<blocks, 256, 0, st>>>(reinterpret_cast<const float*>(sorted), n_rows, n, s, n_groups, nv); break;
    case FLOXB_F64: k_segment_valid<double><<<blocks, 256, 0, st>>>(reinterpret_cast<const double*>(sorted), n_rows, n, s, n_groups, nv); break;
    default: return FLOXB_EDTYPE;
  }
  return cudaGetLastError() == cudaSuccess ? FLOXB_OK : FLOXB_ECUDA;
}

extern "C" int floxb_segment_quantile(const void* sorted, int value_dtype, int64_t n_rows, int64_t n, const int64_t* starts,
                                      int64_t n_groups, const int64_t* nvalid, const int64_t* before, const double* q,
                                      int n_q, int skipna, double fill, int out_dtype, void* out, void* stream) {
  if (n_rows < 0 || n < 0 || n_groups < 0 || n_q < 0) return FLOXB_EINVAL;
  if (n_rows * n_groups * n_q == 0) return FLOXB_OK;
  if (!sorted || !starts || !nvalid || !before || !q || !out) return FLOXB_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int blocks = seg_blocks(n_rows * n_groups);
  const long long* s = reinterpret_cast<const long long*>(starts);
  const long long* nv = reinterpret_cast<const long long*>(nvalid);
  const long long* bf = reinterpret_cast<const long long*>(before);
#define FLOXB_SEGQ(VT, OT)                                                                                              \
  k_segment_quantile<VT, OT><<<blocks, 256, 0, st>>>(reinterpret_cast<const VT*>(sorted), n_rows, n, s, n_groups, nv, bf, q, \
                                                     n_q, skipna, (OT)fill, reinterpret_cast<OT*>(out))
  if (value_dtype == FLOXB_F32 && out_dtype == FLOXB_F32) FLOXB_SEGQ(float, float);
  else if (value_dtype == FLOXB_F32 && out_dtype == FLOXB_F64) FLOXB_SEGQ(float, double);
  else if (value_dtype == FLOXB_F64 && out_dtype == FLOXB_F64) FLOXB_SEGQ(double, double);
  else if (value_dtype == FLOXB_F64 && out_dtype == FLOXB_F32) FLOXB_SEGQ(double, float);
  else return FLOXB_EDTYPE;
#undef FLOXB_SEGQ
  return cudaGetLastError() == cudaSuccess ? FLOXB_OK : FLOXB_ECUDA;
}

// ---------------------------------------------------------------------------------------------------------------------
// grouped scans: cumsum / nancumsum / ffill within groups, in order of appearance
// ---------------------------------------------------------------------------------------------------------------------
// Reference: flox/aggregate_flox.py:296-329 (_np_grouped_scan) and :269-293 (ffill): stable argsort of the codes, ONE
// np.nancumsum over the code-sorted row from which the total preceding each group is subtracted, result un-permuted.
// Here: `perm` is the stable argsort of the LABELS (the caller's device radix sort); the values are never copied into
// sorted order — every kernel gathers values[row, perm[i]] and scatters to out[row, perm[i]] — and the running state is a
// true SEGMENTED scan over the sorted positions (segment heads = code changes), so a group's sums neither carry the rounding
// of earlier groups' totals nor need the reference's NaN-count mask (a NaN propagates through its own segment only).
// Sums accumulate in f64 / i64 and are rounded once to the output dtype.
// Three launches: (A) per 2048-element block the aggregate of its open trailing segment and whether it contains a head,
// (B) per row the carry into every block, (C) the block-local scan with the carry applied up to the block's first head.
namespace floxb {

constexpr int kScanThreads = 256, kScanItems = 8, kScanBlock = kScanThreads * kScanItems;

template <typename VT> struct SumOp {
  using T = typename Acc<VT>::type;
  bool skipna;
  __device__ T identity() const { return (T)0; }
  __device__ T lift(VT v) const { return (skipna && is_nan(v)) ? (T)0 : (T)v; }
  __device__ T combine(T a, T b) const { return a + b; }
  template <typename OT> __device__ OT finish(T s, VT) const { return (OT)s; }
};
template <typename VT> struct FfillState { VT v; int has; };
template <typename VT> struct FfillOp {
  using T = FfillState<VT>;
  __device__ T identity() const { return T{VT(0), 0}; }
  __device__ T lift(VT v) const { return T{v, is_nan(v) ? 0 : 1}; }
  __device__ T combine(T a, T b) const { return b.has ? b : a; }
  template <typename OT> __device__ OT finish(T s, VT own) const { return (OT)(s.has ? s.v : own); }  // leading NaNs stay
};

// Per-thread sequential pass over its kScanItems consecutive sorted positions.  Returns the aggregate of the thread's open
// trailing segment and whether the thread saw a head.  With `emit`, writes the finished elements given the carry-in.
template <typename VT, typename CT, typename OT, typename Op, bool EMIT>
__device__ __forceinline__ void scan_thread(const Op& op, const VT* __restrict__ row, const long long* __restrict__ perm,
                                            const CT* __restrict__ scodes, int64_t n, int64_t base, typename Op::T carry,
                                            bool carry_valid, OT* __restrict__ out_row, typename Op::T& agg, bool& any_head) {
  agg = op.identity();
  any_head = false;
  typename Op::T run = carry;
  bool run_valid = carry_valid;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const int64_t i = base + k;
    if (i >= n) break;
    const bool head = (i == 0) || (scodes[i] != scodes[i - 1]);
    const long long src = perm[i];
    const VT v = row[src];
    const typename Op::T x = op.lift(v);
    if (head) {
      any_head = true;
      agg = x;
      run = x;
      run_valid = true;
    } else {
      agg = op.combine(agg, x);
      run = run_valid ? op.combine(run, x) : x;
      run_valid = true;
    }
    if constexpr (EMIT) out_row[src] = op.template finish<OT>(run, v);
  }
}

// block-level segmented scan of the per-thread (aggregate, head flag) pairs in shared memory; returns for every thread the
// aggregate of everything BEFORE it back to the nearest head (exclusive), and whether a head lies before it in the block
template <typename T, typename Op>
__device__ __forceinline__ void block_exclusive(const Op& op, T agg, bool head, T* s_val, int* s_flag, T& excl, bool& head_before,
                                                T& block_agg, bool& block_head) {
  const int t = threadIdx.x;
  s_val[t] = agg;
  s_flag[t] = head ? 1 : 0;
  __syncthreads();
  for (int d = 1; d < kScanThreads; d <<= 1) {  // Hillis-Steele, inclusive, segmented
    T v = s_val[t];
    int f = s_flag[t];
    if (t >= d && !f) {
      v = op.combine(s_val[t - d], v);
      f = s_flag[t - d];
    }
    __syncthreads();
    s_val[t] = v;
    s_flag[t] = f;
    __syncthreads();
  }
  excl = t > 0 ? s_val[t - 1] : op.identity();
  head_before = t > 0 ? (s_flag[t - 1] != 0) : false;
  block_agg = s_val[kScanThreads - 1];
  block_head = s_flag[kScanThreads - 1] != 0;
  __syncthreads();
}

template <typename VT, typename CT, typename Op>
__global__ void __launch_bounds__(kScanThreads)
k_segscan_totals(Op op, const VT* __restrict__ vals, const long long* __restrict__ perm, const CT* __restrict__ scodes,
                 int64_t n, int64_t nblocks, typename Op::T* __restrict__ blk_agg, int* __restrict__ blk_head) {
  __shared__ typename Op::T s_val[kScanThreads];
  __shared__ int s_flag[kScanThreads];
  const int64_t r = blockIdx.y, b = blockIdx.x;
  typename Op::T agg, excl, bagg;
  bool head, hb, bhead;
  scan_thread<VT, CT, VT, Op, false>(op, vals + r * n, perm, scodes, n, b * kScanBlock + (int64_t)threadIdx.x * kScanItems,
                                     op.identity(), false, nullptr, agg, head);
  block_exclusive(op, agg, head, s_val, s_flag, excl, hb, bagg, bhead);
  if (threadIdx.x == 0) {
    blk_agg[r * nblocks + b] = bagg;
    blk_head[r * nblocks + b] = bhead ? 1 : 0;
  }
}

template <typename Op>
__global__ void k_segscan_carry(Op op, int64_t R, int64_t nblocks, const typename Op::T* __restrict__ blk_agg,
                                const int* __restrict__ blk_head, typename Op::T* __restrict__ carry) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < R; r += (int64_t)gridDim.x * blockDim.x) {
    typename Op::T run = op.identity();
    for (int64_t b = 0; b < nblocks; ++b) {
      carry[r * nblocks + b] = run;   // state entering block b (block 0 starts at a head: never used there)
      const typename Op::T a = blk_agg[r * nblocks + b];
      run = blk_head[r * nblocks + b] ? a : op.combine(run, a);
    }
  }
}

template <typename VT, typename CT, typename OT, typename Op>
__global__ void __launch_bounds__(kScanThreads)
k_segscan_apply(Op op, const VT* __restrict__ vals, const long long* __restrict__ perm, const CT* __restrict__ scodes,
                int64_t n, int64_t nblocks, const typename Op::T* __restrict__ carry, OT* __restrict__ out) {
  __shared__ typename Op::T s_val[kScanThreads];
  __shared__ int s_flag[kScanThreads];
  const int64_t r = blockIdx.y, b = blockIdx.x;
  const int64_t base = b * kScanBlock + (int64_t)threadIdx.x * kScanItems;
  typename Op::T agg, excl, bagg;
  bool head, hb, bhead;
  scan_thread<VT, CT, VT, Op, false>(op, vals + r * n, perm, scodes, n, base, op.identity(), false, nullptr, agg, head);
  block_exclusive(op, agg, head, s_val, s_flag, excl, hb, bagg, bhead);
  // state entering this thread: everything back to the nearest head; if none lies inside the block, the block's carry-in too
  typename Op::T in = excl;
  bool valid = threadIdx.x > 0;
  if (!hb && b > 0) {
    in = valid ? op.combine(carry[r * nblocks + b], excl) : carry[r * nblocks + b];
    valid = true;
  }
  scan_thread<VT, CT, OT, Op, true>(op, vals + r * n, perm, scodes, n, base, in, valid, out + r * n, agg, head);
}

template <typename VT, typename CT, typename OT, typename Op>
int launch_segscan(Op op, const void* vals, const int64_t* perm, const void* scodes, int64_t R, int64_t n, void* scratch,
                   void* out, cudaStream_t st) {
  using T = typename Op::T;
  const int64_t nblocks = (n + kScanBlock - 1) / kScanBlock;
  if (R > 65535) return FLOXB_ERANGE;
  // scratch: [R*nblocks] T aggregates, [R*nblocks] T carries, [R*nblocks] int head flags
  T* agg = reinterpret_cast<T*>(scratch);
  T* carry = agg + R * nblocks;
  int* heads = reinterpret_cast<int*>(carry + R * nblocks);
  dim3 grid((unsigned)nblocks, (unsigned)R);
  const long long* p = reinterpret_cast<const long long*>(perm);
  k_segscan_totals<VT, CT, Op><<<grid, kScanThreads, 0, st>>>(op, reinterpret_cast<const VT*>(vals), p,
                                                              reinterpret_cast<const CT*>(scodes), n, nblocks, agg, heads);
  k_segscan_carry<Op><<<(unsigned)((R + 63) / 64), 64, 0, st>>>(op, R, nblocks, agg, heads, carry);
  k_segscan_apply<VT, CT, OT, Op><<<grid, kScanThreads, 0, st>>>(op, reinterpret_cast<const VT*>(vals), p,
                                                                 reinterpret_cast<const CT*>(scodes), n, nblocks, carry,
                                                                 reinterpret_cast<OT*>(out));
  return cudaGetLastError() == cudaSuccess ? FLOXB_OK : FLOXB_ECUDA;
}

}  // namespace floxb

extern "C" int64_t floxb_segment_scan_scratch_bytes(int64_t n_rows, int64_t n) {
  const int64_t nblocks = (n + kScanBlock - 1) / kScanBlock;
  return n_rows * nblocks * (16 + 16 + 4) + 64;
}

extern "C" int floxb_segment_scan(int kind, const void* values, int value_dtype, const int64_t* perm, const int64_t* sorted_codes,
                                  int64_t n_rows, int64_t n, int skipna, void* scratch, int64_t scratch_bytes, void* out,
                                  void* stream) {
  if (n_rows < 0 || n < 0) return FLOXB_EINVAL;
  if (n_rows * n == 0) return FLOXB_OK;
  if (!values || !perm || !sorted_codes || !scratch || !out) return FLOXB_EINVAL;
  if (scratch_bytes < floxb_segment_scan_scratch_bytes(n_rows, n)) return FLOXB_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define FLOXB_SCAN(VT)                                                                                                      \
  do {                                                                                                                      \
    if (kind == FLOXB_SCAN_CUMSUM) return launch_segscan<VT, int64_t, VT>(SumOp<VT>{skipna != 0}, values, perm, sorted_codes, \
                                                                          n_rows, n, scratch, out, st);                      \
    if (kind == FLOXB_SCAN_FFILL) return launch_segscan<VT, int64_t, VT>(FfillOp<VT>{}, values, perm, sorted_codes, n_rows, n, \
                                                                         scratch, out, st);                                  \
    return FLOXB_EINVAL;                                                                                                    \
  } while (0)
  switch (value_dtype) {
    case FLOXB_F32: FLOXB_SCAN(float);
    case FLOXB_F64: FLOXB_SCAN(double);
    case FLOXB_I32: FLOXB_SCAN(int32_t);
    case FLOXB_I64: FLOXB_SCAN(int64_t);
  }
#undef FLOXB_SCAN
  return FLOXB_EDTYPE;
}
