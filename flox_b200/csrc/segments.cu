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
