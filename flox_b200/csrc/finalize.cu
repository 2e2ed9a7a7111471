// Typed results from the wide accumulators, combine of stacked blocks, Chan variance merge,
// and the small label-encoding kernels.  Everything here is O(#bins) or O(n) elementwise.
//
//   floxb_finalize    flox/core.py:410-475 (_finalize_results) arithmetic: absent-group fill
//                     (aggregate_flox.py:174-180), NaN propagation, all-NaN -> fill for
//                     nanmax/nanmin (aggregate_flox.py:203-206), mean (aggregations.py:323-325),
//                     var/std (aggregations.py:460-472), min_count mask (core.py:438-459).
//   floxb_combine     flox/dask.py:131-143 with agg.simple_combine in {np.sum, np.max, np.min, np.prod}.
//   floxb_var_combine flox/aggregations.py:392-451 (_var_combine).
//   floxb_digitize    flox/factorize.py:55-82;  floxb_ravel_codes  flox/factorize.py:102-108.
#include <math_constants.h>

#include "common.cuh"

namespace floxb {

template <typename OT> __device__ __forceinline__ OT make_nan() { return (OT)0; }
template <> __device__ __forceinline__ float make_nan<float>() { return CUDART_NAN_F; }
template <> __device__ __forceinline__ double make_nan<double>() { return CUDART_NAN; }

template <typename VT, typename OT>
__global__ void __launch_bounds__(256)
k_finalize(int kind, const void* __restrict__ acc_, const void* __restrict__ acc2_, const long long* __restrict__ cnt,
           const long long* __restrict__ nnan, int64_t nbins, int skipna, OT fill, long long min_count, OT mc_fill,
           long long ddof, OT* __restrict__ out, int64_t gm_ab, int64_t gm_g) {
  using AccT = typename Acc<VT>::type;
  constexpr bool OF = is_float<OT>::value;
  // gm_ab != 0: the accumulators are group-major [G][AB] (column-owner kernels); `out` is always [AB][G].
  // Threads walk the ACCUMULATOR order so that the wide reads coalesce; the narrow writes merge in L2.
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nbins; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = j;                                             // accumulator index
    const int64_t o = gm_ab ? (j % gm_ab) * gm_g + (j / gm_ab) : j;   // output index
    const long long c = cnt ? cnt[i] : 0, m = nnan ? nnan[i] : 0;
    const long long rows = c + m;
    const bool poisoned = (!skipna) && m > 0;  // a NaN reached a NaN-propagating reduction
    OT r;
    switch (kind) {
      case FLOXB_FIN_SUM: {
        const AccT a = reinterpret_cast<const AccT*>(acc_)[i];
        r = rows == 0 ? fill : (poisoned ? make_nan<OT>() : (OT)a);
        break;
      }
      case FLOXB_FIN_NANLEN: r = rows == 0 ? fill : (OT)c; break;
      case FLOXB_FIN_MEAN: {
        const AccT a = reinterpret_cast<const AccT*>(acc_)[i];
        double s;
        if (rows == 0) s = (double)fill;           // the reference divides the fill too: fill / 0
        else if (poisoned) s = CUDART_NAN;
        else s = (sizeof(OT) == 4 && OF) ? (double)(float)a : (double)a;
        r = (OT)(s / (double)c);
        break;
      }
      case FLOXB_FIN_MAX:
      case FLOXB_FIN_MIN: {
        if (rows == 0) { r = fill; break; }
        if (poisoned) { r = make_nan<OT>(); break; }
        if (c == 0) { r = fill; break; }            // all NaN under nanmax/nanmin
        const VT v = Key<VT>::dec((typename Key<VT>::type) reinterpret_cast<const long long*>(acc_)[i]);
        r = (OT)v;
        if constexpr (is_float<VT>::value) {
          // reference quirk: a nan-skipping result equal to its +-inf identity counts as "all NaN"
          if (skipna && ((kind == FLOXB_FIN_MAX && v == -CUDART_INF) || (kind == FLOXB_FIN_MIN && v == CUDART_INF))) r = fill;
        }
        break;
      }
      case FLOXB_FIN_VAR:
      case FLOXB_FIN_STD: {
        const double m2 = reinterpret_cast<const double*>(acc_)[i];
        const double den = (double)(c - ddof);
        double v = poisoned ? CUDART_NAN : m2 / den;
        if (den < 0) v = CUDART_NAN;
        if (kind == FLOXB_FIN_STD) v = sqrt(v);
        r = (OT)v;
        break;
      }
      case FLOXB_FIN_PICK: {
        const AccT a = reinterpret_cast<const AccT*>(acc_)[i];
        r = (rows == 0 || (skipna && c == 0)) ? fill : (OT)a;
        break;
      }
      case FLOXB_FIN_ARG: {
        const long long ix = reinterpret_cast<const long long*>(acc_)[i];
        r = (ix == INT64_MAX) ? fill : (OT)ix;
        break;
      }
      default: r = fill;
    }
    if (min_count > 0 && c < min_count) r = mc_fill;
    out[o] = r;
  }
}

template <typename T, int KIND>
__global__ void __launch_bounds__(256)
k_combine(const T* __restrict__ stacked, int64_t nblk, int64_t nbins, T* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbins; i += (int64_t)gridDim.x * blockDim.x) {
    T r = stacked[i];
    for (int64_t k = 1; k < nblk; ++k) {
      const T v = stacked[k * nbins + i];
      if constexpr (KIND == FLOXB_COMB_SUM) r += v;
      else if constexpr (KIND == FLOXB_COMB_PROD) r *= v;
      else if constexpr (KIND == FLOXB_COMB_MAX) r = (v > r || v != v) ? v : r;   // NaN propagates like np.max
      else if constexpr (KIND == FLOXB_COMB_MIN) r = (v < r || v != v) ? v : r;
      else if constexpr (KIND == FLOXB_COMB_NANMAX) r = (v > r || r != r) ? v : r;  // NaN skipped like np.nanmax
      else r = (v < r || r != r) ? v : r;
    }
    out[i] = r;
  }
}

// Sequential pairwise merge along the stacked axis == the reference's cumsum formulation:
//   adj_k = (N_<k * S_k - n_k * S_<k)^2 / (N_<k * n_k * (N_<k + n_k)), 0 when either side is empty.
__global__ void __launch_bounds__(256)
k_var_combine(const double* __restrict__ m2, const double* __restrict__ s, const double* __restrict__ n, int64_t nblk,
              int64_t nbins, double* __restrict__ om2, double* __restrict__ os, double* __restrict__ on) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbins; i += (int64_t)gridDim.x * blockDim.x) {
    double M = m2[i], S = s[i], N = n[i], adj = 0.0;
    for (int64_t k = 1; k < nblk; ++k) {
      const double mk = m2[k * nbins + i], sk = s[k * nbins + i], nk = n[k * nbins + i];
      const bool degenerate = (N == 0.0) || (nk == 0.0);
      const double num = N * sk - nk * S;
      adj += (num * num) / (N * nk * (N + nk) + (degenerate ? 1.0 : 0.0));
      M += mk;
      S += sk;
      N += nk;
    }
    om2[i] = M + adj;
    os[i] = S;
    on[i] = N;
  }
}

template <typename LT>
__global__ void __launch_bounds__(256)
k_digitize(const LT* __restrict__ labels, int64_t n, const double* __restrict__ edges, int64_t ne, int right,
           long long* __restrict__ codes) {
  const double lo = edges[0], hi = edges[ne - 1];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double x = (double)labels[i];
    long long code = -1;
    const bool inside = right ? (x > lo && x <= hi) : (x >= lo && x < hi);
    if (inside) {  // NaN fails both comparisons -> -1
      // np.digitize(x, edges, right) - 1 for increasing edges
      int64_t a = 0, b = ne;  // first edge index with (right ? edge >= x : edge > x)
      while (a < b) {
        const int64_t mid = (a + b) >> 1;
        const double e = edges[mid];
        const bool go_right = right ? (e < x) : (e <= x);
        if (go_right) a = mid + 1; else b = mid;
      }
      code = a - 1;
    }
    codes[i] = code;
  }
}

__global__ void __launch_bounds__(256)
k_ravel_codes(const long long* __restrict__ c0, const long long* __restrict__ c1, int64_t n, long long s0, long long s1,
              long long* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const long long a = c0[i], b = c1[i];
    out[i] = (a < 0 || b < 0 || a >= s0 || b >= s1) ? -1 : a * s1 + b;
  }
}

static inline int ew_blocks(int64_t n) {
  const int64_t cap = (int64_t)sm_count() * 8;
  const int64_t b = (n + 255) / 256;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

template <typename VT, typename OT>
int launch_finalize(int kind, const void* acc, const void* acc2, const int64_t* cnt, const int64_t* nnan, int64_t nbins,
                    int skipna, double fill_f, int64_t fill_i, int64_t min_count, double mc_f, int64_t mc_i, int64_t ddof,
                    void* out, int64_t gm_ab, cudaStream_t st) {
  const OT fill = is_float<OT>::value ? (OT)fill_f : (OT)fill_i;
  const OT mcf = is_float<OT>::value ? (OT)mc_f : (OT)mc_i;
  k_finalize<VT, OT><<<ew_blocks(nbins), 256, 0, st>>>(kind, acc, acc2, reinterpret_cast<const long long*>(cnt),
                                                      reinterpret_cast<const long long*>(nnan), nbins, skipna, fill,
                                                      (long long)min_count, mcf, (long long)ddof, reinterpret_cast<OT*>(out),
                                                      gm_ab, gm_ab ? nbins / gm_ab : 0);
  return cudaGetLastError() == cudaSuccess ? FLOXB_OK : FLOXB_ECUDA;
}

template <typename VT>
int finalize_out(int out_dtype, int kind, const void* acc, const void* acc2, const int64_t* cnt, const int64_t* nnan,
                 int64_t nbins, int skipna, double fill_f, int64_t fill_i, int64_t min_count, double mc_f, int64_t mc_i,
                 int64_t ddof, void* out, int64_t gm_ab, cudaStream_t st) {
  switch (out_dtype) {
    case FLOXB_F32: return launch_finalize<VT, float>(kind, acc, acc2, cnt, nnan, nbins, skipna, fill_f, fill_i, min_count, mc_f, mc_i, ddof, out, gm_ab, st);
    case FLOXB_F64: return launch_finalize<VT, double>(kind, acc, acc2, cnt, nnan, nbins, skipna, fill_f, fill_i, min_count, mc_f, mc_i, ddof, out, gm_ab, st);
    case FLOXB_I32: return launch_finalize<VT, int32_t>(kind, acc, acc2, cnt, nnan, nbins, skipna, fill_f, fill_i, min_count, mc_f, mc_i, ddof, out, gm_ab, st);
    case FLOXB_I64: return launch_finalize<VT, int64_t>(kind, acc, acc2, cnt, nnan, nbins, skipna, fill_f, fill_i, min_count, mc_f, mc_i, ddof, out, gm_ab, st);
  }
  return FLOXB_EDTYPE;
}

template <typename T>
int combine_t(int kind, const void* stacked, int64_t nblk, int64_t nbins, void* out, cudaStream_t st) {
  const T* s = reinterpret_cast<const T*>(stacked);
  T* o = reinterpret_cast<T*>(out);
  const int b = ew_blocks(nbins);
  switch (kind) {
    case FLOXB_COMB_SUM: k_combine<T, FLOXB_COMB_SUM><<<b, 256, 0, st>>>(s, nblk, nbins, o); break;
    case FLOXB_COMB_MAX: k_combine<T, FLOXB_COMB_MAX><<<b, 256, 0, st>>>(s, nblk, nbins, o); break;
    case FLOXB_COMB_MIN: k_combine<T, FLOXB_COMB_MIN><<<b, 256, 0, st>>>(s, nblk, nbins, o); break;
    case FLOXB_COMB_PROD: k_combine<T, FLOXB_COMB_PROD><<<b, 256, 0, st>>>(s, nblk, nbins, o); break;
    case FLOXB_COMB_NANMAX: k_combine<T, FLOXB_COMB_NANMAX><<<b, 256, 0, st>>>(s, nblk, nbins, o); break;
    case FLOXB_COMB_NANMIN: k_combine<T, FLOXB_COMB_NANMIN><<<b, 256, 0, st>>>(s, nblk, nbins, o); break;
    default: return FLOXB_EINVAL;
  }
  return cudaGetLastError() == cudaSuccess ? FLOXB_OK : FLOXB_ECUDA;
}

}  // namespace floxb

using namespace floxb;

extern "C" int floxb_finalize(int kind, int value_dtype, int out_dtype, const void* acc, const void* acc2,
                              const int64_t* cnt, const int64_t* nnan, int64_t n_bins, int skipna, double fill_f,
                              int64_t fill_i, int64_t min_count, double mc_fill_f, int64_t mc_fill_i, int64_t ddof,
                              void* out, int64_t group_major_ab, void* stream) {
  if (group_major_ab < 0 || (group_major_ab > 0 && n_bins % group_major_ab != 0)) return FLOXB_EINVAL;
  if (n_bins < 0) return FLOXB_EINVAL;
  if (n_bins == 0) return FLOXB_OK;
  if (!out || !cnt) return FLOXB_EINVAL;
  if (kind != FLOXB_FIN_NANLEN && !acc) return FLOXB_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (value_dtype) {
    case FLOXB_F32: return finalize_out<float>(out_dtype, kind, acc, acc2, cnt, nnan, n_bins, skipna, fill_f, fill_i, min_count, mc_fill_f, mc_fill_i, ddof, out, group_major_ab, st);
    case FLOXB_F64: return finalize_out<double>(out_dtype, kind, acc, acc2, cnt, nnan, n_bins, skipna, fill_f, fill_i, min_count, mc_fill_f, mc_fill_i, ddof, out, group_major_ab, st);
    case FLOXB_I32: return finalize_out<int32_t>(out_dtype, kind, acc, acc2, cnt, nnan, n_bins, skipna, fill_f, fill_i, min_count, mc_fill_f, mc_fill_i, ddof, out, group_major_ab, st);
    case FLOXB_I64: return finalize_out<int64_t>(out_dtype, kind, acc, acc2, cnt, nnan, n_bins, skipna, fill_f, fill_i, min_count, mc_fill_f, mc_fill_i, ddof, out, group_major_ab, st);
  }
  return FLOXB_EDTYPE;
}

extern "C" int floxb_combine(int kind, int dtype, const void* stacked, int64_t nblk, int64_t n_bins, void* out,
                             void* stream) {
  if (nblk < 1 || n_bins < 0) return FLOXB_EINVAL;
  if (n_bins == 0) return FLOXB_OK;
  if (!stacked || !out) return FLOXB_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == FLOXB_F64) return combine_t<double>(kind, stacked, nblk, n_bins, out, st);
  if (dtype == FLOXB_I64) return combine_t<long long>(kind, stacked, nblk, n_bins, out, st);
  return FLOXB_EDTYPE;
}

extern "C" int floxb_var_combine(const double* m2, const double* s, const double* n, int64_t nblk, int64_t n_bins,
                                 double* out_m2, double* out_s, double* out_n, void* stream) {
  if (nblk < 1 || n_bins < 0) return FLOXB_EINVAL;
  if (n_bins == 0) return FLOXB_OK;
  if (!m2 || !s || !n || !out_m2 || !out_s || !out_n) return FLOXB_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  k_var_combine<<<ew_blocks(n_bins), 256, 0, st>>>(m2, s, n, nblk, n_bins, out_m2, out_s, out_n);
  return cudaGetLastError() == cudaSuccess ? FLOXB_OK : FLOXB_ECUDA;
}

extern "C" int floxb_digitize(const void* labels, int label_dtype, int64_t n, const double* edges, int64_t n_edges,
                              int closed_right, int64_t* codes, void* stream) {
  if (n < 0 || n_edges < 0) return FLOXB_EINVAL;
  if (n == 0) return FLOXB_OK;
  if (!labels || !codes || (n_edges > 0 && !edges)) return FLOXB_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  long long* out = reinterpret_cast<long long*>(codes);
  if (n_edges < 2) {  // factorize.py:80-81: fewer than two edges -> everything is outside
    cudaMemsetAsync(out, 0xff, (size_t)n * 8, st);
    return cudaGetLastError() == cudaSuccess ? FLOXB_OK : FLOXB_ECUDA;
  }
  const int b = ew_blocks(n);
  switch (label_dtype) {
    case FLOXB_F32: k_digitize<float><<<b, 256, 0, st>>>(reinterpret_cast<const float*>(labels), n, edges, n_edges, closed_right, out); break;
    case FLOXB_F64: k_digitize<double><<<b, 256, 0, st>>>(reinterpret_cast<const double*>(labels), n, edges, n_edges, closed_right, out); break;
    case FLOXB_I64: k_digitize<long long><<<b, 256, 0, st>>>(reinterpret_cast<const long long*>(labels), n, edges, n_edges, closed_right, out); break;
    case FLOXB_I32: k_digitize<int><<<b, 256, 0, st>>>(reinterpret_cast<const int*>(labels), n, edges, n_edges, closed_right, out); break;
    default: return FLOXB_EDTYPE;
  }
  return cudaGetLastError() == cudaSuccess ? FLOXB_OK : FLOXB_ECUDA;
}

extern "C" int floxb_ravel_codes(const int64_t* c0, const int64_t* c1, int64_t n, int64_t size0, int64_t size1,
                                 int64_t* out, void* stream) {
  if (n < 0 || size0 < 0 || size1 < 0) return FLOXB_EINVAL;
  if (n == 0) return FLOXB_OK;
  if (!c0 || !c1 || !out) return FLOXB_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  k_ravel_codes<<<ew_blocks(n), 256, 0, st>>>(reinterpret_cast<const long long*>(c0), reinterpret_cast<const long long*>(c1),
                                              n, size0, size1, reinterpret_cast<long long*>(out));
  return cudaGetLastError() == cudaSuccess ? FLOXB_OK : FLOXB_ECUDA;
}

static thread_local char g_cuda_err[256] = "";

extern "C" int floxb_version(void) { return 100; }

extern "C" const char* floxb_strerror(int code) {
  switch (code) {
    case FLOXB_OK: return "ok";
    case FLOXB_EINVAL: return "invalid argument";
    case FLOXB_EDTYPE: return "unsupported dtype combination";
    case FLOXB_ECUDA: return "CUDA runtime error";
    case FLOXB_ERANGE: return "size out of range";
  }
  return "unknown error";
}

extern "C" const char* floxb_last_cuda_error(void) {
  cudaError_t e = cudaGetLastError();
  const char* s = cudaGetErrorString(e);
  int i = 0;
  for (; s && s[i] && i < 255; ++i) g_cuda_err[i] = s[i];
  g_cuda_err[i] = 0;
  return g_cuda_err;
}
