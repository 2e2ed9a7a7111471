// flox_b200 — label staging for host-resident chunks (include/flox_b200.h: floxb_host_pack_codes,
// floxb_widen_codes).  flox hands every block int64 codes (flox/factorize.py:201-210); for an f32 reduction they are
// 8 of the 12 bytes per row that cross PCIe.  With fewer than 2^15 groups the host narrows them to int16 on the
// way into the pinned staging buffer (one pass, parallelised over chunks by the caller) and the device widens them
// to the int32 the reduction kernels read: 6 instead of 12 bytes per row over the link.
#include "common.cuh"

#if defined(__x86_64__)
#include <emmintrin.h>
#define FLOXB_HAVE_SSE2 1
#endif

namespace floxb {

// The staging buffer is written once and next read by the DMA engine: streaming (non-temporal) 16-byte stores
// keep it out of the cache and skip the read-for-ownership of every output line (the pack is bound by host
// memory bandwidth: tools/e2e_probe.py).
template <typename IT, typename OT>
static void pack_block(const IT* __restrict__ in, int64_t n, uint64_t g, OT* __restrict__ out) {
  for (int64_t i = 0; i < n; ++i) {
    const IT c = in[i];
    out[i] = ((uint64_t)(int64_t)c < g) ? (OT)c : (OT)-1;  // negative codes wrap to huge unsigned values
  }
}

template <typename IT, typename OT>
static void pack_codes(const IT* __restrict__ in, int64_t n, int64_t n_groups, OT* __restrict__ out) {
  const uint64_t g = (uint64_t)n_groups;
  int64_t i = 0;
#ifdef FLOXB_HAVE_SSE2
  // blocks of 4096 codes are packed into an L1-resident buffer (auto-vectorised loop) and streamed out
  constexpr int64_t BLK = 4096, V = 16 / sizeof(OT);
  const int64_t head = (int64_t)(((16 - (reinterpret_cast<uintptr_t>(out) & 15)) & 15) / sizeof(OT));
  if (head <= n && (reinterpret_cast<uintptr_t>(out) % sizeof(OT)) == 0) {
    pack_block(in, head, g, out);
    i = head;
    alignas(16) OT tmp[BLK];
    for (; i + BLK <= n; i += BLK) {
      pack_block(in + i, BLK, g, tmp);
      for (int64_t k = 0; k < BLK; k += V)
        _mm_stream_si128(reinterpret_cast<__m128i*>(out + i + k), _mm_load_si128(reinterpret_cast<const __m128i*>(tmp + k)));
    }
    _mm_sfence();
  }
#endif
  pack_block(in + i, n - i, g, out + i);
}

__global__ void __launch_bounds__(256)
k_widen_codes(const short4* __restrict__ in4, const int16_t* __restrict__ in, int64_t n, int4* __restrict__ out4,
              int32_t* __restrict__ out) {
  const int64_t nvec = in4 ? n / 4 : 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nvec; q += stride) {
    const short4 s = in4[q];
    out4[q] = make_int4(s.x, s.y, s.z, s.w);
  }
  for (int64_t i = nvec * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = in[i];
}

// ---- discovery of unknown integer labels -----------------------------------------------------------------
__global__ void k_minmax_init(long long* mm) {
  mm[0] = INT64_MAX;
  mm[1] = INT64_MIN;
}

template <typename CT>
__global__ void __launch_bounds__(256)
k_label_minmax(const CT* __restrict__ labels, int64_t n, long long* __restrict__ mm) {
  long long lo = INT64_MAX, hi = INT64_MIN;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const long long c = (long long)labels[i];
    lo = c < lo ? c : lo;
    hi = c > hi ? c : hi;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const long long a = __shfl_xor_sync(0xffffffffu, lo, d), b = __shfl_xor_sync(0xffffffffu, hi, d);
    lo = a < lo ? a : lo;
    hi = b > hi ? b : hi;
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&mm[0], lo);
    atomicMax(&mm[1], hi);
  }
}

template <typename CT>
__global__ void __launch_bounds__(256)
k_label_presence(const CT* __restrict__ labels, int64_t n, long long lo, unsigned long long range, uint8_t* __restrict__ table) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long rel = (unsigned long long)((long long)labels[i] - lo);
    if (rel < range && !table[rel]) table[rel] = 1;  // benign race: every writer stores the same byte
  }
}

template <typename CT>
__global__ void __launch_bounds__(256)
k_lut_codes(const CT* __restrict__ labels, int64_t n, long long lo, unsigned long long range, const long long* __restrict__ lut,
            long long* __restrict__ codes) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long rel = (unsigned long long)((long long)labels[i] - lo);
    codes[i] = rel < range ? lut[rel] : -1ll;
  }
}

static inline unsigned label_blocks(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (b > cap) b = cap;
  return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace floxb

using namespace floxb;

extern "C" int floxb_label_minmax(const void* labels, int label_dtype, int64_t n, int64_t* minmax, void* stream) {
  if (n <= 0 || !labels || !minmax) return FLOXB_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  auto* mm = reinterpret_cast<long long*>(minmax);
  k_minmax_init<<<1, 1, 0, st>>>(mm);
  if (label_dtype == FLOXB_I64) k_label_minmax<<<label_blocks(n), 256, 0, st>>>(static_cast<const int64_t*>(labels), n, mm);
  else if (label_dtype == FLOXB_I32) k_label_minmax<<<label_blocks(n), 256, 0, st>>>(static_cast<const int32_t*>(labels), n, mm);
  else return FLOXB_EDTYPE;
  return cudaGetLastError() == cudaSuccess ? FLOXB_OK : FLOXB_ECUDA;
}

extern "C" int floxb_label_presence(const void* labels, int label_dtype, int64_t n, int64_t lo, int64_t range,
                                    uint8_t* table, void* stream) {
  if (n < 0 || range < 0) return FLOXB_EINVAL;
  if (range == 0) return FLOXB_OK;
  if (!table || (n > 0 && !labels)) return FLOXB_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (cudaMemsetAsync(table, 0, (size_t)range, st) != cudaSuccess) return FLOXB_ECUDA;
  if (n == 0) return FLOXB_OK;
  if (label_dtype == FLOXB_I64)
    k_label_presence<<<label_blocks(n), 256, 0, st>>>(static_cast<const int64_t*>(labels), n, (long long)lo, (unsigned long long)range, table);
  else if (label_dtype == FLOXB_I32)
    k_label_presence<<<label_blocks(n), 256, 0, st>>>(static_cast<const int32_t*>(labels), n, (long long)lo, (unsigned long long)range, table);
  else return FLOXB_EDTYPE;
  return cudaGetLastError() == cudaSuccess ? FLOXB_OK : FLOXB_ECUDA;
}

extern "C" int floxb_lut_codes(const void* labels, int label_dtype, int64_t n, int64_t lo, int64_t range, const int64_t* lut,
                               int64_t* codes, void* stream) {
  if (n < 0 || range < 0) return FLOXB_EINVAL;
  if (n == 0) return FLOXB_OK;
  if (!labels || !codes || (range > 0 && !lut)) return FLOXB_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  auto* l = reinterpret_cast<const long long*>(lut);
  auto* c = reinterpret_cast<long long*>(codes);
  if (label_dtype == FLOXB_I64)
    k_lut_codes<<<label_blocks(n), 256, 0, st>>>(static_cast<const int64_t*>(labels), n, (long long)lo, (unsigned long long)range, l, c);
  else if (label_dtype == FLOXB_I32)
    k_lut_codes<<<label_blocks(n), 256, 0, st>>>(static_cast<const int32_t*>(labels), n, (long long)lo, (unsigned long long)range, l, c);
  else return FLOXB_EDTYPE;
  return cudaGetLastError() == cudaSuccess ? FLOXB_OK : FLOXB_ECUDA;
}

extern "C" int floxb_host_pack_codes(const void* codes, int code_dtype, int64_t n, int64_t n_groups, void* out, int out_bits) {
  if (n < 0 || n_groups < 0) return FLOXB_EINVAL;
  if (n == 0) return FLOXB_OK;
  if (!codes || !out) return FLOXB_EINVAL;
  if (out_bits == 16) {
    if (n_groups > 32767) return FLOXB_ERANGE;
    if (code_dtype == FLOXB_I64) pack_codes(static_cast<const int64_t*>(codes), n, n_groups, static_cast<int16_t*>(out));
    else if (code_dtype == FLOXB_I32) pack_codes(static_cast<const int32_t*>(codes), n, n_groups, static_cast<int16_t*>(out));
    else return FLOXB_EDTYPE;
  } else if (out_bits == 32) {
    if (n_groups > 2147483647LL) return FLOXB_ERANGE;
    if (code_dtype == FLOXB_I64) pack_codes(static_cast<const int64_t*>(codes), n, n_groups, static_cast<int32_t*>(out));
    else if (code_dtype == FLOXB_I32) pack_codes(static_cast<const int32_t*>(codes), n, n_groups, static_cast<int32_t*>(out));
    else return FLOXB_EDTYPE;
  } else {
    return FLOXB_EINVAL;
  }
  return FLOXB_OK;
}

extern "C" int floxb_widen_codes(const int16_t* codes16, int64_t n, int32_t* codes32, void* stream) {
  if (n < 0) return FLOXB_EINVAL;
  if (n == 0) return FLOXB_OK;
  if (!codes16 || !codes32) return FLOXB_EINVAL;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool vec = (reinterpret_cast<uintptr_t>(codes16) & 7) == 0 && (reinterpret_cast<uintptr_t>(codes32) & 15) == 0;
  int64_t blocks = (n / 4 + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  k_widen_codes<<<(unsigned)blocks, 256, 0, st>>>(vec ? reinterpret_cast<const short4*>(codes16) : nullptr, codes16, n,
                                                 reinterpret_cast<int4*>(codes32), codes32);
  return cudaGetLastError() == cudaSuccess ? FLOXB_OK : FLOXB_ECUDA;
}
