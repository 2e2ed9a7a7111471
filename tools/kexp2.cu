// Experiments for the L2-atomic path with kept trailing dims (BASELINE config 4: f64 (n, 8), G = 1e5):
// does the accumulator layout ([B][G] vs [G][B]) or a once-per-row count change the L2 atomic rate?
// NOT part of the product library; timed by tools/kbench2.py.
#include <cuda_runtime.h>
#include <stdint.h>

// variant 0: bins [b][g], per element: RED.f64 + RED.u64 count            (current product behaviour)
// variant 1: bins [g][b], per element: RED.f64 + RED.u64 count
// variant 2: bins [g][b], per element: RED.f64; per ROW one RED.u64 row count; NaN elements RED a NaN count
// variant 3: bins [b][g], like 2
// variant 4: loads only
template <int VARIANT>
__global__ void __launch_bounds__(256)
k_exp2(const double* __restrict__ vals, const long long* __restrict__ codes, int64_t n, int B, int64_t G,
       double* __restrict__ acc, unsigned long long* __restrict__ cnt, unsigned long long* __restrict__ nan,
       unsigned long long* __restrict__ rows) {
  const int64_t total = n * B;
  double sink = 0;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = j / B;
    const int b = (int)(j - i * B);
    const long long c = codes[i];
    const double v = vals[j];
    if (VARIANT == 4) { sink += v + (double)c; continue; }
    if ((unsigned long long)c >= (unsigned long long)G) continue;
    const int64_t bin = (VARIANT == 0 || VARIANT == 3) ? (int64_t)b * G + c : c * B + b;
    if (VARIANT <= 1) {
      if (v != v) atomicAdd(&nan[bin], 1ull); else { atomicAdd(&acc[bin], v); atomicAdd(&cnt[bin], 1ull); }
    } else {
      if (v != v) atomicAdd(&nan[bin], 1ull); else atomicAdd(&acc[bin], v);
      if (b == 0) atomicAdd(&rows[c], 1ull);
    }
  }
  if (VARIANT == 4 && sink == 1.2345) acc[0] = sink;
}

extern "C" int kexp2_run(int variant, const double* v, const long long* c, int64_t n, int B, int64_t G, double* acc,
                         unsigned long long* cnt, unsigned long long* nan, unsigned long long* rows, int grid, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  switch (variant) {
    case 0: k_exp2<0><<<grid, 256, 0, st>>>(v, c, n, B, G, acc, cnt, nan, rows); break;
    case 1: k_exp2<1><<<grid, 256, 0, st>>>(v, c, n, B, G, acc, cnt, nan, rows); break;
    case 2: k_exp2<2><<<grid, 256, 0, st>>>(v, c, n, B, G, acc, cnt, nan, rows); break;
    case 3: k_exp2<3><<<grid, 256, 0, st>>>(v, c, n, B, G, acc, cnt, nan, rows); break;
    case 4: k_exp2<4><<<grid, 256, 0, st>>>(v, c, n, B, G, acc, cnt, nan, rows); break;
    default: return -1;
  }
  return (int)cudaGetLastError();
}
