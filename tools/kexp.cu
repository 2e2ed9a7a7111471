// Kernel-design experiments for the grouped-sum hot loop (NOT part of the product library).
// Built into tools/libkexp.so by tools/kbench.py and timed on the B200 to choose the production
// configuration of k_moments_smem (threads, unroll) and to attribute time between HBM streaming,
// the 64-bit CAS loop and the native 32-bit counter atomics.
#include <cuda_runtime.h>
#include <stdint.h>

enum Mode { FULL = 0, SUM_ONLY = 1, CNT_ONLY = 2, LOAD_ONLY = 3, PACKED32 = 4, F32_CAS = 5, F32_FLUSH4 = 6, F32_FLUSH8 = 7,
            F32_FLUSH16 = 8, HYBRID25 = 9, HYBRID50 = 10, F32_BATCH = 11, F32_CAS_RET = 12, F32_CAS_SWEEP = 13 };

__device__ __forceinline__ float4 ldg4(const float* p);
__device__ __forceinline__ longlong2 ldg2(const long long* p);

// 4 rows at a time with an OPTIMISTIC compare-and-swap each: the four LDS, the four FADDs and the four
// ATOMS.CAS are independent instructions (ILP), only losers fall back to the retry loop.
__device__ __forceinline__ void batch_add4(float* sf32, unsigned* scnt, const int* b, const float* x, const bool* ok) {
  unsigned old[4], nw[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) old[k] = ok[k] ? __float_as_uint(sf32[b[k]]) : 0u;
#pragma unroll
  for (int k = 0; k < 4; ++k) nw[k] = __float_as_uint(__uint_as_float(old[k]) + x[k]);
  unsigned got[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) got[k] = ok[k] ? atomicCAS(reinterpret_cast<unsigned*>(&sf32[b[k]]), old[k], nw[k]) : old[k];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (ok[k]) {
      if (got[k] != old[k]) atomicAdd(&sf32[b[k]], x[k]);
      atomicAdd(&scnt[b[k]], 1u);
    }
  }
}

// Same work as F32_CAS (f32 CAS loop + native count) but with the loads of the NEXT tile in flight while
// the atomics of the current tile run (register double buffering, tile = THREADS*4 rows).
template <int THREADS, bool BATCH>
__global__ void __launch_bounds__(THREADS, 1)
k_exp_pipe(const float* __restrict__ vals, const long long* __restrict__ codes, int64_t n, int G, double* __restrict__ gacc,
           unsigned long long* __restrict__ gcnt) {
  constexpr int TILE = THREADS * 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sf32 = reinterpret_cast<float*>(smem_raw);
  unsigned* scnt = reinterpret_cast<unsigned*>(smem_raw + (size_t)G * 8);
  for (int g = threadIdx.x; g < G; g += THREADS) {
    sf32[g] = 0.f;
    scnt[g] = 0u;
  }
  __syncthreads();
  const int64_t tiles = n / TILE;
  int64_t tile = blockIdx.x;
  float4 v0, v1;
  longlong2 a0, b0, a1, b1;
  if (tile < tiles) {
    const int off = threadIdx.x * 4;
    v0 = ldg4(vals + tile * TILE + off);
    a0 = ldg2(codes + tile * TILE + off);
    b0 = ldg2(codes + tile * TILE + off + 2);
  }
  while (tile < tiles) {
    const int64_t next = tile + gridDim.x;
    if (next < tiles) {
      const int off = threadIdx.x * 4;
      v1 = ldg4(vals + next * TILE + off);
      a1 = ldg2(codes + next * TILE + off);
      b1 = ldg2(codes + next * TILE + off + 2);
    }
    const float x[4] = {v0.x, v0.y, v0.z, v0.w};
    const long long cc[4] = {a0.x, a0.y, b0.x, b0.y};
    if (BATCH) {
      int b[4];
      bool ok[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        ok[k] = (unsigned long long)cc[k] < (unsigned long long)G;
        b[k] = ok[k] ? (int)cc[k] : 0;
      }
      batch_add4(sf32, scnt, b, x, ok);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if ((unsigned long long)cc[k] >= (unsigned long long)G) continue;
        atomicAdd(&sf32[(int)cc[k]], x[k]);
        atomicAdd(&scnt[(int)cc[k]], 1u);
      }
    }
    v0 = v1; a0 = a1; b0 = b1;
    tile = next;
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += THREADS) {
    atomicAdd(&gacc[g], (double)sf32[g]);
    atomicAdd(&gcnt[g], (unsigned long long)scnt[g]);
  }
}

__device__ __forceinline__ float4 ldg4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ longlong2 ldg2(const long long* p) {
  longlong2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.s64 {%0,%1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
  return r;
}

template <int THREADS, int ITERS, int MODE>
__global__ void __launch_bounds__(THREADS, 1)
k_exp(const float* __restrict__ vals, const long long* __restrict__ codes, int64_t n, int G, double* __restrict__ gacc,
      unsigned long long* __restrict__ gcnt) {
  constexpr int TILE = THREADS * 4 * ITERS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sacc = reinterpret_cast<double*>(smem_raw);
  unsigned* scnt = reinterpret_cast<unsigned*>(sacc + G);
  unsigned long long* spk = reinterpret_cast<unsigned long long*>(smem_raw);  // PACKED32: [f32 sum | u32 cnt]
  float* sf32 = reinterpret_cast<float*>(smem_raw);                            // F32_CAS
  for (int g = threadIdx.x; g < G; g += THREADS) {
    sacc[g] = 0.0;
    scnt[g] = 0u;
  }
  __syncthreads();
  float sink = 0.f;
  long long csink = 0;
  const int64_t tiles = n / TILE;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const float* v = vals + tile * TILE;
    const long long* c = codes + tile * TILE;
    float4 vv[ITERS];
    longlong2 c0[ITERS], c1[ITERS];
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int off = (it * THREADS + threadIdx.x) * 4;
      vv[it] = ldg4(v + off);
      c0[it] = ldg2(c + off);
      c1[it] = ldg2(c + off + 2);
    }
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const float x[4] = {vv[it].x, vv[it].y, vv[it].z, vv[it].w};
      const long long cc[4] = {c0[it].x, c0[it].y, c1[it].x, c1[it].y};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (MODE == LOAD_ONLY) {
          sink += x[k];
          csink += cc[k];
          continue;
        }
        if ((unsigned long long)cc[k] >= (unsigned long long)G) continue;
        const int b = (int)cc[k];
        if (MODE == FULL || MODE == SUM_ONLY) atomicAdd(&sacc[b], (double)x[k]);
        if (MODE == FULL || MODE == CNT_ONLY) atomicAdd(&scnt[b], 1u);
        if (MODE == F32_CAS) {
          atomicAdd(&sf32[b], x[k]);
          atomicAdd(&scnt[b], 1u);
        }
        if (MODE == F32_CAS_RET) {
          atomicAdd(&sf32[b], x[k]);
          csink += atomicAdd(&scnt[b], 1u);   // RETURNING counter add (ATOMS.ADD instead of ATOMS.POPC.INC)
        }
        if (MODE == F32_CAS_SWEEP) {
          atomicAdd(&sf32[b], x[k]);
          atomicAdd(&scnt[b], 1u);
        }
        if (MODE == F32_BATCH && k == 0) {
          int bb[4];
          bool ok[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            ok[q] = (unsigned long long)cc[q] < (unsigned long long)G;
            bb[q] = ok[q] ? (int)cc[q] : 0;
          }
          batch_add4(sf32, scnt, bb, x, ok);
        }
        if (MODE == F32_FLUSH4 || MODE == F32_FLUSH8 || MODE == F32_FLUSH16) {
          constexpr unsigned every = MODE == F32_FLUSH4 ? 4u : (MODE == F32_FLUSH8 ? 8u : 16u);
          atomicAdd(&sf32[b], x[k]);
          const unsigned seen = atomicAdd(&scnt[b], 1u);
          if ((seen & (every - 1)) == every - 1) atomicAdd(&gacc[b], (double)atomicExch(&sf32[b], 0.0f));
        }
        if (MODE == HYBRID25 || MODE == HYBRID50) {
          const bool to_l2 = MODE == HYBRID25 ? ((threadIdx.x & 3) == 0) : ((threadIdx.x & 1) == 0);
          if (to_l2) atomicAdd(&gacc[b], (double)x[k]); else atomicAdd(&sacc[b], (double)x[k]);
          atomicAdd(&scnt[b], 1u);
        }
        if (MODE == PACKED32) {
          unsigned long long old = spk[b], assumed;
          do {
            assumed = old;
            const float s = __uint_as_float((unsigned)(assumed >> 32)) + x[k];
            const unsigned long long nw = ((unsigned long long)__float_as_uint(s) << 32) | (unsigned)((unsigned)assumed + 1u);
            old = atomicCAS(&spk[b], assumed, nw);
          } while (old != assumed);
        }
      }
    }
    if (MODE == F32_CAS_SWEEP && (((tile - blockIdx.x) / gridDim.x) & 7) == 7) {
      // barrier-fenced sweep of the f32 partials into an f64 level (kept after the counters), every 8 tiles
      double* lvl2 = reinterpret_cast<double*>(smem_raw + (size_t)G * 12);
      __syncthreads();
      for (int g = threadIdx.x; g < G; g += THREADS) {
        lvl2[g] += (double)sf32[g];
        sf32[g] = 0.f;
      }
      __syncthreads();
    }
  }
  __syncthreads();
  if (MODE == F32_CAS_RET && csink == 42) gcnt[0] = 1;
  if (MODE == LOAD_ONLY) {
    if (sink == 123.456f && csink == 42) gacc[0] = sink;  // keep the loads alive
    return;
  }
  for (int g = threadIdx.x; g < G; g += THREADS) {
    if (MODE == PACKED32) {
      const unsigned long long p = spk[g];
      atomicAdd(&gacc[g], (double)__uint_as_float((unsigned)(p >> 32)));
      atomicAdd(&gcnt[g], (unsigned long long)(unsigned)p);
    } else if (MODE == F32_CAS || MODE == F32_CAS_RET || MODE == F32_CAS_SWEEP || MODE == F32_BATCH || MODE == F32_FLUSH4 || MODE == F32_FLUSH8 || MODE == F32_FLUSH16) {
      atomicAdd(&gacc[g], (double)sf32[g]);
      atomicAdd(&gcnt[g], (unsigned long long)scnt[g]);
    } else {
      atomicAdd(&gacc[g], sacc[g]);
      atomicAdd(&gcnt[g], (unsigned long long)scnt[g]);
    }
  }
}

template <int THREADS, int ITERS, int MODE>
static int launch(const float* v, const long long* c, int64_t n, int G, double* acc, unsigned long long* cnt, int grid,
                  cudaStream_t st) {
  auto k = k_exp<THREADS, ITERS, MODE>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  const size_t smem = (size_t)G * 20 + 16;
  k<<<grid, THREADS, smem, st>>>(v, c, n, G, acc, cnt);
  return (int)cudaGetLastError();
}

#define CASE(ID, T, I, M) \
  case ID: return launch<T, I, M>(v, c, n, G, acc, cnt, grid, st);

extern "C" int kexp_run(int variant, const float* v, const long long* c, int64_t n, int G, double* acc,
                        unsigned long long* cnt, int grid, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  switch (variant) {
    CASE(0, 1024, 2, FULL)
    CASE(1, 1024, 1, FULL)
    CASE(2, 512, 4, FULL)
    CASE(3, 512, 2, FULL)
    CASE(4, 768, 2, FULL)
    CASE(5, 1024, 2, SUM_ONLY)
    CASE(6, 1024, 2, CNT_ONLY)
    CASE(7, 1024, 2, LOAD_ONLY)
    CASE(8, 1024, 2, PACKED32)
    CASE(9, 1024, 2, F32_CAS)
    CASE(10, 512, 4, LOAD_ONLY)
    CASE(11, 256, 4, FULL)
    CASE(12, 1024, 4, LOAD_ONLY)
    CASE(13, 1024, 4, FULL)
    CASE(14, 1024, 2, F32_FLUSH4)
    CASE(15, 1024, 2, F32_FLUSH8)
    CASE(16, 1024, 2, F32_FLUSH16)
    CASE(17, 512, 4, F32_FLUSH8)
    CASE(18, 1024, 2, HYBRID25)
    CASE(19, 1024, 2, HYBRID50)
    CASE(20, 1024, 4, F32_FLUSH8)
    CASE(21, 1024, 2, F32_BATCH)
    CASE(24, 1024, 2, F32_CAS_RET)
    CASE(25, 1024, 2, F32_CAS_SWEEP)
    case 22: {
      auto k = k_exp_pipe<1024, false>;
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      k<<<grid, 1024, (size_t)G * 12 + 16, st>>>(v, c, n, G, acc, cnt);
      return (int)cudaGetLastError();
    }
    case 23: {
      auto k = k_exp_pipe<1024, true>;
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      k<<<grid, 1024, (size_t)G * 12 + 16, st>>>(v, c, n, G, acc, cnt);
      return (int)cudaGetLastError();
    }
  }
  return -1;
}

extern "C" const char* kexp_name(int variant) {
  static const char* names[] = {"full t1024 i2", "full t1024 i1", "full t512 i4",  "full t512 i2",      "full t768 i2",
                                "sum-only t1024 i2", "cnt-only t1024 i2", "load-only t1024 i2", "packed f32|u32 CAS64 t1024 i2",
                                "f32 CAS + cnt t1024 i2", "load-only t512 i4", "full t256 i4", "load-only t1024 i4", "full t1024 i4",
                                "f32 flush4 t1024 i2", "f32 flush8 t1024 i2", "f32 flush16 t1024 i2", "f32 flush8 t512 i4",
                                "hybrid25 f64 t1024 i2", "hybrid50 f64 t1024 i2", "f32 flush8 t1024 i4",
                                "f32 batch4 optimistic CAS t1024 i2", "f32 CAS + cnt pipelined t1024 i1", "f32 batch4 pipelined t1024 i1",
                                "f32 CAS + RETURNING cnt t1024 i2", "f32 CAS + cnt + sweeps t1024 i2"};
  return (variant >= 0 && variant < 26) ? names[variant] : "?";
}
