// Bisect harness: which feature of the production two-level kernel (flox_b200/csrc/moments.cu) costs what,
// relative to the bare software-pipelined loop of tools/kexp.cu variant 22.  Uses the production walk_tiles.
// Built and timed by tools/kbench3.py on the GPU box.  Variants 1-5, 7, 9 measured the earlier three-buffer
// walk (removed; numbers kept in profiles/r01_kbench3_bisect_oldwalk.json).
#include "../flox_b200/csrc/common.cuh"

using namespace floxb;

namespace {

enum : unsigned { F_WALK = 1, F_NAN = 2, F_RET = 4, F_SWEEP = 8, F_WS = 16, F_NOPF = 32, F_PF8 = 64, F_WALK2 = 128, F_SWEEP2 = 256, F_SW64 = 512, F_PACKED = 1024 };

template <unsigned F>
__global__ void __launch_bounds__(1024, 1)
k_bisect(const __grid_constant__ ChunkTable tab, const float* __restrict__ vals,
         const long long* __restrict__ codes, int64_t n, int G, double* __restrict__ gacc,
         unsigned long long* __restrict__ gcnt, double* __restrict__ workspace) {
  constexpr int THREADS = 1024;
  constexpr int TILE = THREADS * 4;
  constexpr bool WS = (F & F_WS) != 0;
  constexpr bool SWEEP2 = (F & F_SWEEP2) != 0;  // barrier-free sweeps (atomicExch), level 2 in the workspace
  constexpr bool SWEEP = (F & F_SWEEP) != 0 || SWEEP2;
  constexpr unsigned kDrainEvery = (F & F_SW64) ? 64 : 32;
  constexpr int kSweepTiles = (F & F_SW64) ? 64 : 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int S = G;
  const size_t ws_stride = (size_t)S * 8 + (((size_t)S * 4 + 7) & ~(size_t)7);
  double* sacc = WS ? reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(workspace) + (size_t)blockIdx.x * ws_stride)
                    : reinterpret_cast<double*>(smem_raw);
  unsigned* stot = WS ? reinterpret_cast<unsigned*>(sacc + S) : reinterpret_cast<unsigned*>(smem_raw + (size_t)S * 8);
  const bool l2_in_smem = SWEEP && !WS;
  float* spart = reinterpret_cast<float*>(smem_raw + (l2_in_smem ? (size_t)S * 12 : 0));
  unsigned* scnt = reinterpret_cast<unsigned*>(spart + S);
  unsigned* snanbits = scnt + S;
  for (int s = threadIdx.x; s < S; s += THREADS) {
    if (SWEEP) { sacc[s] = 0.0; stot[s] = 0u; }
    spart[s] = 0.f;
    scnt[s] = 0u;
  }
  for (int w = threadIdx.x; w < (G + 31) / 32; w += THREADS) snanbits[w] = 0u;
  __syncthreads();

  auto add_row = [&](float v, int64_t c) {
    if ((unsigned long long)c >= (unsigned long long)G) return;
    if constexpr ((F & F_NAN) != 0) {
      if (v != v) {
        const unsigned bit = 1u << ((int)c & 31);
        if (!(snanbits[(int)c >> 5] & bit)) atomicOr(&snanbits[(int)c >> 5], bit);
        return;
      }
    }
    const int sl = (int)c;
    if constexpr ((F & F_PACKED) != 0) {
      // upper bound of a "count and fixed-point sum in ONE native 32-bit ATOMS.ADD" design: [count 6 bits | sum 26 bits].
      // Throughput probe only (no drains, the fields overflow): what would removing the LDS + CAS loop buy?
      atomicAdd(&scnt[sl], (1u << 26) + (unsigned)__float2int_rn(v * 524288.0f));
      return;
    }
    atomicAdd(&spart[sl], v);
    if constexpr ((F & F_RET) != 0) {
      const unsigned seen = atomicAdd(&scnt[sl], 1u);
      if constexpr (SWEEP) {
        if (seen >= kDrainEvery - 1) {
          const float part = atomicExch(&spart[sl], 0.0f);
          atomicAdd(&sacc[sl], (double)part);
          atomicAdd(&stot[sl], atomicExch(&scnt[sl], 0u));
        }
      } else if (seen == 0xffffffffu) {  // never true: keeps the atomic a RETURNING one
        gcnt[0] = 1ull;
      }
    } else {
      atomicAdd(&scnt[sl], 1u);
    }
  };
  int since_sweep = 0;
  auto after_tile = [&] {
    if constexpr (SWEEP) {
      if (SWEEP2) {
        if (++since_sweep == kSweepTiles) {
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
      } else if (++since_sweep == kSweepTiles) {
        since_sweep = 0;
        __syncthreads();
        for (int sl = threadIdx.x; sl < S; sl += THREADS) {
          const unsigned c = scnt[sl];
          if (c) {
            sacc[sl] += (double)spart[sl];
            spart[sl] = 0.0f;
            stot[sl] += c;
            scnt[sl] = 0u;
          }
        }
        __syncthreads();
      }
    }
  };

  if constexpr ((F & F_WALK2) != 0) {
    walk_tiles<float, int64_t, THREADS>(tab, add_row, after_tile);
  } else {
    // bare depth-1 register pipeline over one flat array (kexp.cu variant 22)
    const int64_t tiles = n / TILE;
    int64_t tile = blockIdx.x;
    Vec4<float> v0, v1;
    Vec4<int64_t> c0, c1;
    if (tile < tiles) {
      v0 = load4<float>(vals + tile * TILE + threadIdx.x * 4);
      c0 = load4<int64_t>(reinterpret_cast<const int64_t*>(codes) + tile * TILE + threadIdx.x * 4);
    }
    while (tile < tiles) {
      const int64_t next = tile + gridDim.x;
      if (next < tiles) {
        v1 = load4<float>(vals + next * TILE + threadIdx.x * 4);
        c1 = load4<int64_t>(reinterpret_cast<const int64_t*>(codes) + next * TILE + threadIdx.x * 4);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) add_row(v0.v[k], c0.v[k]);
      after_tile();
      v0 = v1;
      c0 = c1;
      tile = next;
    }
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += THREADS) {
    double tot = (double)spart[g];
    unsigned long long c = scnt[g];
    if (SWEEP) { tot += sacc[g]; c += stot[g]; }
    if (c) {
      atomicAdd(&gacc[g], tot);
      atomicAdd(&gcnt[g], c);
    }
  }
}

template <unsigned F>
int launch(const ChunkTable& tab, const float* v, const long long* c, int64_t n, int G, double* acc,
           unsigned long long* cnt, double* ws, int grid, cudaStream_t st) {
  const bool l2_in_smem = (F & F_SWEEP) && !(F & F_WS);
  const size_t smem = (size_t)G * (l2_in_smem ? 20 : 8) + ((G + 31) / 32) * 4 + 16;
  auto k = k_bisect<F>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  k<<<grid, 1024, smem, st>>>(tab, v, c, n, G, acc, cnt, ws);
  return (int)cudaGetLastError();
}

}  // namespace

// `n` rows in one flat allocation, presented to walk_tiles as `nchunks` equal slices
extern "C" int kexp3_run(int variant, const float* v, const long long* c, int64_t n, int nchunks, int G, double* acc,
                         unsigned long long* cnt, double* ws, int grid, void* stream) {
  ChunkTable tab{};
  const int64_t per = n / nchunks;
  tab.nchunks = nchunks;
  for (int i = 0; i < nchunks; ++i) {
    tab.values[i] = v + (int64_t)i * per;
    tab.codes[i] = c + (int64_t)i * per;
    tab.n[i] = per;
  }
  cudaStream_t st = (cudaStream_t)stream;
#define V(ID, FLAGS) case ID: return launch<(FLAGS)>(tab, v, c, n, G, acc, cnt, ws, grid, st);
  switch (variant) {
    V(0, 0u)
    V(6, F_NAN | F_RET | F_SWEEP | F_WS)
    V(8, F_NAN | F_RET)
    V(10, F_SWEEP | F_WS)
    V(11, F_NAN)
    V(12, F_RET)
    V(13, F_WALK2)
    V(14, F_WALK2 | F_NAN | F_RET)
    V(15, F_WALK2 | F_NAN | F_RET | F_SWEEP | F_WS)
    V(16, F_WALK2 | F_NAN | F_RET | F_SWEEP2 | F_WS)
    V(17, F_NAN | F_RET | F_SWEEP2 | F_WS)
    V(18, F_WALK2 | F_NAN | F_RET | F_SWEEP2 | F_WS | F_SW64)
    V(19, F_WALK2 | F_NAN | F_RET | F_SWEEP | F_WS | F_SW64)
    V(20, F_PACKED)
    V(21, F_WALK2 | F_NAN | F_PACKED)
  }
  return -1;
}

extern "C" const char* kexp3_name(int variant) {
  static const char* names[] = {"bare pipe", "walk", "walk+nan", "walk+nan+ret", "walk+nan+ret+sweep(smem)", "walk+nan+ret+sweep(ws) = production",
                                "bare+nan+ret+sweep(ws)", "walk, no L2 prefetch", "bare+nan+ret", "walk, L2 prefetch 8 ahead",
                                "bare+sweep(ws)", "bare+nan", "bare+ret", "walk2", "walk2+nan+ret", "walk2+nan+ret+sweep(ws)",
                                "walk2+nan+ret+sweep2(ws, no barrier)", "bare+nan+ret+sweep2", "walk2+nan+ret+sweep2 every 64",
                                "walk2+nan+ret+sweep(ws) every 64", "bare, ONE native atomic per row (packed probe)",
                                "walk+nan, ONE native atomic per row (packed probe)"};
  return (variant >= 0 && variant < 22) ? names[variant] : "?";
}
