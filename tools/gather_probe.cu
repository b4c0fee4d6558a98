// gather_probe.cu -- how fast can one B200 SM gather 64-byte rows?  (measurement tool, not product code)
//
// The sampling kernel's inner loop is "fetch a 64-byte (pixel, head) row at a data-dependent index".  This probe
// measures the ceiling of that primitive on its own, for every way the hardware offers to issue it, so the
// kernel's roofline has a measured denominator instead of a guess:
//
//   mode 0  LDG.128, 4 lanes per row    (8 rows  / warp instruction)   <- the round-1 kernel's shape
//   mode 1  LDG.256, 2 lanes per row    (16 rows / warp instruction)
//   mode 2  LDG.128, 8 lanes per 128-byte line (4 full lines / instruction): the L1's best case
//   mode 3  LDS.128 from a shared-memory table, 4 lanes per row, neighbouring lane groups on opposite bank halves
//   mode 4  LDS.128, random bank halves
//   mode 5  TMA tile::gather4 (one elected lane per warp, 4 rows per instruction) into shared memory
//   mode 6  cp.async.bulk, one 64-byte row per instruction (one elected lane per warp)
//   mode 7  LDG.64, 8 lanes per row     (4 rows / warp instruction)
//
// Rows are drawn with a per-lane-group LCG from a table of `rows` rows (default 147,312 = one 1152x768 pyramid,
// 9.4 MB: L2-resident, L1-missing) or from a window of `window` rows around a moving base (L1-hitting).
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o build_variants/gather_probe tools/gather_probe.cu -lcuda
// Run:    build_variants/gather_probe <mode> [ctas_per_sm] [unroll] [window] [rows]
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      printf("{\"error\": \"%s at line %d\"}\n", cudaGetErrorString(e_), __LINE__);     \
      return 1;                                                                            \
    }                                                                                      \
  } while (0)

constexpr int kThreads = 256;

struct Args {
  const unsigned char *table;
  unsigned rows, window;
  int iters;
  unsigned long long *sink;
  unsigned long long *cycles;  // per CTA
};

__device__ __forceinline__ unsigned lcg(unsigned &s) {
  s = s * 1664525u + 1013904223u;
  return s;
}
__device__ __forceinline__ unsigned pick(unsigned &s, unsigned base, const Args &a) {
  const unsigned r = lcg(s);
  if (a.window) {
    unsigned i = base + __umulhi(r, a.window);
    return i >= a.rows ? i - a.rows : i;
  }
  return __umulhi(r, a.rows);
}

struct U8 {
  uint4 a, b;
};
__device__ __forceinline__ uint4 ldg128v(const void *p) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ldg64v(const void *p) {
  uint2 r;
  asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 lds128v(const void *p) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"((uint32_t)__cvta_generic_to_shared(p)));
  return r;
}
__device__ __forceinline__ U8 ldg256(const void *p) {
  U8 r;
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.a.x), "=r"(r.a.y), "=r"(r.a.z), "=r"(r.a.w), "=r"(r.b.x), "=r"(r.b.y), "=r"(r.b.z), "=r"(r.b.w)
               : "l"(p));
  return r;
}

template <int MODE, int U>
__global__ void __launch_bounds__(kThreads) ldg_probe(const Args a) {
  constexpr int LPR = MODE == 0 ? 4 : MODE == 1 ? 2 : 8;  // lanes per row (mode 2: per 128-byte line)
  const unsigned lane = threadIdx.x & 31;
  const unsigned grp = (blockIdx.x * kThreads + threadIdx.x) / LPR;
  const unsigned sub = lane % LPR;
  unsigned s = grp * 2654435761u + 12345u;
  unsigned acc = 0;
  const long long t0 = clock64();
  unsigned base = __umulhi(grp * 40503u, a.rows);
  for (int it = 0; it < a.iters; ++it) {
    if constexpr (MODE == 0) {
      uint4 r[U];
#pragma unroll
      for (int u = 0; u < U; ++u) r[u] = ldg128v(a.table + (size_t)pick(s, base, a) * 64 + sub * 16);
#pragma unroll
      for (int u = U - 1; u >= 0; --u) acc = (acc ^ r[u].x) * 0x9E3779B1u + (r[u].y ^ r[u].z ^ r[u].w);
    } else if constexpr (MODE == 1) {
      U8 r[U];
#pragma unroll
      for (int u = 0; u < U; ++u) r[u] = ldg256(a.table + (size_t)pick(s, base, a) * 64 + sub * 32);
#pragma unroll
      for (int u = U - 1; u >= 0; --u) acc = (acc ^ r[u].a.x) * 0x9E3779B1u + (r[u].a.y ^ r[u].a.z ^ r[u].a.w ^ r[u].b.x ^ r[u].b.y ^ r[u].b.z ^ r[u].b.w);
    } else if constexpr (MODE == 2) {
      uint4 r[U];
#pragma unroll
      for (int u = 0; u < U; ++u) r[u] = ldg128v(a.table + (size_t)(pick(s, base, a) & ~1u) * 64 + sub * 16);
#pragma unroll
      for (int u = U - 1; u >= 0; --u) acc = (acc ^ r[u].x) * 0x9E3779B1u + (r[u].y ^ r[u].z ^ r[u].w);
    } else {  // MODE 7: LDG.64, 8 lanes per row
      uint2 r[U];
#pragma unroll
      for (int u = 0; u < U; ++u) r[u] = ldg64v(a.table + (size_t)pick(s, base, a) * 64 + sub * 8);
#pragma unroll
      for (int u = U - 1; u >= 0; --u) acc = (acc ^ r[u].x) * 0x9E3779B1u + r[u].y;
    }
    base += 7;
    if (base >= a.rows) base -= a.rows;
  }
  const long long t1 = clock64();
  if (acc == 0x12345u) a.sink[0] = acc;
  if (threadIdx.x == 0) a.cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

// shared-memory table: `SROWS` rows of 64 bytes, natural [pixel][head][64 B] layout (head parity = bank half)
template <int MODE, int U>
__global__ void __launch_bounds__(kThreads) lds_probe(const Args a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const unsigned srows = a.window;  // rows resident in shared memory
  for (unsigned i = threadIdx.x; i < srows * 4; i += kThreads)
    reinterpret_cast<uint4 *>(smem)[i] = __ldg(reinterpret_cast<const uint4 *>(a.table) + i);
  __syncthreads();
  const unsigned lane = threadIdx.x & 31;
  const unsigned grp = (blockIdx.x * kThreads + threadIdx.x) / 4;
  const unsigned sub = lane & 3;
  const unsigned par = (lane >> 2) & 1;  // lane group parity
  unsigned s = grp * 2654435761u + 12345u;
  unsigned acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < a.iters; ++it) {
    uint4 r[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      unsigned row = __umulhi(lcg(s), srows);
      if constexpr (MODE == 3) row = (row & ~1u) | par;  // opposite bank halves for neighbouring lane groups
      r[u] = lds128v(smem + (size_t)row * 64 + sub * 16);
    }
#pragma unroll
    for (int u = U - 1; u >= 0; --u) acc = (acc ^ r[u].x) * 0x9E3779B1u + (r[u].y ^ r[u].z ^ r[u].w);
  }
  const long long t1 = clock64();
  if (acc == 0x12345u) a.sink[0] = acc;
  if (threadIdx.x == 0) a.cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// modes 5 / 6: one elected lane per warp issues TMA copies into a per-warp ring of U slots x 16 rows; the warp
// then reads the rows back with conflict-free LDS.128 (what a consumer would do).
template <int MODE, int U>
__global__ void __launch_bounds__(kThreads) tma_probe(const Args a, const __grid_constant__ CUtensorMap tmap) {
  constexpr int ROWS_PER_SLOT = 8;  // rows one warp consumes per LDS.128 (4 lanes per row)
  __shared__ __align__(128) unsigned char ring[kThreads / 32][U][ROWS_PER_SLOT * 64];
  __shared__ __align__(8) uint64_t bar[kThreads / 32][U];
  const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    for (int u = 0; u < U; ++u) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[warp][u])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  unsigned s = (blockIdx.x * (kThreads / 32) + warp) * 2654435761u + 12345u;
  unsigned acc = 0;
  auto issue = [&](int u) {
    if (lane == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[warp][u])), "r"(ROWS_PER_SLOT * 64)
                   : "memory");
      if constexpr (MODE == 5) {
#pragma unroll
        for (int g = 0; g < ROWS_PER_SLOT / 4; ++g) {
          const int r0 = (int)__umulhi(lcg(s), a.rows), r1 = (int)__umulhi(lcg(s), a.rows), r2 = (int)__umulhi(lcg(s), a.rows),
                    r3 = (int)__umulhi(lcg(s), a.rows);
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
                  smem_u32(&ring[warp][u][g * 256])),
              "l"(&tmap), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(&bar[warp][u]))
              : "memory");
        }
      } else {
#pragma unroll
        for (int g = 0; g < ROWS_PER_SLOT; ++g) {
          const unsigned r = __umulhi(lcg(s), a.rows);
          asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           smem_u32(&ring[warp][u][g * 64])),
                       "l"(a.table + (size_t)r * 64), "r"(64), "r"(smem_u32(&bar[warp][u]))
                       : "memory");
        }
      }
    }
  };
  const long long t0 = clock64();
  for (int u = 0; u < U; ++u) issue(u);
  for (int it = 0; it < a.iters; ++it) {
    const int u = it % U;
    mbar_wait(&bar[warp][u], (unsigned)((it / U) & 1));
    const uint4 r = *reinterpret_cast<const uint4 *>(&ring[warp][u][lane * 16]);
    acc ^= r.x ^ r.y ^ r.z ^ r.w;
    __syncwarp();
    if (it + U < a.iters) issue(u);
  }
  const long long t1 = clock64();
  if (acc == 0x12345u) a.sink[0] = acc;
  if (lane == 0 && warp == 0) a.cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template <int MODE, int U>
cudaError_t launch(int grid, const Args &a, const CUtensorMap &tmap, size_t smem) {
  if constexpr (MODE == 3 || MODE == 4) {
    cudaFuncSetAttribute(lds_probe<MODE, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    lds_probe<MODE, U><<<grid, kThreads, smem>>>(a);
  } else if constexpr (MODE == 5 || MODE == 6) {
    tma_probe<MODE, U><<<grid, kThreads>>>(a, tmap);
  } else {
    ldg_probe<MODE, U><<<grid, kThreads>>>(a);
  }
  return cudaGetLastError();
}

template <int MODE>
cudaError_t launch_u(int unroll, int grid, const Args &a, const CUtensorMap &tmap, size_t smem) {
  switch (unroll) {
    case 1: return launch<MODE, 1>(grid, a, tmap, smem);
    case 2: return launch<MODE, 2>(grid, a, tmap, smem);
    case 4: return launch<MODE, 4>(grid, a, tmap, smem);
    default: return launch<MODE, 8>(grid, a, tmap, smem);
  }
}

int main(int argc, char **argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  const int ctas_per_sm = argc > 2 ? atoi(argv[2]) : 4;
  const int unroll = argc > 3 ? atoi(argv[3]) : 4;
  unsigned window = argc > 4 ? (unsigned)atoi(argv[4]) : 0u;
  const unsigned rows = argc > 5 ? (unsigned)atoi(argv[5]) : 147312u;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  const int grid = sms * ctas_per_sm;
  unsigned char *table;
  CK(cudaMalloc(&table, (size_t)rows * 64));
  std::vector<unsigned> host((size_t)rows * 16);
  for (size_t i = 0; i < host.size(); ++i) host[i] = (unsigned)(i * 2654435761u);
  CK(cudaMemcpy(table, host.data(), host.size() * 4, cudaMemcpyHostToDevice));
  unsigned long long *sink, *cycles;
  CK(cudaMalloc(&sink, 8));
  CK(cudaMalloc(&cycles, sizeof(unsigned long long) * grid));
  size_t smem = 0;
  if (mode == 3 || mode == 4) {
    if (!window) window = 2160;  // 138,240 bytes: the two coarsest levels of the 1152x768 pyramid, all heads
    smem = (size_t)window * 64;
  }
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (mode == 5) {
    typedef CUresult (*Fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                           const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                           CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult st;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &st));
    const cuuint64_t dims[2] = {32, rows};
    const cuuint64_t strides[1] = {64};
    const cuuint32_t box[2] = {32, 1};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = ((Fn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, table, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      printf("{\"error\": \"cuTensorMapEncodeTiled %d\"}\n", (int)r);
      return 1;
    }
  }
  // rows fetched per thread per iteration
  const double rows_per_thread_iter = mode == 0 ? unroll / 4.0 : mode == 1 ? unroll / 2.0 : mode == 2 ? unroll * 2 / 8.0 : mode == 7 ? unroll / 8.0
                                      : (mode == 3 || mode == 4)                                                                  ? unroll / 4.0
                                                                                                                                  : 8.0 / 32.0;
  int iters = (mode == 5 || mode == 6) ? 4000 : 2000;
  Args a{table, rows, window, iters, sink, cycles};
  auto run = [&]() -> cudaError_t {
    switch (mode) {
      case 0: return launch_u<0>(unroll, grid, a, tmap, smem);
      case 1: return launch_u<1>(unroll, grid, a, tmap, smem);
      case 2: return launch_u<2>(unroll, grid, a, tmap, smem);
      case 3: return launch_u<3>(unroll, grid, a, tmap, smem);
      case 4: return launch_u<4>(unroll, grid, a, tmap, smem);
      case 5: return launch_u<5>(unroll, grid, a, tmap, smem);
      case 6: return launch_u<6>(unroll, grid, a, tmap, smem);
      default: return launch_u<7>(unroll, grid, a, tmap, smem);
    }
  };
  CK(run());
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(e0));
    CK(run());
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  std::vector<unsigned long long> cyc(grid);
  CK(cudaMemcpy(cyc.data(), cycles, sizeof(unsigned long long) * grid, cudaMemcpyDeviceToHost));
  unsigned long long cmax = 0;
  for (auto c : cyc) cmax = c > cmax ? c : cmax;
  const double total_rows = rows_per_thread_iter * (double)iters * (double)grid * kThreads;
  const double rows_per_clk_sm = total_rows / sms / (double)cmax;
  printf(
      "{\"mode\": %d, \"ctas_per_sm\": %d, \"unroll\": %d, \"window\": %u, \"rows\": %u, \"ms\": %.4f, \"rows_per_us\": %.1f, "
      "\"gather_GBps\": %.1f, \"cycles_max\": %llu, \"rows_per_clk_per_sm\": %.4f, \"equiv_headline_us\": %.2f}\n",
      mode, ctas_per_sm, unroll, window, rows, best, total_rows / (best * 1e3), total_rows * 64 / (best * 1e6), cmax, rows_per_clk_sm,
      // the headline call gathers 18,414 queries x 8 heads x 80 corner rows, 78 % of them live
      18414.0 * 8 * 80 * 0.78 / (total_rows / (best * 1e3)));
  return 0;
}
