// SPDX-License-Identifier: Apache-2.0
//
// value_proj_sm100.cu -- the two GEMM-shaped neighbours of the sampling kernels for NVIDIA B200 (sm_100a):
//
//     value[r, :] = key_padding_mask[r] ? 0 : x[r, :] @ W^T + bias                 (msda_b200_value_proj)
//     out[r, :]   = round(attended[r, :] @ W^T + bias) + residual[r, :]             (msda_b200_output_proj)
//
// i.e. nn.Linear + masked_fill + head split (/root/reference/codetr/multi_scale_deformable_attention.py:173-176)
// and output_proj + inference-mode dropout + residual (:212-218) of the calling module, each in ONE kernel; the
// first one's output already is the [B, S, M, D] tensor the sampling kernels read.
// These are the only GEMMs next to the hot path, so they run on the 5th-generation tensor cores, hand-written:
//
//   * operands staged by TMA (cp.async.bulk.tensor.2d, 128-byte swizzle) into shared memory, one mbarrier per
//     64-wide K chunk so the first MMAs start while the later chunks are still in flight;
//   * one elected thread issues tcgen05.mma (cta_group::1, kind::f16, UMMA 128 x N x 16), fp32 accumulators in
//     tensor memory, completion signalled with tcgen05.commit on mbarriers;
//   * epilogue warps read their TMEM lane quarter with tcgen05.ld (32x32b.x32), add the bias, zero the padded
//     rows (or add the residual), round to the 16-bit element type, stage the result in 128-byte-swizzled shared
//     memory and write it back with TMA stores (rows past the end are clipped by the TMA).
//
// Two kernels share those pieces: value_proj_persistent_kernel (the default: persistent CTAs, resident weights,
// x ring, double-buffered accumulators, warp-specialised; described at its definition) and value_proj_kernel (the
// first version, one tile per CTA, kept selectable for A/B runs).  DESIGN.md section 4.6 has the measurements.
//
// Nothing here allocates, synchronises or reads device memory on the host; the tensor maps are encoded per
// call on the host (cuTensorMapEncodeTiled through cudaGetDriverEntryPoint, ~1 us) and passed as
// __grid_constant__ parameters.

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "msda_b200.h"
#include "msda_internal.hpp"

namespace {

constexpr int kTileRows = 128;    // UMMA M: one accumulator row per TMEM lane
constexpr int kChunkK = 64;       // 16-bit elements per 128-byte swizzle row
constexpr int kUmmaK = 16;        // K of one tcgen05.mma for 16-bit inputs
constexpr int kMaxChunks = 4;     // K <= 256
constexpr int kMaxN = 256;
constexpr int kEpilogueWarps = 4;
constexpr int kThreads = (kEpilogueWarps + 1) * 32;  // + one producer / MMA warp

struct ProjParams {
  const void *bias;            // [N] in the element type, or nullptr
  const unsigned char *mask;   // [rows], non-zero = padded key (row of zeros), or nullptr
  const void *residual;        // [rows, n_total] in the element type, added after the bias, or nullptr (output_proj mode)
  int rows, K, N;              // N = output columns of ONE CTA (n_total / n_split)
  int n_total, n_split;        // all output columns; column blocks a row tile is split into (1, 2 or 4)
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Spin on the barrier's phase; a transfer that never completes (a bad tensor map) traps instead of hanging
// the GPU until the watchdog.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
  unsigned done = 0;
  for (unsigned spins = 0; !done; ++spins) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (spins > (1u << 26)) __trap();
  }
}

__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
// the same load delivered to the same shared-memory offset (and mbarrier) of every CTA of the cluster named in `mask`
__device__ __forceinline__ void tma_load_2d_multicast(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar,
                                                      unsigned short mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ unsigned cluster_cta_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, const void *src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(src)),
               "r"(c0), "r"(c1)
               : "memory");
}

// Shared-memory matrix descriptor of a K-major tile whose rows are 128 bytes, 128-byte swizzled (what the
// TMA wrote): start address >> 4 in bits [0,14), stride between 8-row groups (1024 B) >> 4 in bits [32,46),
// descriptor version 1 in bits [46,48), layout type 2 (SWIZZLE_128B) in bits [61,64).
__device__ __forceinline__ uint64_t umma_desc_k_major_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// D[tmem] (+)= A[smem] * B[smem]^T, issued by one thread for the whole CTA
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_load_32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// same without the wait: the caller overlaps it with work on registers it already holds, then waits
__device__ __forceinline__ void tmem_load_32_async(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// (a, b) = (x0 + y0, x1 + y1), one packed fp32 instruction; each half is an ordinary round-to-nearest add
__device__ __forceinline__ void add2(float &a, float &b, float x0, float x1, float y0, float y1) {
  unsigned long long x, y, r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(x0), "f"(x1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(y0), "f"(y1));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(x), "l"(y));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r));
}

template <bool BF16>
__device__ __forceinline__ float elem_to_float(unsigned short bits) {
  if constexpr (BF16) return __uint_as_float((unsigned)bits << 16);
  else return __half2float(__ushort_as_half(bits));
}
template <bool BF16>
__device__ __forceinline__ unsigned pack_pair(float a, float b) {
  if constexpr (BF16) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const unsigned *>(&h);
  } else {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const unsigned *>(&h);
  }
}

// Dynamic shared memory (1024-byte aligned, the swizzle atom):
//   [0, tile_bytes)        x tile: K/64 chunks of 128 rows x 128 B; reused as the output staging tile
//                          (N/64 chunks of 128 rows x 128 B), tile_bytes = 16 KB * max(K, N) / 64
//   [.., + K/64 * N*128)   weight: K/64 chunks of N rows x 128 B
//   bias as fp32 [N], then the barriers and the TMEM base address
template <bool BF16>
__global__ void __launch_bounds__(kThreads, 1)
value_proj_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                  const __grid_constant__ CUtensorMap map_out, const ProjParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // the swizzle atom needs 1024-byte alignment; the launch asks for 1 KB of slack to round up into
  unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int chunks = p.K / kChunkK;
  const int out_chunks = p.N / kChunkK;
  const int tile_bytes = kTileRows * 128 * max(chunks, out_chunks);
  unsigned char *x_tile = smem;
  unsigned char *w_tile = smem + tile_bytes;
  float *bias_f = reinterpret_cast<float *>(w_tile + (size_t)chunks * p.N * 128);
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(bias_f + kMaxN);  // [kMaxChunks]
  uint64_t *done_bar = full_bar + kMaxChunks;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done_bar + 1);

  const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
  const int row0 = (int)blockIdx.x * kTileRows;
  const uint32_t tmem_cols = p.N <= 32 ? 32u : p.N <= 64 ? 64u : p.N <= 128 ? 128u : 256u;

  // ---- set-up: barriers + TMEM allocation (producer warp), bias to fp32 (epilogue warps) ----
  if (warp == kEpilogueWarps) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_out) : "memory");
      for (int c = 0; c < chunks; ++c) mbar_init(&full_bar[c], 1);
      mbar_init(done_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    for (int n = (int)threadIdx.x; n < p.N; n += kEpilogueWarps * 32) {
      bias_f[n] = p.bias ? elem_to_float<BF16>(static_cast<const unsigned short *>(p.bias)[n]) : 0.f;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kEpilogueWarps) {
    // ---- producer + MMA issuer: one thread ----
    if (lane == 0) {
      const unsigned chunk_bytes = (unsigned)(kTileRows * 128 + p.N * 128);
      for (int c = 0; c < chunks; ++c) {
        mbar_expect_tx(&full_bar[c], chunk_bytes);
        tma_load_2d(x_tile + (size_t)c * kTileRows * 128, &map_x, c * kChunkK, row0, &full_bar[c]);
        tma_load_2d(w_tile + (size_t)c * p.N * 128, &map_w, c * kChunkK, 0, &full_bar[c]);
      }
      // instruction descriptor: D = fp32 (bits 4-5 = 1), A / B format (bits 7-9 / 10-12: 0 = f16, 1 = bf16),
      // both operands K-major (bits 15, 16 = 0), N >> 3 in bits 17-22, M >> 4 in bits 24-28
      const uint32_t fmt = BF16 ? 1u : 0u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(kTileRows >> 4) << 24);
      for (int c = 0; c < chunks; ++c) {
        mbar_wait(&full_bar[c], 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_addr = smem_u32(x_tile + (size_t)c * kTileRows * 128);
        const uint32_t b_addr = smem_u32(w_tile + (size_t)c * p.N * 128);
#pragma unroll
        for (int k = 0; k < kChunkK / kUmmaK; ++k) {
          // inside the 128-byte swizzle row a K step of 16 elements is 32 bytes of start address
          umma_f16(tmem_base, umma_desc_k_major_sw128(a_addr + k * kUmmaK * 2), umma_desc_k_major_sw128(b_addr + k * kUmmaK * 2),
                   idesc, (c | k) != 0 ? 1u : 0u);
        }
      }
      umma_commit(done_bar);  // arrives when every MMA above has written TMEM (and read its operands)
    }
    __syncwarp();
  } else {
    // ---- epilogue: TMEM -> registers -> (+bias, mask, round) -> swizzled staging tile ----
    mbar_wait(done_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int rl = warp * 32 + lane;  // row inside the tile = TMEM lane
    const int r = row0 + rl;
    const bool padded = p.mask != nullptr && r < p.rows && p.mask[r] != 0;
    unsigned char *stage = x_tile;
    for (int c32 = 0; c32 < p.N / 32; ++c32) {
      uint32_t v[32];
      tmem_load_32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c32 * 32), v);
      unsigned char *row_base = stage + (size_t)(c32 >> 1) * kTileRows * 128 + (size_t)rl * 128;
#pragma unroll
      for (int i = 0; i < 4; ++i) {  // four 16-byte pieces = 8 elements each
        uint4 o;
        unsigned *ow = reinterpret_cast<unsigned *>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = c32 * 32 + i * 8 + j * 2;
          const float a = padded ? 0.f : __uint_as_float(v[i * 8 + j * 2]) + bias_f[col];
          const float b = padded ? 0.f : __uint_as_float(v[i * 8 + j * 2 + 1]) + bias_f[col + 1];
          ow[j] = pack_pair<BF16>(a, b);
        }
        const int piece = (c32 & 1) * 4 + i;
        *reinterpret_cast<uint4 *>(row_base + ((piece ^ (rl & 7)) << 4)) = o;
      }
    }
    // make the generic-proxy writes visible to the TMA (async proxy), then one thread stores the tile
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("bar.sync 1, %0;" ::"n"(kEpilogueWarps * 32) : "memory");
    if (threadIdx.x == 0) {
      for (int c = 0; c < out_chunks; ++c) tma_store_2d(&map_out, stage + (size_t)c * kTileRows * 128, c * kChunkK, row0);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem must outlive the reads
    }
  }

  // ---- teardown: everyone is done with TMEM, the allocating warp frees it ----
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kEpilogueWarps) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ---------------------------------------------------------------------------
// Persistent, warp-specialised variant (the default).  One CTA per SM loops over row tiles:
//   warp 8  TMA producer: the x tiles as a ring of 64-wide K chunks (kStages x 16 KB) running ahead of the
//           MMAs, and the weight matrix once (resident for the CTA's lifetime), interleaved chunk by chunk with
//           the first tile so the first MMAs start after 48 KB instead of 192 KB; the first loads are issued
//           before the CTA-wide set-up barrier (they only need the mbarriers this warp initialised itself);
//   warp 9  MMA issuer: tcgen05.mma into one of TWO accumulator buffers in tensor memory; tcgen05.commit frees
//           each ring slot and publishes each finished accumulator;
//   warps 0-7  epilogue: tcgen05.ld of their lane quarter (warps w and w+4 share a quarter and split the
//           columns), bias / mask / rounding, 16-byte conflict-free stores into the warp's own 4 KB swizzled
//           staging slab, written back by the warp's own TMA store (no CTA-wide barrier).  The accumulator is
//           released as soon as its last columns are in registers, so the epilogue of tile i overlaps the
//           loads and MMAs of tile i+1.
// Measured phase timeline (tools/vproj_trace.py, -DMSDA_VPROJ_TRACE) is in DESIGN.md section 4.
// ---------------------------------------------------------------------------
#ifdef MSDA_VPROJ_TRACE
// Debug build only: %globaltimer stamps of the phases of every CTA, read back by msda_b200_debug_vproj_trace.
constexpr int kTraceSlots = 8;
__device__ unsigned long long g_vproj_trace[1024][kTraceSlots];
__device__ __forceinline__ void trace_stamp(int slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  if (blockIdx.x < 1024) g_vproj_trace[blockIdx.x][slot] = t;
}
#if MSDA_VPROJ_TRACE == 2  // epilogue detail of the first tile instead of the kernel phases
#define VPROJ_TRACE(slot)
#define VPROJ_TRACE_EPI(slot) trace_stamp(slot)
#else
#define VPROJ_TRACE(slot) trace_stamp(slot)
#define VPROJ_TRACE_EPI(slot)
#endif
#else
#define VPROJ_TRACE(slot)
#define VPROJ_TRACE_EPI(slot)
#endif

constexpr int kStages = 4;          // x ring slots (one K chunk of one tile each)
constexpr int kStageChunks = 2;     // output staging: 2 x 16 KB = one 4 KB slab (32 rows x 64 columns) per epilogue warp
constexpr int kPEpilogueWarps = 8;
constexpr int kProducerWarp = kPEpilogueWarps, kMmaWarp = kPEpilogueWarps + 1;
constexpr int kPersistentThreads = (kPEpilogueWarps + 2) * 32;

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Dynamic shared memory (1024-byte aligned): weight chunks (K/64 x N x 128 B), x ring (kStages x 16 KB), output
// staging (kStageChunks x 16 KB), bias as fp32 [N], barriers, TMEM base address.
template <bool BF16, bool RESIDUAL>
__global__ void __launch_bounds__(kPersistentThreads, 1)
value_proj_persistent_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                             const __grid_constant__ CUtensorMap map_out, const ProjParams p, int pdl, int cluster) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int chunks = p.K / kChunkK;
  const int out_chunks = p.N / kChunkK;
  unsigned char *w_tile = smem;
  unsigned char *ring = smem + (size_t)chunks * p.N * 128;
  unsigned char *staging = ring + (size_t)kStages * kTileRows * 128;
  float *bias_f = reinterpret_cast<float *>(staging + (size_t)kStageChunks * kTileRows * 128);
  uint64_t *w_full = reinterpret_cast<uint64_t *>(bias_f + kMaxN);  // [kMaxChunks]
  uint64_t *a_full = w_full + kMaxChunks;   // [kStages]
  uint64_t *a_empty = a_full + kStages;     // [kStages]
  uint64_t *acc_full = a_empty + kStages;   // [2]
  uint64_t *acc_empty = acc_full + 2;       // [2]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);

  const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
  if (threadIdx.x == 0) VPROJ_TRACE(0);  // kernel entry
  const int tiles = (p.rows + kTileRows - 1) / kTileRows;
  // n_split > 1 (few row tiles): n_split CTAs share a row tile, each owns N = n_total / n_split output columns
  // starting at n0 and loads only that slice of the weight matrix; the host launches exactly tiles * n_split CTAs
  const int tile0 = (int)blockIdx.x / p.n_split, tile_step = (int)gridDim.x / p.n_split;
  const int n0 = ((int)blockIdx.x % p.n_split) * p.N;
  const uint32_t acc_cols = p.N <= 32 ? 32u : p.N <= 64 ? 64u : p.N <= 128 ? 128u : 256u;  // per accumulator buffer
  const uint32_t tmem_cols = 2 * acc_cols;

  // producer state (lives across the set-up barrier)
  int p_stage = 0;
  unsigned p_phase = 0;
  auto load_x_chunk = [&](int tile, int c) {
    mbar_wait(&a_empty[p_stage], p_phase ^ 1u);  // a fresh barrier passes the wait on the opposite parity
    mbar_expect_tx(&a_full[p_stage], (unsigned)(kTileRows * 128));
    tma_load_2d(ring + (size_t)p_stage * kTileRows * 128, &map_x, c * kChunkK, tile * kTileRows, &a_full[p_stage]);
    if (++p_stage == kStages) {
      p_stage = 0;
      p_phase ^= 1u;
    }
  };

  if (warp == kProducerWarp) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_out) : "memory");
      for (int c = 0; c < kMaxChunks; ++c) mbar_init(&w_full[c], 1);
      for (int s = 0; s < kStages; ++s) {
        mbar_init(&a_full[s], 1);
        mbar_init(&a_empty[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&acc_full[b], 1);
        mbar_init(&acc_empty[b], kPEpilogueWarps);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      // first tile + weights, chunk by chunk, before the rest of the CTA has finished its set-up
      if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
      // cluster > 1: the weights arrive as multicast slices issued by every CTA of the cluster after the cluster-wide
      // barrier below (map_w's box is then N / cluster rows); every CTA expects the whole matrix on its own barriers
      for (int c = 0; c < chunks; ++c) {
        if (tile0 < tiles) load_x_chunk(tile0, c);
        mbar_expect_tx(&w_full[c], (unsigned)(p.N * 128));
        if (cluster == 1) tma_load_2d(w_tile + (size_t)c * p.N * 128, &map_w, c * kChunkK, n0, &w_full[c]);
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int n = (int)threadIdx.x; n < p.N; n += kPEpilogueWarps * 32) {
      bias_f[n] = p.bias ? elem_to_float<BF16>(static_cast<const unsigned short *>(p.bias)[n0 + n]) : 0.f;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (cluster > 1) cluster_sync_all();  // every CTA's barriers exist before anybody multicasts into them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) VPROJ_TRACE(1);  // set-up done (barriers, TMEM, bias)
  if (cluster > 1 && warp == kProducerWarp && lane == 0) {
    // this CTA's slice of every weight chunk, delivered to all CTAs of the cluster: one L2 read per cluster instead
    // of one per SM (148 SMs pulling the same 128 KB is what bounds the first tile)
    const int slice_rows = p.N / cluster;
    const int rank = (int)cluster_cta_rank();
    const unsigned short mask = (unsigned short)((1u << cluster) - 1u);
    for (int c = 0; c < chunks; ++c) {
      tma_load_2d_multicast(w_tile + (size_t)c * p.N * 128 + (size_t)rank * slice_rows * 128, &map_w, c * kChunkK, rank * slice_rows,
                            &w_full[c], mask);
    }
  }
  // let the next kernel of the stream start its own launch / set-up (it waits for our memory in its own
  // griddepcontrol.wait; kernels launched without the attribute are unaffected)
  if (pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == kProducerWarp) {
    // ===== TMA producer: the remaining tiles =====
    if (lane == 0) {
      for (int tile = tile0 + tile_step; tile < tiles; tile += tile_step) {
        for (int c = 0; c < chunks; ++c) load_x_chunk(tile, c);
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t fmt = BF16 ? 1u : 0u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(kTileRows >> 4) << 24);
      int stage = 0;
      unsigned phase = 0;
      int t = 0;
      for (int tile = tile0; tile < tiles; tile += tile_step, ++t) {
        const int buf = t & 1;
        mbar_wait(&acc_empty[buf], (((unsigned)t >> 1) & 1u) ^ 1u);  // the epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t acc = tmem_base + (uint32_t)buf * acc_cols;
        for (int c = 0; c < chunks; ++c) {
          if (t == 0) {
            mbar_wait(&w_full[c], 0);
            if (c == 0) VPROJ_TRACE(2);  // first weight chunk landed
          }
          mbar_wait(&a_full[stage], phase);
          if (t == 0 && c == 0) VPROJ_TRACE(3);  // first x chunk landed
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_addr = smem_u32(ring + (size_t)stage * kTileRows * 128);
          const uint32_t b_addr = smem_u32(w_tile + (size_t)c * p.N * 128);
#pragma unroll
          for (int k = 0; k < kChunkK / kUmmaK; ++k) {
            umma_f16(acc, umma_desc_k_major_sw128(a_addr + k * kUmmaK * 2), umma_desc_k_major_sw128(b_addr + k * kUmmaK * 2), idesc,
                     (c | k) != 0 ? 1u : 0u);
          }
          umma_commit(&a_empty[stage]);  // the slot is free once these MMAs have read it
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(&acc_full[buf]);
        if (t == 0) VPROJ_TRACE(4);  // first tile's MMAs issued
      }
      if (t == 0) {  // a CTA without tiles still receives the cluster's weight multicasts: let them land before exit
        for (int c = 0; c < chunks; ++c) mbar_wait(&w_full[c], 0);
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue =====
    // warp w reads TMEM lanes [32*(w%4), +32) (the hardware ties a warp to that lane quarter); warps w and w+4
    // share a quarter and split the 64-column output chunks between them (even / odd).  Every warp owns a 4 KB
    // slab of the staging buffer (its 32 rows x 64 columns, 128-byte swizzled) and writes it back with its own
    // TMA store, so the eight warps never wait for each other: no CTA-wide barrier in the epilogue, and a slab's
    // TMA read overlaps the TMEM load and the conversion of the warp's next chunk.
    const int quarter = warp & 3, half = warp >> 2;
    const int rl = quarter * 32 + lane;
    unsigned char *slab = staging + (size_t)warp * (32 * 128);
    unsigned char *slab_row = slab + (size_t)lane * 128;
    int t = 0;
    for (int tile = tile0; tile < tiles; tile += tile_step, ++t) {
      const int buf = t & 1;
      const int r = tile * kTileRows + rl;
      const bool padded = p.mask != nullptr && r < p.rows && p.mask[r] != 0;
      const unsigned char *res_row = nullptr;  // RESIDUAL: this thread's row of the tensor added after the bias
      if constexpr (RESIDUAL) {
        if (r < p.rows) res_row = static_cast<const unsigned char *>(p.residual) + ((size_t)r * p.n_total + n0) * 2;
      }
      mbar_wait(&acc_full[buf], ((unsigned)t >> 1) & 1u);
      if (t == 0 && threadIdx.x == 0) VPROJ_TRACE(5);  // first accumulator complete
      if (t == 0 && threadIdx.x == 0) VPROJ_TRACE_EPI(0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t acc = tmem_base + (uint32_t)buf * acc_cols + ((uint32_t)(quarter * 32) << 16);
      if (half >= out_chunks) {
        // N == 64: the second warp of the quarter has no chunk, but the accumulator hand-back counts all 8 warps
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
      }
      for (int oc = half; oc < out_chunks; oc += 2) {  // this warp's output chunks (64 columns each)
        const bool tr = t == 0 && threadIdx.x == 0 && oc == half;
        uint32_t v0[32], v1[32];  // named (not indexed) so they stay in registers
        tmem_load_32_async(acc + (uint32_t)(oc * 64), v0);
        tmem_load_32_async(acc + (uint32_t)(oc * 64 + 32), v1);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (tr) VPROJ_TRACE_EPI(1);
        if (oc + 2 >= out_chunks) {
          // the warp's share of the accumulator is in registers: hand the buffer back to the MMA issuer
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        // the slab's previous TMA store (issued by lane 0) must have finished reading it
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        if (tr) VPROJ_TRACE_EPI(2);
        // the chunk's 64 bias values, loaded before the first slab store: the compiler cannot prove that the
        // slab stores do not alias the bias array, so loads left inside the loop are serialised behind them
        float4 bias_r[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) bias_r[i] = *reinterpret_cast<const float4 *>(bias_f + oc * 64 + i * 4);
        auto emit = [&](const uint32_t(&cur)[32], int g) {  // g: which 32-column half of the chunk
#pragma unroll
          for (int i = 0; i < 4; ++i) {  // four 16-byte pieces = 8 elements each
            const int piece = g * 4 + i;
            uint4 *dst = reinterpret_cast<uint4 *>(slab_row + ((piece ^ (lane & 7)) << 4));
            const float4 b0 = bias_r[g * 8 + i * 2], b1 = bias_r[g * 8 + i * 2 + 1];
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            uint4 res = make_uint4(0u, 0u, 0u, 0u);
            if constexpr (RESIDUAL) {
              if (res_row != nullptr) res = __ldg(reinterpret_cast<const uint4 *>(res_row + (size_t)(oc * 64 + g * 32 + i * 8) * 2));
            }
            const unsigned rw[4] = {res.x, res.y, res.z, res.w};
            uint4 o;
            unsigned *ow = reinterpret_cast<unsigned *>(&o);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              // accumulator + bias in fp32, two columns per instruction (add.f32x2 -> SASS FADD2)
              float a, b;
              add2(a, b, __uint_as_float(cur[i * 8 + j * 2]), __uint_as_float(cur[i * 8 + j * 2 + 1]), bb[j * 2], bb[j * 2 + 1]);
              if constexpr (RESIDUAL) {
                // Linear output rounded to the element type first, then the residual added and rounded again: the
                // same two roundings as output_proj followed by a separate add in 16 bits
                const unsigned lin = pack_pair<BF16>(a, b);
                a = elem_to_float<BF16>((unsigned short)(lin & 0xffffu)) + elem_to_float<BF16>((unsigned short)(rw[j] & 0xffffu));
                b = elem_to_float<BF16>((unsigned short)(lin >> 16)) + elem_to_float<BF16>((unsigned short)(rw[j] >> 16));
              }
              ow[j] = padded ? 0u : pack_pair<BF16>(a, b);  // a padded key: zeros whatever the GEMM produced
            }
            *dst = o;  // no branch in this loop: the 16 bias loads of a chunk are all issued up front
          }
        };
        emit(v0, 0);
        emit(v1, 1);
        // generic-proxy writes -> visible to the TMA (async proxy), then lane 0 stores the slab
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (tr) VPROJ_TRACE_EPI(3);
        if (lane == 0) {
          tma_store_2d(&map_out, slab, n0 + oc * kChunkK, tile * kTileRows + quarter * 32);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (tr) VPROJ_TRACE_EPI(4);
      }
      if (t == 0 && threadIdx.x == 0) VPROJ_TRACE(6);  // first tile handed to the TMA
      if (t == 0 && threadIdx.x == 0) VPROJ_TRACE_EPI(7);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem must outlive the reads
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
  if (cluster > 1) cluster_sync_all();  // nobody leaves while a peer could still be reading or writing its shared memory
  if (threadIdx.x == 0) VPROJ_TRACE(7);  // exit
}

size_t persistent_smem_bytes(int K, int N) {
  return (size_t)(K / kChunkK) * N * 128 + (size_t)(kStages + kStageChunks) * kTileRows * 128 + kMaxN * sizeof(float) +
         (kMaxChunks + 2 * kStages + 4) * sizeof(uint64_t) + 16;
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static std::atomic<void *> cached{nullptr};
  void *fn = cached.load(std::memory_order_acquire);
  if (!fn) {
    cudaDriverEntryPointQueryResult status;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &status) != cudaSuccess ||
        status != cudaDriverEntryPointSuccess) {
      return nullptr;
    }
    cached.store(fn, std::memory_order_release);
  }
  return reinterpret_cast<EncodeTiledFn>(fn);
}

// [outer, inner] row-major 16-bit matrix, box = 64 x box_outer elements, 128-byte swizzle
bool make_map(CUtensorMap *map, const void *base, bool bf16, uint64_t inner, uint64_t outer, uint32_t box_outer) {
  EncodeTiledFn encode = encode_tiled_fn();
  if (!encode) return false;
  const cuuint64_t dims[2] = {inner, outer};
  const cuuint64_t strides[1] = {inner * 2};
  const cuuint32_t box[2] = {(cuuint32_t)kChunkK, box_outer};
  const cuuint32_t elem_strides[2] = {1, 1};
  return encode(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), dims,
                strides, box, elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

size_t proj_smem_bytes(int K, int N) {
  const int chunks = K / kChunkK, out_chunks = N / kChunkK;
  const size_t tile = (size_t)kTileRows * 128 * (chunks > out_chunks ? chunks : out_chunks);
  return tile + (size_t)chunks * N * 128 + kMaxN * sizeof(float) + (kMaxChunks + 1) * sizeof(uint64_t) + 16;
}

}  // namespace

extern "C" {

int msda_b200_value_proj_supported(int64_t in_features, int64_t out_features, int dtype) {
  return (dtype == MSDA_F16 || dtype == MSDA_BF16) && in_features >= kChunkK && in_features <= kMaxChunks * kChunkK &&
                 in_features % kChunkK == 0 && out_features >= kChunkK && out_features <= kMaxN && out_features % kChunkK == 0
             ? 1
             : 0;
}

static int launch_projection(const void *x, const void *weight, const void *bias, const unsigned char *key_padding_mask,
                             const void *residual, void *value, int64_t rows, int64_t in_features, int64_t out_features, int dtype,
                             void *stream_) {
  if (rows < 0 || in_features <= 0 || out_features <= 0) return MSDA_ERR_BAD_SHAPE;
  if (dtype != MSDA_F16 && dtype != MSDA_BF16 && dtype != MSDA_F32 && dtype != MSDA_F64) return MSDA_ERR_BAD_DTYPE;
  if (!msda_b200_value_proj_supported(in_features, out_features, dtype)) return MSDA_ERR_UNSUPPORTED;
  if (rows == 0) return MSDA_OK;
  if (rows > (int64_t)INT32_MAX - kTileRows) return MSDA_ERR_BAD_SHAPE;
  if (!x || !weight || !value) return MSDA_ERR_NULL_POINTER;
  if (((uintptr_t)x | (uintptr_t)weight | (uintptr_t)value) & 15u) return MSDA_ERR_UNSUPPORTED;  // TMA base alignment
  if (bias && ((uintptr_t)bias & 1u)) return MSDA_ERR_UNSUPPORTED;
  if (residual && ((uintptr_t)residual & 15u)) return MSDA_ERR_UNSUPPORTED;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool bf16 = dtype == MSDA_BF16;
  const int K = (int)in_features, N = (int)out_features;

  const char *single = getenv("MSDA_B200_VPROJ_SINGLE_TILE");
  const bool single_tile = single && *single == '1' && residual == nullptr;
  CUtensorMap map_x, map_w, map_out;
  if (!make_map(&map_x, x, bf16, (uint64_t)K, (uint64_t)rows, kTileRows) || !make_map(&map_w, weight, bf16, (uint64_t)K, (uint64_t)N, (uint32_t)N) ||
      // output boxes: a whole 128-row chunk for the single-tile kernel, one warp's 32 rows for the persistent one
      !make_map(&map_out, value, bf16, (uint64_t)N, (uint64_t)rows, single_tile ? kTileRows : 32)) {
    return MSDA_ERR_UNSUPPORTED;
  }
  ProjParams p;
  p.bias = bias;
  p.mask = key_padding_mask;
  p.residual = residual;
  p.rows = (int)rows;
  p.K = K;
  p.N = N;
  p.n_total = N;
  p.n_split = 1;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return MSDA_ERR_UNSUPPORTED;
  static std::atomic<int> sm_count[64];
  int sms = sm_count[dev].load(std::memory_order_acquire);
  if (sms == 0) {
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    sm_count[dev].store(sms, std::memory_order_release);
  }
  const unsigned tiles = (unsigned)((rows + kTileRows - 1) / kTileRows);
  // opt in to > 48 KB of dynamic shared memory once per (device, kernel)
  static std::atomic<int> attr_set[64][6];
  auto opt_in = [&](const void *fn, int slot, size_t bytes) -> cudaError_t {
    if (attr_set[dev][slot].load(std::memory_order_acquire)) return cudaSuccess;
    const cudaError_t ae = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (ae == cudaSuccess) attr_set[dev][slot].store(1, std::memory_order_release);
    return ae;
  };
  if (single_tile) {
    // one tile per CTA, output staged in shared memory and written by TMA (the first version; kept for A/B runs)
    const size_t smem = proj_smem_bytes(K, N) + 1024;  // + slack for the 1024-byte round-up in the kernel
    auto kernel = bf16 ? value_proj_kernel<true> : value_proj_kernel<false>;
    const cudaError_t ae = opt_in(reinterpret_cast<const void *>(kernel), bf16 ? 1 : 0, proj_smem_bytes(kMaxChunks * kChunkK, kMaxN) + 1024);
    if (ae != cudaSuccess) return (int)ae;
    kernel<<<tiles, kThreads, smem, stream>>>(map_x, map_w, map_out, p);
    msda_detail::set_last_variant(bf16 ? "value_proj<bf16>/tcgen05/single-tile" : "value_proj<f16>/tcgen05/single-tile");
  } else {
    const size_t smem = persistent_smem_bytes(K, N) + 1024;
    auto kernel = residual ? (bf16 ? value_proj_persistent_kernel<true, true> : value_proj_persistent_kernel<false, true>)
                           : (bf16 ? value_proj_persistent_kernel<true, false> : value_proj_persistent_kernel<false, false>);
    const cudaError_t ae = opt_in(reinterpret_cast<const void *>(kernel), 2 + (bf16 ? 1 : 0) + (residual ? 2 : 0),
                                  persistent_smem_bytes(kMaxChunks * kChunkK, kMaxN) + 1024);
    if (ae != cudaSuccess) return (int)ae;
    // thread-block clusters (opt-in, MSDA_B200_VPROJ_CLUSTER = 2 or 4): the CTAs of a cluster share one L2 read of the
    // weight matrix through TMA multicast.  Measured slower than every CTA loading its own copy (6.5 -> 8.2 us with
    // pairs, 14.0 us with clusters of 4 at the headline shape: the two cluster-wide barriers and the later start of the
    // weight loads cost more than the 9.5 MB of L2 reads they save), so the default is 1.
    int cluster = 1;
    if (const char *ce = getenv("MSDA_B200_VPROJ_CLUSTER")) cluster = atoi(ce);
    if (cluster != 1 && cluster != 2 && cluster != 4) cluster = 1;
    if ((N / cluster) % 8 != 0 || N / cluster > 256) cluster = 1;
    if (cluster > 1 && !make_map(&map_w, weight, bf16, (uint64_t)K, (uint64_t)N, (uint32_t)(N / cluster))) return MSDA_ERR_UNSUPPORTED;
    unsigned grid = tiles < (unsigned)sms ? tiles : (unsigned)sms;
    // few row tiles (R50-sized encoders, decoder queries): split every tile's output columns over 2 or 4 CTAs so more
    // SMs work and each ingests only its slice of the weight matrix (MSDA_B200_VPROJ_NSPLIT=1/2/4 forces it)
    int n_split = 1;
    if (cluster == 1) {
      if (tiles * 4 <= (unsigned)sms && N % 256 == 0) n_split = 4;
      else if (tiles * 2 <= (unsigned)sms && N % 128 == 0) n_split = 2;
      if (const char *ne = getenv("MSDA_B200_VPROJ_NSPLIT")) {
        const int forced = atoi(ne);
        if ((forced == 1 || forced == 2 || forced == 4) && N % (64 * forced) == 0) n_split = forced;
      }
    }
    if (n_split > 1) {
      p.N = N / n_split;
      p.n_split = n_split;
      // CTA j owns column block j % n_split for good and walks the row tiles j / n_split, + grid / n_split, ...
      grid = tiles * (unsigned)n_split;
      const unsigned fit = (unsigned)sms / (unsigned)n_split * (unsigned)n_split;
      if (grid > fit) grid = fit;
      if (!make_map(&map_w, weight, bf16, (uint64_t)K, (uint64_t)N, (uint32_t)p.N)) return MSDA_ERR_UNSUPPORTED;
    }
    if (cluster > 1) {
      grid = (grid + cluster - 1) / cluster * cluster;            // whole clusters; surplus CTAs only relay weights
      const unsigned fit = (unsigned)sms / cluster * cluster;
      if (grid > fit) grid = fit;
    }
    // programmatic dependent launch: this kernel's set-up overlaps the tail of the previous kernel of the
    // stream, and it releases its own dependents right after set-up (MSDA_B200_PDL=0 launches the plain way)
    const char *pdl_env = getenv("MSDA_B200_PDL");
    const int pdl = (pdl_env && *pdl_env == '0') ? 0 : 1;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kPersistentThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    unsigned n_attr = 0;
    if (pdl) {
      attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
      ++n_attr;
    }
    if (cluster > 1) {
      attr[n_attr].id = cudaLaunchAttributeClusterDimension;
      attr[n_attr].val.clusterDim.x = (unsigned)cluster;
      attr[n_attr].val.clusterDim.y = 1;
      attr[n_attr].val.clusterDim.z = 1;
      ++n_attr;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n_attr;
    const cudaError_t le = cudaLaunchKernelEx(&cfg, kernel, map_x, map_w, map_out, p, pdl, cluster);
    if (le != cudaSuccess) return (int)le;
    char name[96];
    snprintf(name, sizeof(name), "%s<%s>/tcgen05/cluster%d/nsplit%d/persistent", residual ? "output_proj" : "value_proj", bf16 ? "bf16" : "f16",
             cluster, n_split);
    msda_detail::set_last_variant(name);
  }
  msda_detail::launch_count.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

int msda_b200_value_proj(const void *x, const void *weight, const void *bias, const unsigned char *key_padding_mask, void *value,
                         int64_t rows, int64_t in_features, int64_t out_features, int dtype, unsigned flags, void *stream) {
  (void)flags;
  return launch_projection(x, weight, bias, key_padding_mask, nullptr, value, rows, in_features, out_features, dtype, stream);
}

int msda_b200_output_proj(const void *attended, const void *weight, const void *bias, const void *residual, void *out, int64_t rows,
                          int64_t in_features, int64_t out_features, int dtype, unsigned flags, void *stream) {
  (void)flags;
  if (rows > 0 && !residual) return MSDA_ERR_NULL_POINTER;
  return launch_projection(attended, weight, bias, nullptr, residual, out, rows, in_features, out_features, dtype, stream);
}

#ifdef MSDA_VPROJ_TRACE
int msda_b200_debug_vproj_trace(unsigned long long *host_out, int ctas) {
  if (ctas > 1024) ctas = 1024;
  return (int)cudaMemcpyFromSymbol(host_out, g_vproj_trace, sizeof(unsigned long long) * kTraceSlots * (size_t)ctas);
}
#endif

}  // extern "C"
