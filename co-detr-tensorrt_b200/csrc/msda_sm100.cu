// msda_sm100.cu -- multi-scale deformable attention for NVIDIA B200 (sm_100a): forward, opt-in
// producer-fused forward, backward, and the C ABI of include/msda_b200.h.
//
// Written from scratch for Blackwell; it is NOT a port of the reference's mmcv-derived kernels
// (codetr/csrc/ms_deform_attn.cu: forward :211-261, one thread per output channel with scalar loads;
// backward :263-760).  What it computes is the reference's operator, in the reference's layouts and with
// its boundary semantics:
//
//   out[b,q,m,:] = sum_{l,p} w[b,q,m,l,p] * bilinear(value_l[b,:,m,:], loc[b,q,m,l,p])
//
// with  x_pix = x*W - 0.5,  y_pix = y*H - 0.5  (align_corners=False), zero padding per corner and the
// whole-sample range test of ms_deform_attn.cu:246-249.
//
// Kernels (DESIGN.md section 4 has the measurements behind every choice, including the ones that lost)
//   msda_fwd_vec      the hot kernel.  A lane group of G = D*sizeof(T)/16 lanes owns one (query, head)
//                     pair and fetches each 64-byte corner row with one LDG.E.128 per lane; lane k of the
//                     group works out the geometry of point k of the level once and broadcasts it by
//                     shuffle; every "does not contribute" case is a weight of exactly zero, which
//                     predicates off the load and the FMAs; fp16 multiplies in place with Blackwell's
//                     mixed-precision FMA (PTX fma.rn.f32.f16 -> SASS FHFMA) and accumulates in fp32, the
//                     exact paths (bf16, fp32) use the packed fp32 FMA (fma.rn.f32x2 -> FFMA2); the
//                     grid is persistent (resident CTA count) and the next unit's first sample is loaded
//                     while the current one is computed.  Template switches: P (4 or run-time), SPLIT
//                     (points dealt to 2/4 lane groups), MATH (fhfma / exact), STAGE (TMA bulk-copy staging
//                     of locations and weights, opt-in), FUSED (softmax + location arithmetic in-kernel), DYN
//                     (warps draw units from a device counter, opt-in).
//   msda_fwd_small    decoder-sized launches: 4-way point split, 128-thread CTAs, level table held in
//                     lanes, no shared memory, no barrier.
//   msda_pack_value + msda_fwd_packed   opt-in packed-pyramid path: 128-byte (pixel, head) entries that also
//                     hold the right-hand neighbour, fetched with 256-bit loads (LDG.E.256, sm_100+).
//   msda_fwd_generic  any shape / any dtype (incl. fp64), one thread per output element.
//   msda_bwd_vec / msda_bwd_generic     backward; grad_value scattered with 16-byte vector reductions
//                     (red.global.add.v4.f32 / .noftz.v4.f16x2 -> SASS REDG.E.ADD.F16x8).
//   read_probe_kernel L2 / HBM read-bandwidth probe for the roofline denominators.
//
// Level shapes / start indices stay device-resident (TensorRT hands them over as device buffers,
// deformable_attention_plugin.cpp:339-341): the kernels read them, the launcher never does, so every entry
// point is capture-safe and sync-free.  Tensor cores are not used: the op is a gather plus a weighted sum.

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "msda_b200.h"
#include "msda_internal.hpp"

namespace msda_detail __attribute__((visibility("hidden"))) {
std::atomic<uint64_t> launch_count{0};
}  // namespace msda_detail

namespace {

constexpr int kMaxLevelsSmem = 16;  // levels cached in shared memory by the fast kernels
constexpr int kThreads = 256;
#ifndef MSDA_MINB
#define MSDA_MINB 4
#endif
#ifndef MSDA_NB
#define MSDA_NB 1
#endif
#ifndef MSDA_PIPE
#define MSDA_PIPE 0
#endif
#ifndef MSDA_FFMA2
#define MSDA_FFMA2 1  // packed fp32 FMA (FFMA2) on the exact-arithmetic paths
#endif
#ifndef MSDA_BF16_UNPACK
#define MSDA_BF16_UNPACK 2  // low bf16 -> fp32: 0 = compiler's shift (IMAD.U32, FMA pipe), 1 = PRMT (ALU pipe), 2 = alternate
                            // the two (measured 60.4 / 59.5 / 59.0 us at the headline shape for 1 / 0 / 2)
#endif
#ifndef MSDA_LB
#define MSDA_LB 1   // levels whose row loads are issued together in the split-points path (measured: 1 is best,
                    // more registers per thread push the 450-CTA decoder grid into a second wave)
#endif
#ifndef MSDA_MINB_SPLIT
#define MSDA_MINB_SPLIT 4
#endif

std::atomic<uint64_t> &g_launch_count = msda_detail::launch_count;
thread_local char g_last_variant[128] = "none";

// ---------------------------------------------------------------------------
// element traits
// ---------------------------------------------------------------------------
template <typename T>
struct Elem;

template <>
struct Elem<float> {
  using acc_t = float;
  static __device__ __forceinline__ float to_acc(float v) { return v; }
  static __device__ __forceinline__ float from_acc(float v) { return v; }
};
template <>
struct Elem<double> {
  using acc_t = double;
  static __device__ __forceinline__ double to_acc(double v) { return v; }
  static __device__ __forceinline__ double from_acc(double v) { return v; }
};
template <>
struct Elem<__half> {
  using acc_t = float;
  static __device__ __forceinline__ float to_acc(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half from_acc(float v) { return __float2half_rn(v); }
};
template <>
struct Elem<__nv_bfloat16> {
  using acc_t = float;
  static __device__ __forceinline__ float to_acc(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_acc(float v) { return __float2bfloat16_rn(v); }
};

// ---------------------------------------------------------------------------
// Dynamic unit scheduling.  Each launch that uses it owns one {next, done} counter pair out of this pool
// (picked round-robin on the host).  Warps draw (tile, pass, warp-slice) units with atomicAdd(next); the last
// warp of the grid to run dry resets the pair, so the pool needs no host-side memset and stays zero between
// launches.  Not used under stream capture (a replayed graph would share its slot with itself).
// ---------------------------------------------------------------------------
constexpr int kSchedSlots = 4096;
__device__ unsigned g_sched[kSchedSlots][2];

// ---------------------------------------------------------------------------
// kernel parameters
// ---------------------------------------------------------------------------
struct MsdaParams {
  const void *value;
  const int64_t *shapes;  // [L,2] (H,W), device
  const int64_t *starts;  // [L], device
  const void *loc;        // [B,Q,M,L,P,2]            (plain mode)
  const void *weight;     // [B,Q,M,L,P]              (plain mode)
  const void *ref;        // [B,Q,L,ref_dim]          (fused mode)
  const void *offsets;    // [B,Q,M,L,P,2]            (fused mode)
  const void *logits;     // [B,Q,M,L*P]              (fused mode)
  void *out;              // [B,Q,M*D]
  void *packed;           // workspace: pixel-pair packed value pyramid (packed path), else nullptr
  int B, S, M, D, L, Q, P;
  int ref_dim;  // 0 = plain mode, 2 or 4 = fused mode
  int tile_w_log2, tile_h_log2;  // query tile is 2^tile_w_log2 x 2^tile_h_log2 (tiled) or that many consecutive queries (linear)
  int passes;           // CTA passes per tile = ceil(tile queries * M / pairs per pass)
  float inv_passes, inv_M;       // reciprocals for fast_div
  int qpp;              // queries per pass when PAIRS_PER_PASS is a whole number of queries, else 0
  int stage_loc_row, stage_w_row;  // padded shared-memory row pitch (bytes) of one query's locations / weights
  int want_tiled;       // 1: use 2-D tiles when sum(H*W) == Q
  int head_major;       // 1: a warp holds one head of 32/G neighbouring queries
  int chunked;          // 1: each CTA owns a contiguous run of (tile, pass) units instead of a strided set
  int pdl_early_tables; // 1: (MSDA_FLAG_PDL) read the level tables before waiting for the preceding kernel
  int l2_prefetch;      // 1: every CTA starts by asking L2 to fetch its share of the image's value tensor
  unsigned *sched;      // dynamic unit scheduling: {next warp-unit, finished warps} counters of this launch, or nullptr
  int hp_smem_bytes;    // head-pair kernel: dynamic shared memory available for cached pyramid levels
};

struct LevelGeom {
  int H, W, start, qstart;
  float rcpW, rcpH;  // correctly rounded 1/W, 1/H for the fused producers' offset normalisation
};

// ---------------------------------------------------------------------------
// sample geometry, shared by every kernel
//   follows ms_deform_attn.cu:246-249 (un-normalise, whole-sample test) and
//   :35-42, :53-73 (floor, corner validity, corner weights)
// ---------------------------------------------------------------------------
template <typename A>
struct Sample {
  int idx[4];   // pixel index (h*W + w) of the four corners inside the level
  A cw[4];      // corner weight, already multiplied by the attention weight
  bool ok[4];   // corner inside the level (false -> contributes zero)
};

template <typename A>
__device__ __forceinline__ A floor_acc(A v);
template <>
__device__ __forceinline__ float floor_acc<float>(float v) { return floorf(v); }
template <>
__device__ __forceinline__ double floor_acc<double>(double v) { return floor(v); }

// product rounded on its own (never contracted into an FMA with the following subtraction), like the
// reference's scalar_t arithmetic `loc * size - 0.5`: at exact-integer pixel coordinates the choice of the
// bilinear cell -- and with it the one-sided derivative the backward returns -- depends on this rounding
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }

template <typename A>
__device__ __forceinline__ Sample<A> make_sample(A x, A y, A aw, int H, int W) {
  Sample<A> s;
  const A w_im = mul_rn(x, (A)W) - (A)0.5;
  const A h_im = mul_rn(y, (A)H) - (A)0.5;
  const bool inside = (h_im > (A)-1) && (w_im > (A)-1) && (h_im < (A)H) && (w_im < (A)W);
  const A hf = floor_acc<A>(h_im), wf = floor_acc<A>(w_im);
  const int h_lo = inside ? (int)hf : 0;
  const int w_lo = inside ? (int)wf : 0;
  const A lh = h_im - hf, lw = w_im - wf;
  const A hh = (A)1 - lh, hw = (A)1 - lw;
  const bool top = h_lo >= 0, bot = h_lo + 1 <= H - 1;
  const bool lef = w_lo >= 0, rig = w_lo + 1 <= W - 1;
  s.ok[0] = inside && top && lef;
  s.ok[1] = inside && top && rig;
  s.ok[2] = inside && bot && lef;
  s.ok[3] = inside && bot && rig;
  const int base = h_lo * W + w_lo;
  s.idx[0] = base;
  s.idx[1] = base + 1;
  s.idx[2] = base + W;
  s.idx[3] = base + W + 1;
  // a sample that fails the range test contributes nothing; its weights may be inf/NaN
  // (non-finite or huge locations), so they are forced to zero rather than multiplied by zero rows
  const A awi = inside ? aw : (A)0;
  s.cw[0] = inside ? (hh * hw) * awi : (A)0;
  s.cw[1] = inside ? (hh * lw) * awi : (A)0;
  s.cw[2] = inside ? (lh * hw) * awi : (A)0;
  s.cw[3] = inside ? (lh * lw) * awi : (A)0;
  return s;
}

// ---------------------------------------------------------------------------
// Fused-producer mode (msda_b200_forward_fused): the softmax over L*P and the sampling-location
// arithmetic of the calling module (codetr/multi_scale_deformable_attention.py:180-200) happen in the
// kernel.  The module runs those steps as separate PyTorch ops in the tensor dtype, so for fp16/bf16 every
// intermediate (off / normaliser, ref + ..., off / P, ... * wh, softmax output) is rounded to 16 bits; the
// helpers below round at the same places, so the fused result tracks the unfused pipeline instead of
// being "more exact" than it.  For float / double the roundings are identities.
// ---------------------------------------------------------------------------
template <typename T, typename A>
__device__ __forceinline__ A round_like(A v) {
  return (A)Elem<T>::to_acc(Elem<T>::from_acc((typename Elem<T>::acc_t)v));
}

// exp of the softmax: full-precision expf for fp32 tensors, the fast intrinsic when the result is rounded to 16 bits
template <typename T>
__device__ __forceinline__ float fused_exp(float v) {
  if constexpr (sizeof(T) == 4) return expf(v);
  else return __expf(v);
}

// a / b for fp32 with y = RN(1/b) given: one multiply and two FMAs that land on the correctly rounded quotient
// (Markstein: q = RN(a*y), r = a - b*q exactly by FMA, RN(q + r*y) = RN(a/b)); replaces the ~10-instruction IEEE
// division sequence in the hot loop of the fused kernel without changing a bit of the result.
__device__ __forceinline__ float div_with_rcp(float a, float b, float rcp_b) {
  const float q = __fmul_rn(a, rcp_b);
  const float r = fmaf(-b, q, a);
  return fmaf(r, rcp_b, q);
}

// the vector kernel's variant of fused_location below: fp32 arithmetic, reciprocals of W, H and P precomputed
template <typename T>
__device__ __forceinline__ void fused_location_fast(const T *rf, int ref_dim, float ox, float oy, float Wf, float Hf, float rcpW,
                                                    float rcpH, float &x, float &y) {
  if (ref_dim == 2) {
    x = round_like<T, float>(Elem<T>::to_acc(rf[0]) + round_like<T, float>(div_with_rcp(ox, Wf, rcpW)));
    y = round_like<T, float>(Elem<T>::to_acc(rf[1]) + round_like<T, float>(div_with_rcp(oy, Hf, rcpH)));
  } else {
    // P == 4 on this path: off / 4 is exact
    const float tx = round_like<T, float>(round_like<T, float>(ox * 0.25f) * Elem<T>::to_acc(rf[2])) * 0.5f;
    const float ty = round_like<T, float>(round_like<T, float>(oy * 0.25f) * Elem<T>::to_acc(rf[3])) * 0.5f;
    x = round_like<T, float>(Elem<T>::to_acc(rf[0]) + tx);
    y = round_like<T, float>(Elem<T>::to_acc(rf[1]) + ty);
  }
}

template <typename T, typename A>
__device__ __forceinline__ void fused_location(const T *rf, int ref_dim, A ox, A oy, int H, int W, int P, A &x, A &y) {
  if (ref_dim == 2) {
    // reference_points + sampling_offsets / (W, H)        (:186-191)
    x = round_like<T, A>((A)Elem<T>::to_acc(rf[0]) + round_like<T, A>(ox / (A)W));
    y = round_like<T, A>((A)Elem<T>::to_acc(rf[1]) + round_like<T, A>(oy / (A)H));
  } else {
    // reference_points[..., :2] + sampling_offsets / P * reference_points[..., 2:] * 0.5   (:192-196)
    const A tx = round_like<T, A>(round_like<T, A>(ox / (A)P) * (A)Elem<T>::to_acc(rf[2])) * (A)0.5;
    const A ty = round_like<T, A>(round_like<T, A>(oy / (A)P) * (A)Elem<T>::to_acc(rf[3])) * (A)0.5;
    x = round_like<T, A>((A)Elem<T>::to_acc(rf[0]) + tx);
    y = round_like<T, A>((A)Elem<T>::to_acc(rf[1]) + ty);
  }
}

// ---------------------------------------------------------------------------
// Generic kernel: any D, any L, any P, any dtype (incl. double).  One thread per
// output element; used for shapes the vector kernel does not cover and as the
// in-library cross-check of the fast paths.  Fused mode is supported here too.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) msda_fwd_generic(const MsdaParams p) {
  using A = typename Elem<T>::acc_t;
  const T *__restrict__ value = static_cast<const T *>(p.value);
  T *__restrict__ out = static_cast<T *>(p.out);
  const int64_t n = (int64_t)p.B * p.Q * p.M * p.D;
  const int LP = p.L * p.P;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % p.D);
    const int64_t pair = i / p.D;  // (b*Q + q)*M + m
    const int m = (int)(pair % p.M);
    const int64_t bq = pair / p.M;
    const int64_t b = bq / p.Q;
    const T *vb = value + ((int64_t)b * p.S) * p.M * p.D + (int64_t)m * p.D + c;

    A mx = 0, denom = 1;
    if (p.ref_dim) {  // softmax statistics over the L*P logits of this pair
      const T *lg = static_cast<const T *>(p.logits) + pair * LP;
      mx = Elem<T>::to_acc(lg[0]);
      for (int j = 1; j < LP; ++j) {
        const A v = Elem<T>::to_acc(lg[j]);
        mx = v > mx ? v : mx;
      }
      denom = 0;
      for (int j = 0; j < LP; ++j) denom += exp(Elem<T>::to_acc(lg[j]) - mx);
    }

    A acc = 0;
    for (int l = 0; l < p.L; ++l) {
      const int H = (int)p.shapes[2 * l], W = (int)p.shapes[2 * l + 1];
      const T *vl = vb + p.starts[l] * (int64_t)p.M * p.D;
      for (int k = 0; k < p.P; ++k) {
        A x, y, aw;
        const int64_t si = pair * LP + (int64_t)l * p.P + k;
        if (p.ref_dim == 0) {
          const T *lc = static_cast<const T *>(p.loc) + si * 2;
          x = Elem<T>::to_acc(lc[0]);
          y = Elem<T>::to_acc(lc[1]);
          aw = Elem<T>::to_acc(static_cast<const T *>(p.weight)[si]);
        } else {
          const T *of = static_cast<const T *>(p.offsets) + si * 2;
          const T *rf = static_cast<const T *>(p.ref) + (bq * p.L + l) * p.ref_dim;
          const A ox = Elem<T>::to_acc(of[0]), oy = Elem<T>::to_acc(of[1]);
          fused_location<T, A>(rf, p.ref_dim, ox, oy, H, W, p.P, x, y);
          aw = round_like<T, A>(exp((A)Elem<T>::to_acc(static_cast<const T *>(p.logits)[si]) - mx) / denom);
        }
        const Sample<A> s = make_sample<A>(x, y, aw, H, W);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (s.ok[j]) acc += s.cw[j] * Elem<T>::to_acc(vl[(int64_t)s.idx[j] * p.M * p.D]);
        }
      }
    }
    out[i] = Elem<T>::from_acc(acc);
  }
}

// ---------------------------------------------------------------------------
// Vector kernel helpers
// ---------------------------------------------------------------------------
enum MathMode { kExact = 0, kFhfma = 1 };

// Cache policy knobs (compile-time, for the tuning builds): MSDA_VALUE_HINT 0 = plain read-only load,
// 1 = L1::evict_last (value rows are the reused data); MSDA_STREAM_HINT 0 = plain, 1 = L1::no_allocate for the
// streamed-once inputs (locations, weights) so they do not displace value rows from L1.
#ifndef MSDA_VALUE_HINT
#define MSDA_VALUE_HINT 0
#endif
#ifndef MSDA_STREAM_HINT
#define MSDA_STREAM_HINT 0
#endif

__device__ __forceinline__ uint4 ldg128(const void *ptr) {
#if MSDA_VALUE_HINT == 1
  uint4 r;
  asm volatile("ld.global.nc.L1::evict_last.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(ptr));
  return r;
#else
  return __ldg(static_cast<const uint4 *>(ptr));
#endif
}

__device__ __forceinline__ unsigned ld_stream_u32(const void *ptr) {
#if MSDA_STREAM_HINT == 1
  unsigned r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(ptr));
  return r;
#else
  return __ldg(static_cast<const unsigned *>(ptr));
#endif
}
__device__ __forceinline__ unsigned short ld_stream_u16(const void *ptr) {
#if MSDA_STREAM_HINT == 1
  unsigned short r;
  asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(r) : "l"(ptr));
  return r;
#else
  return __ldg(static_cast<const unsigned short *>(ptr));
#endif
}

// acc[0..VEC) += cw * row, for one 16-byte piece of a corner row
template <typename T, int MATH>
struct RowFma;

// Blackwell packed fp32 FMA: (a0, a1) += (x0, x1) * w in one instruction (SASS FFMA2 with the weight as a scalar
// broadcast operand).  Each half is an ordinary round-to-nearest fmaf, so results are bit-identical to two
// scalar FMAs; the point is one issue slot instead of two on the fp32 and bf16 paths.
__device__ __forceinline__ void fma2(float &a0, float &a1, float x0, float x1, float w) {
#if MSDA_FFMA2
  unsigned long long acc, x, ww;
  asm("mov.b64 %0, {%1, %2};" : "=l"(acc) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(x0), "f"(x1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(ww) : "f"(w));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(x), "l"(ww));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(acc));
#else
  a0 = fmaf(w, x0, a0);
  a1 = fmaf(w, x1, a1);
#endif
}

template <int MATH>
struct RowFma<float, MATH> {
  static __device__ __forceinline__ void run(float (&acc)[4], const uint4 &r, float cw, unsigned /*cw16*/) {
    fma2(acc[0], acc[1], __uint_as_float(r.x), __uint_as_float(r.y), cw);
    fma2(acc[2], acc[3], __uint_as_float(r.z), __uint_as_float(r.w), cw);
  }
};

template <>
struct RowFma<__half, kExact> {
  static __device__ __forceinline__ void run(float (&acc)[8], const uint4 &r, float cw, unsigned) {
    const unsigned w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      // scalar FMAs here: with the HADD2.F32 unpacks on the same pipe FFMA2 measured slower (66.1 vs 62.7 us)
      const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&w[i]));
      acc[2 * i] = fmaf(cw, f.x, acc[2 * i]);
      acc[2 * i + 1] = fmaf(cw, f.y, acc[2 * i + 1]);
    }
  }
};

template <>
struct RowFma<__nv_bfloat16, kExact> {
  static __device__ __forceinline__ void run(float (&acc)[8], const uint4 &r, float cw, unsigned) {
    const unsigned w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      // bf16 -> fp32 is a 16-bit shift: keep it on the integer pipe
#if MSDA_BF16_UNPACK == 0
      const unsigned lo = w[i] << 16;
#else
      unsigned lo;
      if (MSDA_BF16_UNPACK == 1 || (i & 1)) asm("prmt.b32 %0, %1, 0, 0x1044;" : "=r"(lo) : "r"(w[i]));
      else lo = w[i] << 16;
#endif
      fma2(acc[2 * i], acc[2 * i + 1], __uint_as_float(lo), __uint_as_float(w[i] & 0xffff0000u), cw);
    }
  }
};

// Blackwell mixed-precision FMA: d(f32) = a(f16) * b(f16) + c(f32), SASS FHFMA.
__device__ __forceinline__ float fhfma_f16(unsigned short a, unsigned short b, float c) {
  float d;
  asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(d) : "h"(a), "h"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float fhfma_bf16(unsigned short a, unsigned short b, float c) {
  float d;
  asm("fma.rn.f32.bf16 %0, %1, %2, %3;" : "=f"(d) : "h"(a), "h"(b), "f"(c));
  return d;
}

template <>
struct RowFma<__half, kFhfma> {
  static __device__ __forceinline__ void run(float (&acc)[8], const uint4 &r, float, unsigned cw16) {
    const unsigned w[4] = {r.x, r.y, r.z, r.w};
    const unsigned short c = (unsigned short)cw16;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc[2 * i] = fhfma_f16((unsigned short)(w[i] & 0xffffu), c, acc[2 * i]);
      acc[2 * i + 1] = fhfma_f16((unsigned short)(w[i] >> 16), c, acc[2 * i + 1]);
    }
  }
};

template <>
struct RowFma<__nv_bfloat16, kFhfma> {
  static __device__ __forceinline__ void run(float (&acc)[8], const uint4 &r, float, unsigned cw16) {
    const unsigned w[4] = {r.x, r.y, r.z, r.w};
    const unsigned short c = (unsigned short)cw16;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc[2 * i] = fhfma_bf16((unsigned short)(w[i] & 0xffffu), c, acc[2 * i]);
      acc[2 * i + 1] = fhfma_bf16((unsigned short)(w[i] >> 16), c, acc[2 * i + 1]);
    }
  }
};

template <typename T>
__device__ __forceinline__ unsigned weight_to_16(float cw);
template <>
__device__ __forceinline__ unsigned weight_to_16<__half>(float cw) {
  return (unsigned)__half_as_ushort(__float2half_rn(cw));
}
template <>
__device__ __forceinline__ unsigned weight_to_16<__nv_bfloat16>(float cw) {
  return (unsigned)__bfloat16_as_ushort(__float2bfloat16_rn(cw));
}
template <>
__device__ __forceinline__ unsigned weight_to_16<float>(float) {
  return 0u;
}

// two packed 16-bit elements -> two floats
template <typename T>
__device__ __forceinline__ float2 unpack2(unsigned v);
template <>
__device__ __forceinline__ float2 unpack2<__half>(unsigned v) {
  return __half22float2(*reinterpret_cast<const __half2 *>(&v));
}
template <>
__device__ __forceinline__ float2 unpack2<__nv_bfloat16>(unsigned v) {
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}
template <>
__device__ __forceinline__ float2 unpack2<float>(unsigned) {
  return make_float2(0.f, 0.f);
}

template <typename T, int VEC>
__device__ __forceinline__ void store_row(T *dst, const float (&acc)[VEC]);

template <>
__device__ __forceinline__ void store_row<float, 4>(float *dst, const float (&acc)[4]) {
  *reinterpret_cast<float4 *>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
}
template <>
__device__ __forceinline__ void store_row<__half, 8>(__half *dst, const float (&acc)[8]) {
  uint4 o;
  __half2 h;
  h = __floats2half2_rn(acc[0], acc[1]);
  o.x = *reinterpret_cast<unsigned *>(&h);
  h = __floats2half2_rn(acc[2], acc[3]);
  o.y = *reinterpret_cast<unsigned *>(&h);
  h = __floats2half2_rn(acc[4], acc[5]);
  o.z = *reinterpret_cast<unsigned *>(&h);
  h = __floats2half2_rn(acc[6], acc[7]);
  o.w = *reinterpret_cast<unsigned *>(&h);
  *reinterpret_cast<uint4 *>(dst) = o;
}
template <>
__device__ __forceinline__ void store_row<__nv_bfloat16, 8>(__nv_bfloat16 *dst, const float (&acc)[8]) {
  uint4 o;
  __nv_bfloat162 h;
  h = __floats2bfloat162_rn(acc[0], acc[1]);
  o.x = *reinterpret_cast<unsigned *>(&h);
  h = __floats2bfloat162_rn(acc[2], acc[3]);
  o.y = *reinterpret_cast<unsigned *>(&h);
  h = __floats2bfloat162_rn(acc[4], acc[5]);
  o.z = *reinterpret_cast<unsigned *>(&h);
  h = __floats2bfloat162_rn(acc[6], acc[7]);
  o.w = *reinterpret_cast<unsigned *>(&h);
  *reinterpret_cast<uint4 *>(dst) = o;
}

// Lean sample geometry for the vector kernels.  Follows ms_deform_attn.cu:246-249 and :35-73 like
// make_sample(), but folds every "does not contribute" case into a zero weight: a corner outside the
// level, or a sample that fails the whole-sample range test (including NaN / infinite locations, for
// which every comparison is false).  Selects, not multiplications, produce the zeros, so non-finite
// intermediates never leak.  i00 is the pixel index of the top-left corner inside the level; it is only
// meaningful for corners whose weight is non-zero.
__device__ __forceinline__ void make_geo(float x, float y, float aw, int H, int W, int &i00, float (&cw)[4]) {
  const float w_im = __fmul_rn(x, (float)W) - 0.5f;  // product rounded first, like the reference
  const float h_im = __fmul_rn(y, (float)H) - 0.5f;
  const bool inside = (h_im > -1.f) && (w_im > -1.f) && (h_im < (float)H) && (w_im < (float)W);
  const float hf = floorf(h_im), wf = floorf(w_im);
  const int h_lo = (int)hf, w_lo = (int)wf;
  const float lh = h_im - hf, lw = w_im - wf;
  const float hh = 1.f - lh, hw = 1.f - lw;
  const float wy0 = (inside && h_lo >= 0) ? hh * aw : 0.f;
  const float wy1 = (inside && h_lo < H - 1) ? lh * aw : 0.f;
  const float wx0 = (inside && w_lo >= 0) ? hw : 0.f;
  const float wx1 = (inside && w_lo < W - 1) ? lw : 0.f;
  cw[0] = wy0 * wx0;
  cw[1] = wy0 * wx1;
  cw[2] = wy1 * wx0;
  cw[3] = wy1 * wx1;
  i00 = h_lo * W + w_lo;
}

// location (x, y) and attention weight of sample `si` of a (query, head) pair, as floats
template <typename T>
__device__ __forceinline__ void load_sample_inputs(const T *lp, const T *wp, int si, float &x, float &y, float &aw) {
  if constexpr (sizeof(T) == 2) {
    const float2 xy = unpack2<T>(ld_stream_u32(reinterpret_cast<const unsigned *>(lp) + si));
    x = xy.x;
    y = xy.y;
    const unsigned short wraw = ld_stream_u16(reinterpret_cast<const unsigned short *>(wp) + si);
    aw = unpack2<T>((unsigned)wraw).x;
  } else {
    const float2 xy = __ldg(reinterpret_cast<const float2 *>(lp) + si);
    x = xy.x;
    y = xy.y;
    aw = __ldg(reinterpret_cast<const float *>(wp) + si);
  }
}

// same, from the shared-memory staging rows of the pair (generic loads: LDS)
template <typename T>
__device__ __forceinline__ void load_sample_inputs_smem(const unsigned char *lp, const unsigned char *wp, int si, float &x,
                                                        float &y, float &aw) {
  if constexpr (sizeof(T) == 2) {
    const float2 xy = unpack2<T>(reinterpret_cast<const unsigned *>(lp)[si]);
    x = xy.x;
    y = xy.y;
    aw = unpack2<T>((unsigned)reinterpret_cast<const unsigned short *>(wp)[si]).x;
  } else {
    const float2 xy = reinterpret_cast<const float2 *>(lp)[si];
    x = xy.x;
    y = xy.y;
    aw = reinterpret_cast<const float *>(wp)[si];
  }
}

// Undecoded location + weight of one sample: loaded early (next level / next pass), converted at use, so
// the conversion never becomes the point where the warp waits for the load.
struct RawSample {
  unsigned a, b, w;  // 16-bit types: a = packed (x, y), w = weight bits; fp32: a = x, b = y, w = weight
};
template <typename T>
__device__ __forceinline__ RawSample load_raw(const T *lp, const T *wp, int si) {
  RawSample r;
  if constexpr (sizeof(T) == 2) {
    r.a = ld_stream_u32(reinterpret_cast<const unsigned *>(lp) + si);
    r.b = 0u;
    r.w = (unsigned)ld_stream_u16(reinterpret_cast<const unsigned short *>(wp) + si);
  } else {
    const float2 xy = __ldg(reinterpret_cast<const float2 *>(lp) + si);
    r.a = __float_as_uint(xy.x);
    r.b = __float_as_uint(xy.y);
    r.w = __float_as_uint(__ldg(reinterpret_cast<const float *>(wp) + si));
  }
  return r;
}
template <typename T>
__device__ __forceinline__ RawSample load_raw_smem(const unsigned char *lp, const unsigned char *wp, int si) {
  RawSample r;
  if constexpr (sizeof(T) == 2) {
    r.a = reinterpret_cast<const unsigned *>(lp)[si];
    r.b = 0u;
    r.w = (unsigned)reinterpret_cast<const unsigned short *>(wp)[si];
  } else {
    const float2 xy = reinterpret_cast<const float2 *>(lp)[si];
    r.a = __float_as_uint(xy.x);
    r.b = __float_as_uint(xy.y);
    r.w = __float_as_uint(reinterpret_cast<const float *>(wp)[si]);
  }
  return r;
}
template <typename T>
__device__ __forceinline__ void decode_raw(const RawSample &r, float &x, float &y, float &aw) {
  if constexpr (sizeof(T) == 2) {
    const float2 xy = unpack2<T>(r.a);
    x = xy.x;
    y = xy.y;
    aw = unpack2<T>(r.w).x;
  } else {
    x = __uint_as_float(r.a);
    y = __uint_as_float(r.b);
    aw = __uint_as_float(r.w);
  }
}

// two fp32 weights -> packed 16-bit pair in the element type (low half = first)
template <typename T>
__device__ __forceinline__ unsigned pack_weights(float a, float b);
template <>
__device__ __forceinline__ unsigned pack_weights<__half>(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const unsigned *>(&h);
}
template <>
__device__ __forceinline__ unsigned pack_weights<__nv_bfloat16>(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const unsigned *>(&h);
}
template <>
__device__ __forceinline__ unsigned pack_weights<float>(float, float) {
  return 0u;
}

// ---- programmatic dependent launch (PDL): let the next kernel of the stream start its prologue while this
// one drains, and hold this kernel's data reads until the previous kernel has completed.  Both are no-ops
// unless the launch carried cudaLaunchAttributeProgrammaticStreamSerialization.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait_prior_grid() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// One bulk L2 prefetch per CTA (cp.async.bulk.prefetch.L2): the CTAs of image blockIdx.y cover its value tensor, so
// the whole pyramid streams into L2 at HBM rate while the first units compute.  Without it every first touch of a
// value row is a DRAM-latency miss, and a warp waits for the slowest of the 128 row loads of a sample.  Worth
// 0.3 us at the headline shape (which runs within 0.5 us of its all-L2-warm time either way,
// tests/perf_cold_parts.py) and 5-7 % on the small shapes.  A no-op when the producer kernel left `value` in L2.
__device__ __forceinline__ void prefetch_value_l2(const MsdaParams &p, int elem_bytes) {
  if (!p.l2_prefetch || threadIdx.x != 64) return;  // a thread with no set-up work
  const size_t bytes = (size_t)p.S * p.M * p.D * elem_bytes;
  size_t per = (bytes + gridDim.x - 1) / gridDim.x;
  per = (per + 127) & ~(size_t)127;
  const size_t off = (size_t)blockIdx.x * per;
  if (off >= bytes) return;
  size_t n = bytes - off < per ? bytes - off : per;
  n &= ~(size_t)15;
  const char *src = static_cast<const char *>(p.value) + (size_t)blockIdx.y * bytes + off;
  if (n) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"((unsigned)n) : "memory");
}

// ---- TMA 1-D bulk copy + mbarrier (Blackwell/Hopper async proxy) ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------------------
// Tile bookkeeping shared by the vector kernels
// ---------------------------------------------------------------------------
// Exact n / d for 0 <= n < 2^22 with a precomputed float reciprocal: the estimate is off by at most
// one, which the remainder test repairs.  Replaces ~20-instruction integer divisions by run-time
// divisors in the per-pass decode.
__device__ __forceinline__ int fast_div(int n, int d, float inv_d, int &rem) {
  int q = (int)((float)n * inv_d);
  int r = n - q * d;
  if (r < 0) {
    --q;
    r += d;
  } else if (r >= d) {
    ++q;
    r -= d;
  }
  rem = r;
  return q;
}

struct TileSetup {
  LevelGeom lv[kMaxLevelsSmem];
  int tile_first[kMaxLevelsSmem + 1];  // first tile index of each level (tiled order)
  int tiles_x[kMaxLevelsSmem];         // tiles per row of each level
  float inv_tiles_x[kMaxLevelsSmem];
  int n_tiles;                         // tiles per image
  int tiled;                           // 1 = 2-D tiles, 0 = linear chunks
  int layout_ok;                       // 1 = levels are disjoint key ranges inside [0, S) (packed path)
};

__device__ __forceinline__ void setup_tiles(const MsdaParams &p, TileSetup &ts) {
  // executed by warp 0 of the CTA (all 32 lanes); L <= kMaxLevelsSmem <= 32 guaranteed by the host.
  // Lane l loads level l (the loads of all levels are in flight together), then two warp scans
  // give each level its first query and first tile.  Tile extents are powers of two (shifts).
  const int l = threadIdx.x;
  int H = 0, W = 0, start = 0;
  if (l < p.L) {
    H = (int)__ldg(p.shapes + 2 * l);
    W = (int)__ldg(p.shapes + 2 * l + 1);
    start = (int)__ldg(p.starts + l);
  }
  const int nq = H * W;
  const int tx = (W + (1 << p.tile_w_log2) - 1) >> p.tile_w_log2;
  const int ty = (H + (1 << p.tile_h_log2) - 1) >> p.tile_h_log2;
  const int nt = (l < p.L) ? tx * ty : 0;
  int qs = nq, tsum = nt;  // inclusive scans
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int a = __shfl_up_sync(0xffffffffu, qs, off);
    const int c = __shfl_up_sync(0xffffffffu, tsum, off);
    if (l >= off) {
      qs += a;
      tsum += c;
    }
  }
  if (l < p.L) {
    ts.lv[l].H = H;
    ts.lv[l].W = W;
    ts.lv[l].start = start;
    ts.lv[l].qstart = qs - nq;
    ts.lv[l].rcpW = __frcp_rn((float)(W > 0 ? W : 1));
    ts.lv[l].rcpH = __frcp_rn((float)(H > 0 ? H : 1));
    ts.tile_first[l] = tsum - nt;
    ts.tiles_x[l] = tx > 0 ? tx : 1;
    ts.inv_tiles_x[l] = 1.0f / (float)(tx > 0 ? tx : 1);
  }
  const int q_total = __shfl_sync(0xffffffffu, qs, 31);
  const int t_total = __shfl_sync(0xffffffffu, tsum, 31);
  // standard pyramid layout? every level a key range inside [0, S), pairwise disjoint
  bool bad = (l < p.L) && (start < 0 || H < 0 || W < 0 || (long long)start + (long long)nq > (long long)p.S);
  for (int o = 0; o < p.L; ++o) {
    const int os = __shfl_sync(0xffffffffu, start, o), on = __shfl_sync(0xffffffffu, nq, o);
    if (l < p.L && o != l && nq > 0 && on > 0 && start < os + on && os < start + nq) bad = true;
  }
  const unsigned any_bad = __ballot_sync(0xffffffffu, bad);
  if (l == 0) {
    ts.layout_ok = any_bad == 0u;
    ts.tile_first[p.L] = t_total;
    const int tq_log2 = p.tile_w_log2 + p.tile_h_log2;
    if (p.want_tiled && q_total == p.Q) {
      ts.tiled = 1;
      ts.n_tiles = t_total;
    } else {
      ts.tiled = 0;
      ts.n_tiles = (p.Q + (1 << tq_log2) - 1) >> tq_log2;
    }
  }
}

// Maps (tile t, query slot tq inside the tile) to a query index, or -1 for a padding slot.
__device__ __forceinline__ int tile_query(const MsdaParams &p, const TileSetup &ts, int t, int tq) {
  if (!ts.tiled) {
    const int q = (t << (p.tile_w_log2 + p.tile_h_log2)) + tq;
    return q < p.Q ? q : -1;
  }
  int l = 0;
  while (l + 1 < p.L && t >= ts.tile_first[l + 1]) ++l;
  int tcol;
  const int trow = fast_div(t - ts.tile_first[l], ts.tiles_x[l], ts.inv_tiles_x[l], tcol);
  const int y = (trow << p.tile_h_log2) + (tq >> p.tile_w_log2);
  const int x = (tcol << p.tile_w_log2) + (tq & ((1 << p.tile_w_log2) - 1));
  return (y < ts.lv[l].H && x < ts.lv[l].W) ? ts.lv[l].qstart + y * ts.lv[l].W + x : -1;
}

// First query and number of valid queries among the `qpp` consecutive tile slots [tq0, tq0 + qpp) of
// tile t.  Requires qpp to divide the tile width, so the slots lie in one tile row and the queries are
// consecutive in memory.
__device__ __forceinline__ void pass_query_range(const MsdaParams &p, const TileSetup &ts, int t, int tq0, int qpp, int &q0,
                                                 int &nq) {
  if (!ts.tiled) {
    q0 = (t << (p.tile_w_log2 + p.tile_h_log2)) + tq0;
    nq = p.Q - q0;
  } else {
    int l = 0;
    while (l + 1 < p.L && t >= ts.tile_first[l + 1]) ++l;
    int tcol;
    const int trow = fast_div(t - ts.tile_first[l], ts.tiles_x[l], ts.inv_tiles_x[l], tcol);
    const int y = (trow << p.tile_h_log2) + (tq0 >> p.tile_w_log2);
    const int x = (tcol << p.tile_w_log2) + (tq0 & ((1 << p.tile_w_log2) - 1));
    q0 = ts.lv[l].qstart + y * ts.lv[l].W + x;
    nq = (y < ts.lv[l].H) ? ts.lv[l].W - x : 0;
  }
  nq = nq < 0 ? 0 : (nq > qpp ? qpp : nq);
}

// ---------------------------------------------------------------------------
// Vector kernel.
//   T     element type (float / __half / __nv_bfloat16)
//   D     channels per head (D*sizeof(T) multiple of 16, G = D*sizeof(T)/16 <= 32)
//   P_T   points per level known at compile time (4) or 0 = run-time
//   SPLIT lanes groups co-operating on one pair: the P points of a level are
//         dealt round-robin to SPLIT sub-groups and reduced with shuffles
//         (small-Q / decoder shapes, to expose more parallelism)
//   MATH  kExact: fp32 weights ; kFhfma: 16-bit weights + FHFMA
// ---------------------------------------------------------------------------
template <typename T, int D, int P_T, int SPLIT, int MATH, bool STAGE, bool FUSED = false, bool DYN = false>
__global__ void __launch_bounds__(kThreads, (SPLIT > 1 && P_T == 4) ? MSDA_MINB_SPLIT : MSDA_MINB) msda_fwd_vec(const MsdaParams p) {
  static_assert(!(DYN && STAGE), "shared-memory staging needs CTA-wide units");
  constexpr int E = (int)sizeof(T);
  constexpr int VEC = 16 / E;            // channels per lane
  constexpr int G = D / VEC;             // lanes per corner row
  constexpr int GS = G * SPLIT;          // lanes per (query, head) pair
  constexpr int PPW = 32 / GS;           // pairs per warp
  constexpr int PAIRS_PER_PASS = kThreads / GS;
  constexpr int NB = MSDA_NB;            // samples consumed per batch in the broadcast path
  static_assert(D % VEC == 0 && GS <= 32 && (GS & (GS - 1)) == 0, "unsupported D / SPLIT");

  __shared__ TileSetup ts;
  __shared__ __align__(8) uint64_t stage_bar[2];
  extern __shared__ __align__(128) unsigned char stage_mem[];  // STAGE: 2 x (qpp loc rows + qpp weight rows)
  // PDL: the level table (constant for a model / engine) is read before waiting for the preceding kernel;
  // everything that kernel may have produced (value, locations, weights) is read after the wait.
  pdl_launch_dependents();
  if (!p.pdl_early_tables) {
    pdl_wait_prior_grid();  // default: nothing at all is read before the preceding kernel is done
    prefetch_value_l2(p, E);
  }
  if (threadIdx.x < 32) setup_tiles(p, ts);
  if constexpr (STAGE) {
    if (threadIdx.x == 32) {
      mbar_init(&stage_bar[0], 1);
      mbar_init(&stage_bar[1], 1);
      fence_mbar_init();
    }
  }
  __syncthreads();
  if (p.pdl_early_tables) {
    pdl_wait_prior_grid();
    prefetch_value_l2(p, E);
  }

  const char *__restrict__ value = static_cast<const char *>(p.value);
  // fused mode: "loc" are the raw sampling offsets and "wgt" the pre-softmax logits (same layouts)
  const T *__restrict__ loc = static_cast<const T *>(FUSED ? p.offsets : p.loc);
  const T *__restrict__ wgt = static_cast<const T *>(FUSED ? p.logits : p.weight);
  T *__restrict__ out = static_cast<T *>(p.out);

  const int P = P_T ? P_T : p.P;
  const int LP = p.L * P;
  const int M = p.M;
  const unsigned pix_bytes = (unsigned)(M * D * E);
  const int lane_in_pair = threadIdx.x % GS;
  const int sub = lane_in_pair % G;     // which 16-byte piece of the row
  const int split = lane_in_pair / G;   // which share of the points
  const int slots = M << (p.tile_w_log2 + p.tile_h_log2);
  // one loop iteration = one pass of the CTA over PAIRS_PER_PASS (query, head) pairs of one tile of
  // image blockIdx.y; the host guarantees that the pass count stays below 2^22
  const int total = ts.n_tiles * p.passes;
  const int b = blockIdx.y;

  // A pass is split over the CTA's warps: warp-slice k of a pass holds slots [k*32/GS, (k+1)*32/GS).  With
  // static scheduling k is the warp's own index; with dynamic scheduling (DYN) a warp draws (pass, k) units.
  constexpr int WPC = kThreads / 32;
  const int lane_slot = (int)((threadIdx.x & 31) / GS);
  // slot -> (query-in-pass, head) when a pass is a whole number of queries (qpp > 0)
  auto slot_map = [&](int ls, int &tql_out, int &m_out) {
    if (p.head_major) {
      const int g = ls & (PPW - 1);
      const int qb = fast_div(ls / PPW, M, p.inv_M, m_out);
      tql_out = qb * PPW + g;
    } else {
      tql_out = fast_div(ls, M, p.inv_M, m_out);
    }
  };
  int tql = 0, m_fixed = 0;  // static scheduling: this thread's fixed slot
  if (p.qpp) slot_map((int)(threadIdx.x / GS), tql, m_fixed);

  // STAGE: the locations / weights of the pass's queries are copied into shared memory by 1-D TMA bulk
  // copies (one padded row per query, so the lane groups of a warp read distinct banks), double
  // buffered: the copies of pass i+1 are in flight while pass i is computed.
  const int stage_buf_bytes = p.qpp * (p.stage_loc_row + p.stage_w_row);
  auto stage_issue = [&](int w_next, int buf) {
    int pass;
    const int t = fast_div(w_next, p.passes, p.inv_passes, pass);
    int q0, nq;
    pass_query_range(p, ts, t, pass * p.qpp, p.qpp, q0, nq);
    const unsigned loc_bytes = (unsigned)(M * LP * 2 * E), w_bytes = (unsigned)(M * LP * E);
    mbar_expect_tx(&stage_bar[buf], (unsigned)nq * (loc_bytes + w_bytes));
    unsigned char *dl = stage_mem + buf * stage_buf_bytes;
    unsigned char *dw = dl + p.qpp * p.stage_loc_row;
    const char *gl = reinterpret_cast<const char *>(loc) + ((size_t)b * p.Q + q0) * loc_bytes;
    const char *gw = reinterpret_cast<const char *>(wgt) + ((size_t)b * p.Q + q0) * w_bytes;
    for (int i = 0; i < nq; ++i) {
      tma_bulk_g2s(dl + i * p.stage_loc_row, gl + (size_t)i * loc_bytes, loc_bytes, &stage_bar[buf]);
      tma_bulk_g2s(dw + i * p.stage_w_row, gw + (size_t)i * w_bytes, w_bytes, &stage_bar[buf]);
    }
  };

  // (tile, pass) unit w, warp-slice k -> this lane group's query (or -1 for a padding slot) and head
  auto decode_unit = [&](int w, int k, int &q, int &m) {
    int pass;
    const int t = fast_div(w, p.passes, p.inv_passes, pass);
    const int ls = k * (32 / GS) + lane_slot;
    int tq;
    if (p.qpp) {
      int tq_local = tql;
      m = m_fixed;
      if constexpr (DYN) slot_map(ls, tq_local, m);
      tq = pass * p.qpp + tq_local;
      q = tile_query(p, ts, t, tq);
    } else {
      const int s = pass * PAIRS_PER_PASS + ls;
      if (p.head_major) {
        // slots ordered [query block of PPW][head][query in block]: a warp holds one head of PPW
        // neighbouring queries
        const int g = s & (PPW - 1);
        const int qb = fast_div(s / PPW, M, p.inv_M, m);
        tq = qb * PPW + g;
      } else {
        tq = fast_div(s, M, p.inv_M, m);
      }
      q = (s < slots) ? tile_query(p, ts, t, tq) : -1;
    }
  };
  auto pair_index = [&](int q, int m) -> int64_t { return q >= 0 ? ((int64_t)b * p.Q + q) * M + m : 0; };

  // The CTA is persistent (it strides over the units), so the first sample of the NEXT unit is loaded
  // while the current one is computed: `carry` holds its undecoded location / weight.
  constexpr bool kCarry = (P_T == 4 && SPLIT == 1 && G >= 4 && !STAGE);
  int q = -1, m = 0;
  RawSample carry = {0u, 0u, 0u};
  // Unit assignment: `chunked` gives every CTA one contiguous run of units -- it then sweeps whole tiles
  // row by row, so the corner rows shared by vertically adjacent queries are still in this SM's L1 --
  // otherwise units are dealt round-robin (stride = grid size).
  int w_begin, w_end, w_step;
  if (p.chunked) {
    const int per = (total + (int)gridDim.x - 1) / (int)gridDim.x;
    w_begin = (int)blockIdx.x * per;
    w_end = min(total, w_begin + per);
    w_step = 1;
  } else {
    w_begin = blockIdx.x;
    w_end = total;
    w_step = gridDim.x;
  }
  // DYN: warp-units are drawn from the launch's counter; lane 0 draws, the warp shares the value
  const int warp_id = (int)(threadIdx.x >> 5);
  const int total_wu = total * WPC;
  unsigned *const sched = p.sched + (size_t)blockIdx.y * 2;  // one counter pair per image (grid.y)
  auto draw = [&]() -> int {
    unsigned v = 0;
    if ((threadIdx.x & 31) == 0) v = atomicAdd(sched, 1u);
    return (int)__shfl_sync(0xffffffffu, v, 0);
  };
  int w = w_begin, k = warp_id;
  bool have = w_begin < w_end;
  if constexpr (DYN) {
    const int wu = draw();
    have = wu < total_wu;
    w = wu / WPC;
    k = wu % WPC;
  }
  if (have) {
    decode_unit(w, k, q, m);
    if constexpr (kCarry) {
      const int64_t pr = pair_index(q, m);
      carry = load_raw<T>(loc + pr * LP * 2, wgt + pr * LP, sub & 3);
    }
    if constexpr (STAGE) {
      if (threadIdx.x == 0) stage_issue(w_begin, 0);
    }
  }

  int it = 0;
  while (have) {
    // the unit after this one (static: arithmetic; dynamic: drawn now, so the atomic's latency hides behind
    // the work of the current unit)
    int w_next = w + w_step, k_next = k;
    bool have_next = w_next < w_end;
    if constexpr (DYN) {
      const int wu = draw();
      have_next = wu < total_wu;
      w_next = wu / WPC;
      k_next = wu % WPC;
    }
    int q_next = -1, m_next = m;
    if (have_next) decode_unit(w_next, k_next, q_next, m_next);
    const unsigned char *slp = nullptr, *swp = nullptr;
    if constexpr (STAGE) {
      const int buf = it & 1;
      __syncthreads();  // every thread has finished reading buffer buf^1 (previous pass)
      if (threadIdx.x == 0 && have_next) stage_issue(w_next, buf ^ 1);
      mbar_wait(&stage_bar[buf], (unsigned)((it >> 1) & 1));
      slp = stage_mem + buf * stage_buf_bytes + tql * p.stage_loc_row + m * (LP * 2 * E);
      swp = stage_mem + buf * stage_buf_bytes + p.qpp * p.stage_loc_row + tql * p.stage_w_row + m * (LP * E);
    }
    // Padding slots (tile edge, tail of the last pass) stay in the loop with all their weights forced
    // to zero instead of branching out: the warp stays converged, so the shuffles below can name all 32
    // lanes with a compile-time mask.  They read the inputs of pair 0 (always present) and store nothing.
    const bool live = q >= 0;
    {
      const int64_t pair = live ? ((int64_t)b * p.Q + q) * M + m : 0;
      const T *lp = loc + pair * LP * 2;
      const T *wp = wgt + pair * LP;
      const char *vm = value + ((size_t)b * p.S * M + m) * (size_t)(D * E) + (size_t)sub * 16;
      // keep the row base as one opaque 64-bit register pair: every corner address is then a single
      // IMAD.WIDE (index * pixel pitch + base) instead of a re-derivation from the kernel parameters
      asm volatile("" : "+l"(vm));

      float acc[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[i] = 0.f;

      if constexpr (P_T == 4 && SPLIT == 1 && G >= 4) {
        // ---- broadcast path: lane (sub & 3) of the group works out the geometry of sample (sub & 3)
        // of the level once, the G lanes that consume it receive it by warp shuffle.  Everything that
        // must not contribute (corner outside the level, sample outside the range test) has weight
        // zero, and a zero weight predicates both the load and the FMAs of that corner off.
        constexpr unsigned group_mask = 0xffffffffu;
        const int ks = sub & 3;
        RawSample raw;
        if constexpr (STAGE) raw = load_raw_smem<T>(slp, swp, ks);
        else raw = carry;  // loaded during the previous unit (or before the loop)
        // fused mode: softmax statistics of the pair's L*4 logits.  Lane ks holds the logits of point ks
        // of every level; maximum and sum are completed across the four point-lanes with two butterflies.
        float sm_max = 0.f, sm_inv = 1.f;
        const T *rfp = nullptr;
        if constexpr (FUSED) {
          float mx = -INFINITY;
          for (int l = 0; l < p.L; ++l) mx = fmaxf(mx, Elem<T>::to_acc(wp[l * 4 + ks]));
          mx = fmaxf(mx, __shfl_xor_sync(group_mask, mx, 1));
          mx = fmaxf(mx, __shfl_xor_sync(group_mask, mx, 2));
          float sum = 0.f;
          for (int l = 0; l < p.L; ++l) sum += fused_exp<T>(Elem<T>::to_acc(wp[l * 4 + ks]) - mx);
          sum += __shfl_xor_sync(group_mask, sum, 1);
          sum += __shfl_xor_sync(group_mask, sum, 2);
          sm_max = mx;
          sm_inv = 1.f / sum;
          rfp = static_cast<const T *>(p.ref) + ((int64_t)b * p.Q + (live ? q : 0)) * p.L * p.ref_dim;
        }
        // Software-pipelined variant (MSDA_PIPE): the row loads of sample s+1 are issued before the FMAs of
        // sample s, across the samples of a level and across levels, so a lane always has 4-8 row loads in
        // flight instead of 4 -> 0 -> 4.  Two row buffers (32 registers): built with 3 CTAs per SM.
        constexpr bool kPipe = (MSDA_PIPE != 0) && !STAGE && !FUSED;
        if constexpr (kPipe) {
          struct InFlight {
            uint4 r[4];
            unsigned p0, p1;
            float w[4];
          };
          // geometry of this lane's sample (point ks) of level l, from the prefetched undecoded inputs
          auto level_geo = [&](int l, const RawSample &rw, int &gi, unsigned &gp0, unsigned &gp1, float (&gcw)[4]) {
            float x, y, aw;
            decode_raw<T>(rw, x, y, aw);
            aw = live ? aw : 0.f;
            make_geo(x, y, aw, ts.lv[l].H, ts.lv[l].W, gi, gcw);
            gi += ts.lv[l].start;
            gp0 = gp1 = 0u;
            if constexpr (MATH == kFhfma) {
              gp0 = pack_weights<T>(gcw[0], gcw[1]);
              gp1 = pack_weights<T>(gcw[2], gcw[3]);
            }
          };
          auto issue = [&](InFlight &f, int gi, unsigned gp0, unsigned gp1, const float (&gcw)[4], int k, int W) {
            const int bi = __shfl_sync(group_mask, gi, k, G);
            if constexpr (MATH == kFhfma) {
              f.p0 = __shfl_sync(group_mask, gp0, k, G);
              f.p1 = __shfl_sync(group_mask, gp1, k, G);
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) f.w[j] = __shfl_sync(group_mask, gcw[j], k, G);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int idx = bi + (j & 1) + ((j & 2) ? W : 0);
              bool on;
              if constexpr (MATH == kFhfma) on = (((j & 2) ? f.p1 : f.p0) >> ((j & 1) * 16) & 0x7fffu) != 0u;
              else on = f.w[j] != 0.f;
              if (on) f.r[j] = ldg128(vm + (size_t)(unsigned)idx * pix_bytes);
            }
          };
          auto consume = [&](const InFlight &f) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if constexpr (MATH == kFhfma) {
                const unsigned w16 = (((j & 2) ? f.p1 : f.p0) >> ((j & 1) * 16)) & 0xffffu;
                if ((w16 & 0x7fffu) != 0u) RowFma<T, kFhfma>::run(acc, f.r[j], 0.f, w16);
              } else {
                if (f.w[j] != 0.f) RowFma<T, kExact>::run(acc, f.r[j], f.w[j], 0u);
              }
            }
          };
          auto prefetch_after = [&](int l) {  // inputs of level l+1 of this unit, or level 0 of the next unit
            if (l + 1 < p.L) return load_raw<T>(lp, wp, (l + 1) * 4 + ks);
            const int64_t pn = pair_index(q_next, m_next);
            return load_raw<T>(loc + pn * LP * 2, wgt + pn * LP, ks);
          };
          InFlight fa, fb;
          int gi;
          unsigned gp0, gp1;
          float gcw[4];
          level_geo(0, raw, gi, gp0, gp1, gcw);
          raw = prefetch_after(0);
          issue(fa, gi, gp0, gp1, gcw, 0, ts.lv[0].W);
          for (int l = 0; l < p.L; ++l) {
            const int W = ts.lv[l].W;
            issue(fb, gi, gp0, gp1, gcw, 1, W);
            consume(fa);
            issue(fa, gi, gp0, gp1, gcw, 2, W);
            consume(fb);
            issue(fb, gi, gp0, gp1, gcw, 3, W);
            consume(fa);
            if (l + 1 < p.L) {
              level_geo(l + 1, raw, gi, gp0, gp1, gcw);
              raw = prefetch_after(l + 1);
              issue(fa, gi, gp0, gp1, gcw, 0, ts.lv[l + 1].W);
            }
            consume(fb);
          }
        } else
        for (int l = 0; l < p.L; ++l) {
          const int H = ts.lv[l].H, W = ts.lv[l].W;
          float x, y, aw;
          decode_raw<T>(raw, x, y, aw);
          if constexpr (FUSED) {
            // x, y are the raw offsets, aw the logit
            const float ox = x, oy = y;
            fused_location_fast<T>(rfp + l * p.ref_dim, p.ref_dim, ox, oy, (float)W, (float)H, ts.lv[l].rcpW, ts.lv[l].rcpH, x, y);
            aw = round_like<T, float>(fused_exp<T>(aw - sm_max) * sm_inv);
          }
          aw = live ? aw : 0.f;
          // prefetch: next level of this unit, or the first level of the next unit
          if constexpr (STAGE) {
            if (l + 1 < p.L) raw = load_raw_smem<T>(slp, swp, (l + 1) * 4 + ks);
          } else {
            if (l + 1 < p.L) {
              raw = load_raw<T>(lp, wp, (l + 1) * 4 + ks);
            } else {
              const int64_t pn = pair_index(q_next, m_next);
              raw = load_raw<T>(loc + pn * LP * 2, wgt + pn * LP, ks);
            }
          }
          int i00;
          float cw[4];
          make_geo(x, y, aw, H, W, i00, cw);
          i00 += ts.lv[l].start;
          unsigned pk0 = 0, pk1 = 0;
          if constexpr (MATH == kFhfma) {
            pk0 = pack_weights<T>(cw[0], cw[1]);
            pk1 = pack_weights<T>(cw[2], cw[3]);
          }
          // consume the four samples NB at a time: NB*4 row loads in flight per lane
#pragma unroll
          for (int k0 = 0; k0 < 4; k0 += NB) {
            uint4 rows[NB][4];
            float bw[NB][4];
            unsigned bp[NB][2];
#pragma unroll
            for (int kk = 0; kk < NB; ++kk) {
              const int k = k0 + kk;
              const int bi = __shfl_sync(group_mask, i00, k, G);
              if constexpr (MATH == kFhfma) {
                bp[kk][0] = __shfl_sync(group_mask, pk0, k, G);
                bp[kk][1] = __shfl_sync(group_mask, pk1, k, G);
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) bw[kk][j] = __shfl_sync(group_mask, cw[j], k, G);
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int idx = bi + (j & 1) + ((j & 2) ? W : 0);
                bool on;
                if constexpr (MATH == kFhfma) on = ((bp[kk][j >> 1] >> ((j & 1) * 16)) & 0x7fffu) != 0u;
                else on = bw[kk][j] != 0.f;
                if (on) rows[kk][j] = ldg128(vm + (size_t)(unsigned)idx * pix_bytes);
              }
            }
#pragma unroll
            for (int kk = 0; kk < NB; ++kk) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if constexpr (MATH == kFhfma) {
                  const unsigned w16 = (bp[kk][j >> 1] >> ((j & 1) * 16)) & 0xffffu;
                  if ((w16 & 0x7fffu) != 0u) RowFma<T, kFhfma>::run(acc, rows[kk][j], 0.f, w16);
                } else {
                  if (bw[kk][j] != 0.f) RowFma<T, kExact>::run(acc, rows[kk][j], bw[kk][j], 0u);
                }
              }
            }
          }
        }
        if constexpr (kCarry) carry = raw;
      } else if constexpr (P_T == 4) {
        // ---- split-points path (small Q, e.g. the decoder's 900 queries): the four points of a level are
        // dealt to SPLIT lane groups, and the row loads of LB levels are all issued before the first FMA,
        // so one lane has LB * (4/SPLIT) * 4 independent loads in flight -- the problem is too small to
        // hide memory latency with occupancy.
        constexpr int SPL = 4 / SPLIT;   // samples per level per lane group
        constexpr int LB = MSDA_LB;      // levels per batch
        for (int l0 = 0; l0 < p.L; l0 += LB) {
          uint4 rows[LB][SPL][4];
          float cwb[LB][SPL][4];
#pragma unroll
          for (int lb = 0; lb < LB; ++lb) {
            const int l = l0 + lb;
#pragma unroll
            for (int ss = 0; ss < SPL; ++ss) {
#pragma unroll
              for (int j = 0; j < 4; ++j) cwb[lb][ss][j] = 0.f;
            }
            if (l < p.L) {
              const int H = ts.lv[l].H, W = ts.lv[l].W;
              const char *vl = vm + (size_t)ts.lv[l].start * pix_bytes;
#pragma unroll
              for (int ss = 0; ss < SPL; ++ss) {
                float x, y, aw;
                load_sample_inputs<T>(lp, wp, l * 4 + split + ss * SPLIT, x, y, aw);
                aw = live ? aw : 0.f;
                int i00;
                make_geo(x, y, aw, H, W, i00, cwb[lb][ss]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int idx = i00 + (j & 1) + ((j & 2) ? W : 0);
                  if (cwb[lb][ss][j] != 0.f) rows[lb][ss][j] = ldg128(vl + (size_t)(unsigned)idx * pix_bytes);
                }
              }
            }
          }
#pragma unroll
          for (int lb = 0; lb < LB; ++lb) {
#pragma unroll
            for (int ss = 0; ss < SPL; ++ss) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (cwb[lb][ss][j] != 0.f) {
                  if constexpr (MATH == kFhfma) RowFma<T, kFhfma>::run(acc, rows[lb][ss][j], 0.f, weight_to_16<T>(cwb[lb][ss][j]));
                  else RowFma<T, kExact>::run(acc, rows[lb][ss][j], cwb[lb][ss][j], 0u);
                }
              }
            }
          }
        }
      } else {
        // ---- run-time P: every lane works out the geometry of the samples it consumes
        for (int l = 0; l < p.L; ++l) {
          const int H = ts.lv[l].H, W = ts.lv[l].W;
          const char *vl = vm + (size_t)ts.lv[l].start * pix_bytes;
          for (int k = split; k < P; k += SPLIT) {
            float x, y, aw;
            load_sample_inputs<T>(lp, wp, l * P + k, x, y, aw);
            aw = live ? aw : 0.f;
            int i00;
            float cw[4];
            make_geo(x, y, aw, H, W, i00, cw);
            uint4 rows[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int idx = i00 + (j & 1) + ((j & 2) ? W : 0);
              if (cw[j] != 0.f) rows[j] = ldg128(vl + (size_t)(unsigned)idx * pix_bytes);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (cw[j] != 0.f) {
                if constexpr (MATH == kFhfma) {
                  const unsigned w16 = weight_to_16<T>(cw[j]);
                  RowFma<T, kFhfma>::run(acc, rows[j], 0.f, w16);
                } else {
                  RowFma<T, kExact>::run(acc, rows[j], cw[j], 0u);
                }
              }
            }
          }
        }
      }

      if constexpr (SPLIT > 1) {
#pragma unroll
        for (int off = G; off < GS; off <<= 1) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
        }
      }
      if (live && split == 0) store_row<T, VEC>(out + pair * D + sub * VEC, acc);
    }
    q = q_next;
    m = m_next;
    w = w_next;
    k = k_next;
    have = have_next;
    ++it;
  }
  if constexpr (DYN) {
    // this warp has drawn its last unit; the last warp of the image's grid row to get here re-arms the counters
    if ((threadIdx.x & 31) == 0) {
      const unsigned finished = atomicAdd(sched + 1, 1u);
      if (finished == gridDim.x * WPC - 1) {
        sched[0] = 0u;
        sched[1] = 0u;
        __threadfence();
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Small-problem kernel (decoder cross-attention: 900 queries = 7,200 pairs).  The general kernel spends a
// third of its instructions on per-CTA set-up and tile decode, which a launch this small cannot amortise
// (ncu: 3.46 M instructions for 7,200 pairs, issue-bound at 10 us even with warm L2).  Here: no shared
// memory, no barrier, no tiles -- lane l of every warp loads level l's (H, W, start) and hands it out by
// shuffle; 4 lane groups of G lanes share a pair (one point of each level per group) and are reduced
// with butterflies; 128-thread CTAs so that 7,200 pairs spread over all 148 SMs.
// ---------------------------------------------------------------------------
constexpr int kSmallThreads = 128;

template <typename T, int D, int MATH>
__global__ void __launch_bounds__(kSmallThreads, 8) msda_fwd_small(const MsdaParams p) {
  constexpr int E = (int)sizeof(T), VEC = 16 / E, G = D / VEC, GS = G * 4;
  static_assert(GS <= 32, "rows wider than 8 lanes do not fit the 4-way point split");
  const char *__restrict__ value = static_cast<const char *>(p.value);
  const T *__restrict__ loc = static_cast<const T *>(p.loc);
  const T *__restrict__ wgt = static_cast<const T *>(p.weight);
  T *__restrict__ out = static_cast<T *>(p.out);
  const int lane = threadIdx.x & 31;
  const int sub = lane % G, split = (lane / G) & 3;
  const int M = p.M, LP = p.L * 4;
  const unsigned pix_bytes = (unsigned)(M * D * E);

  pdl_launch_dependents();
  if (!p.pdl_early_tables) pdl_wait_prior_grid();
  // level table: lane l holds level l (L <= 32 guaranteed by the host)
  int lvH = 0, lvW = 0, lvS = 0;
  if (lane < p.L) {
    lvH = (int)__ldg(p.shapes + 2 * lane);
    lvW = (int)__ldg(p.shapes + 2 * lane + 1);
    lvS = (int)__ldg(p.starts + lane);
  }
  if (p.pdl_early_tables) pdl_wait_prior_grid();  // value / locations / weights may come from the preceding kernel

  const int64_t pairs = (int64_t)p.B * p.Q * M;
  const int64_t pr = ((int64_t)blockIdx.x * kSmallThreads + threadIdx.x) / GS;
  const bool live = pr < pairs;
  const int64_t pair = live ? pr : 0;
  const int m = (int)(pair % M);
  const int64_t b = pair / M / p.Q;
  const T *lp = loc + pair * LP * 2;
  const T *wp = wgt + pair * LP;
  const char *vm = value + ((size_t)b * p.S * M + m) * (size_t)(D * E) + (size_t)sub * 16;

  // all of this lane's samples (point `split` of every level) are requested before anything is decoded
  constexpr int kMaxL = 8;
  RawSample raw[kMaxL];
#pragma unroll
  for (int l = 0; l < kMaxL; ++l)
    if (l < p.L) raw[l] = load_raw<T>(lp, wp, l * 4 + split);

  float acc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = 0.f;

#pragma unroll
  for (int l = 0; l < kMaxL; ++l) {
    if (l < p.L) {  // uniform across the grid
      const int H = __shfl_sync(0xffffffffu, lvH, l), W = __shfl_sync(0xffffffffu, lvW, l);
      const int start = __shfl_sync(0xffffffffu, lvS, l);
      float x, y, aw;
      decode_raw<T>(raw[l], x, y, aw);
      aw = live ? aw : 0.f;
      int i00;
      float cw[4];
      make_geo(x, y, aw, H, W, i00, cw);
      i00 += start;
      uint4 rows[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int idx = i00 + (j & 1) + ((j & 2) ? W : 0);
        if (cw[j] != 0.f) rows[j] = ldg128(vm + (size_t)(unsigned)idx * pix_bytes);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (cw[j] != 0.f) {
          if constexpr (MATH == kFhfma) RowFma<T, kFhfma>::run(acc, rows[j], 0.f, weight_to_16<T>(cw[j]));
          else RowFma<T, kExact>::run(acc, rows[j], cw[j], 0u);
        }
      }
    }
  }
#pragma unroll
  for (int off = G; off < GS; off <<= 1) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
  }
  if (live && split == 0) store_row<T, VEC>(out + pair * D + sub * VEC, acc);
}

// ---------------------------------------------------------------------------
// Packed path (16-bit types, D = 32, P = 4): pixel-pair packed pyramid + 256-bit loads.
//
// In the channels-last value tensor a (pixel, head) row is 64 B, half an L1 line, so each of the four
// bilinear corners costs a full L1 wavefront.  The pre-pass below re-lays the pyramid out so that one
// 128-byte, line-aligned entry per (pixel, head) holds the row of the pixel AND of its right-hand
// neighbour (zeros at the end of an image row), chunk-interleaved:
//     entry(s, m) = [ v(s)[0:8] | v(s+1)[0:8] | v(s)[8:16] | v(s+1)[8:16] | ... ]   (4 x 32 B)
// A lane group then fetches BOTH horizontal corners of a sample with one LDG.E.256 per lane
// (ld.global.nc.v8.b32, sm_100+): two wavefronts per sample instead of four.
// ---------------------------------------------------------------------------
struct U8 {
  unsigned v[8];
};

__device__ __forceinline__ U8 ldg256(const void *ptr) {
  U8 r;
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(ptr));
  return r;
}
__device__ __forceinline__ void stg256(void *ptr, const uint4 &a, const uint4 &b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x),
               "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}

// pre-pass: one thread per (pixel, head, 16-byte chunk)
template <typename T>
__global__ void __launch_bounds__(kThreads) msda_pack_value(const MsdaParams p) {
  __shared__ TileSetup ts;
  if (threadIdx.x < 32) setup_tiles(p, ts);
  __syncthreads();
  if (!ts.layout_ok) return;  // exotic level layout: the main kernel takes its generic fallback
  const char *__restrict__ value = static_cast<const char *>(p.value);
  char *__restrict__ packed = static_cast<char *>(p.packed);
  const int M = p.M;
  const long long n = (long long)p.B * p.S * M * 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i & 3);
    const long long pm = i >> 2;          // (b*S + s)*M + m
    const long long bs = pm / M;
    const int s = (int)(bs % p.S);
    // does pixel s have a right-hand neighbour in its image row?
    bool has_right = false;
    for (int l = 0; l < p.L; ++l) {
      const int rel = s - ts.lv[l].start;
      if (rel >= 0 && rel < ts.lv[l].H * ts.lv[l].W) {
        const int W = ts.lv[l].W;
        has_right = (rel % W) + 1 < W;
      }
    }
    const uint4 left = __ldg(reinterpret_cast<const uint4 *>(value + pm * 64 + j * 16));
    uint4 right = make_uint4(0u, 0u, 0u, 0u);
    if (has_right) right = __ldg(reinterpret_cast<const uint4 *>(value + (pm + M) * 64 + j * 16));
    stg256(packed + pm * 128 + j * 32, left, right);
  }
}

// fp32-weight FMA of one 16-byte piece (8 channels)
template <typename T, int MATH>
__device__ __forceinline__ void piece_fma(float (&acc)[8], const unsigned *r, float w, unsigned w16) {
  const uint4 q = make_uint4(r[0], r[1], r[2], r[3]);
  RowFma<T, MATH>::run(acc, q, w, w16);
}

template <typename T, int MATH>
__global__ void __launch_bounds__(kThreads, MSDA_MINB) msda_fwd_packed(const MsdaParams p) {
  constexpr int D = 32, E = 2, G = 4, PPW = 8, PAIRS_PER_PASS = kThreads / G;
  __shared__ TileSetup ts;
  if (threadIdx.x < 32) setup_tiles(p, ts);
  __syncthreads();

  const T *__restrict__ loc = static_cast<const T *>(p.loc);
  const T *__restrict__ wgt = static_cast<const T *>(p.weight);
  T *__restrict__ out = static_cast<T *>(p.out);
  const int M = p.M, LP = p.L * 4;

  if (!ts.layout_ok) {
    // Levels that overlap or leave [0, S): the packed pyramid is not defined.  Correct but slow
    // element-wise fallback on the original tensors (same arithmetic as msda_fwd_generic).
    const T *__restrict__ value = static_cast<const T *>(p.value);
    const long long n = (long long)p.Q * M * D;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const int c = (int)(i % D);
      const long long pair_in_img = i / D;
      const int m = (int)(pair_in_img % M);
      const long long pair = (long long)blockIdx.y * p.Q * M + pair_in_img;
      const T *vb = value + ((long long)blockIdx.y * p.S) * M * D + (long long)m * D + c;
      float acc = 0.f;
      for (int l = 0; l < p.L; ++l) {
        const int H = (int)p.shapes[2 * l], W = (int)p.shapes[2 * l + 1];
        const T *vl = vb + p.starts[l] * (long long)M * D;
        for (int k = 0; k < 4; ++k) {
          const long long si = pair * LP + l * 4 + k;
          const Sample<float> sm = make_sample<float>(Elem<T>::to_acc(loc[si * 2]), Elem<T>::to_acc(loc[si * 2 + 1]),
                                                      Elem<T>::to_acc(wgt[si]), H, W);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (sm.ok[j]) acc += sm.cw[j] * Elem<T>::to_acc(vl[(long long)sm.idx[j] * M * D]);
        }
      }
      out[pair * D + c] = Elem<T>::from_acc(acc);
    }
    return;
  }

  const char *__restrict__ packed = static_cast<const char *>(p.packed);
  const unsigned ent_pitch = (unsigned)(M * 128);
  const int sub = threadIdx.x % G;
  const int ks = sub;
  const int ls = (int)(threadIdx.x / G);
  const int slots = M << (p.tile_w_log2 + p.tile_h_log2);
  const int total = ts.n_tiles * p.passes;
  const int b = blockIdx.y;
  int tql = 0, m_fixed = 0;
  if (p.qpp) {
    if (p.head_major) {
      const int g = ls & (PPW - 1);
      const int qb = fast_div(ls / PPW, M, p.inv_M, m_fixed);
      tql = qb * PPW + g;
    } else {
      tql = fast_div(ls, M, p.inv_M, m_fixed);
    }
  }

  for (int w = blockIdx.x; w < total; w += gridDim.x) {
    int q, m;
    {
      int pass;
      const int t = fast_div(w, p.passes, p.inv_passes, pass);
      int tq;
      if (p.qpp) {
        tq = pass * p.qpp + tql;
        m = m_fixed;
        q = tile_query(p, ts, t, tq);
      } else {
        const int s = pass * PAIRS_PER_PASS + ls;
        if (p.head_major) {
          const int g = s & (PPW - 1);
          const int qb = fast_div(s / PPW, M, p.inv_M, m);
          tq = qb * PPW + g;
        } else {
          tq = fast_div(s, M, p.inv_M, m);
        }
        q = (s < slots) ? tile_query(p, ts, t, tq) : -1;
      }
    }
    const bool live = q >= 0;
    const int64_t pair = live ? ((int64_t)b * p.Q + q) * M + m : 0;
    const T *lp = loc + pair * LP * 2;
    const T *wp = wgt + pair * LP;
    const char *vm = packed + ((size_t)b * p.S * M + m) * 128 + (size_t)sub * 32;
    asm volatile("" : "+l"(vm));

    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;

    float nx, ny, naw;
    load_sample_inputs<T>(lp, wp, ks, nx, ny, naw);
    for (int l = 0; l < p.L; ++l) {
      const int H = ts.lv[l].H, W = ts.lv[l].W;
      const float x = nx, y = ny, aw = live ? naw : 0.f;
      if (l + 1 < p.L) load_sample_inputs<T>(lp, wp, (l + 1) * 4 + ks, nx, ny, naw);
      // geometry of sample ks: entry of the top row + weights (left, right) x (top, bottom)
      int e_top;
      float cw[4];
      {
        const float w_im = __fmul_rn(x, (float)W) - 0.5f;
        const float h_im = __fmul_rn(y, (float)H) - 0.5f;
        const bool inside = (h_im > -1.f) && (w_im > -1.f) && (h_im < (float)H) && (w_im < (float)W);
        const float hf = floorf(h_im), wf = floorf(w_im);
        const int h_lo = (int)hf, w_lo = (int)wf;
        const float lh = h_im - hf, lw = w_im - wf;
        const float hh = 1.f - lh, hw = 1.f - lw;
        const float wy0 = (inside && h_lo >= 0) ? hh * aw : 0.f;
        const float wy1 = (inside && h_lo < H - 1) ? lh * aw : 0.f;
        // column -1: the entry of column 0 is used and its LEFT half is the sample's right-hand corner
        const bool neg = w_lo < 0;
        const float wl = inside ? (neg ? lw : hw) : 0.f;
        const float wr = (inside && !neg && w_lo < W - 1) ? lw : 0.f;
        cw[0] = wy0 * wl;
        cw[1] = wy0 * wr;
        cw[2] = wy1 * wl;
        cw[3] = wy1 * wr;
        e_top = ts.lv[l].start + h_lo * W + (neg ? 0 : w_lo);
      }
      unsigned pk0 = 0, pk1 = 0;
      if constexpr (MATH == kFhfma) {
        pk0 = pack_weights<T>(cw[0], cw[1]);
        pk1 = pack_weights<T>(cw[2], cw[3]);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int be = __shfl_sync(0xffffffffu, e_top, k, G);
        U8 rt, rb;
        if constexpr (MATH == kFhfma) {
          const unsigned bt = __shfl_sync(0xffffffffu, pk0, k, G);
          const unsigned bb = __shfl_sync(0xffffffffu, pk1, k, G);
          if ((bt & 0x7fff7fffu) != 0u) rt = ldg256(vm + (size_t)(unsigned)be * ent_pitch);
          if ((bb & 0x7fff7fffu) != 0u) rb = ldg256(vm + (size_t)(unsigned)(be + W) * ent_pitch);
          if ((bt & 0x7fffu) != 0u) piece_fma<T, kFhfma>(acc, &rt.v[0], 0.f, bt & 0xffffu);
          if ((bt & 0x7fff0000u) != 0u) piece_fma<T, kFhfma>(acc, &rt.v[4], 0.f, bt >> 16);
          if ((bb & 0x7fffu) != 0u) piece_fma<T, kFhfma>(acc, &rb.v[0], 0.f, bb & 0xffffu);
          if ((bb & 0x7fff0000u) != 0u) piece_fma<T, kFhfma>(acc, &rb.v[4], 0.f, bb >> 16);
        } else {
          float bw[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) bw[j] = __shfl_sync(0xffffffffu, cw[j], k, G);
          if (bw[0] != 0.f || bw[1] != 0.f) rt = ldg256(vm + (size_t)(unsigned)be * ent_pitch);
          if (bw[2] != 0.f || bw[3] != 0.f) rb = ldg256(vm + (size_t)(unsigned)(be + W) * ent_pitch);
          if (bw[0] != 0.f) piece_fma<T, kExact>(acc, &rt.v[0], bw[0], 0u);
          if (bw[1] != 0.f) piece_fma<T, kExact>(acc, &rt.v[4], bw[1], 0u);
          if (bw[2] != 0.f) piece_fma<T, kExact>(acc, &rb.v[0], bw[2], 0u);
          if (bw[3] != 0.f) piece_fma<T, kExact>(acc, &rb.v[4], bw[3], 0u);
        }
      }
    }
    if (live) store_row<T, 8>(out + pair * D + sub * 8, acc);
  }
}

#include "msda_fwd_hp.cuh"

// ---------------------------------------------------------------------------
// Backward: gradients w.r.t. value (scatter-add with atomics, like the
// reference), sampling locations and attention weights.  Per-sample derivative of
// ms_deform_attn.cu:79-141; the reference's six kernel variants (:263-760) differ only in how the
// per-channel partial sums are reduced -- here a lane group owns a (query, head) pair, each lane holds
// D/G channels, and the four corner dot products are reduced over the group with shuffles.
//   grad_value   accumulated into the caller's zero-initialised buffer (codetr/ops.py:94-96)
//   grad_loc, grad_weight   fully overwritten
// ---------------------------------------------------------------------------
struct BwdParams {
  const void *value;
  const int64_t *shapes;
  const int64_t *starts;
  const void *loc;
  const void *weight;
  const void *grad_out;
  void *grad_value;
  void *grad_loc;
  void *grad_weight;
  int B, S, M, D, L, Q, P;
};

template <typename T>
__device__ __forceinline__ void atomic_add_elem(T *addr, typename Elem<T>::acc_t v) {
  atomicAdd(addr, Elem<T>::from_acc(v));
}

// geometry of one sample for the backward: corner indices, validity, weights and their derivatives
template <typename A>
struct BwdSample {
  int idx[4];
  bool ok[4];
  A cw[4], dh[4], dw[4];
  bool inside;
};

template <typename A>
__device__ __forceinline__ BwdSample<A> make_bwd_sample(A x, A y, int H, int W) {
  BwdSample<A> s;
  const A w_im = mul_rn(x, (A)W) - (A)0.5;
  const A h_im = mul_rn(y, (A)H) - (A)0.5;
  s.inside = (h_im > (A)-1) && (w_im > (A)-1) && (h_im < (A)H) && (w_im < (A)W);
  const A hf = floor_acc<A>(h_im), wf = floor_acc<A>(w_im);
  const int h_lo = s.inside ? (int)hf : 0, w_lo = s.inside ? (int)wf : 0;
  const A lh = h_im - hf, lw = w_im - wf, hh = (A)1 - lh, hw = (A)1 - lw;
  const bool top = h_lo >= 0, bot = h_lo + 1 <= H - 1, lef = w_lo >= 0, rig = w_lo + 1 <= W - 1;
  s.ok[0] = s.inside && top && lef;
  s.ok[1] = s.inside && top && rig;
  s.ok[2] = s.inside && bot && lef;
  s.ok[3] = s.inside && bot && rig;
  const int base = h_lo * W + w_lo;
  s.idx[0] = base;
  s.idx[1] = base + 1;
  s.idx[2] = base + W;
  s.idx[3] = base + W + 1;
  s.cw[0] = hh * hw; s.cw[1] = hh * lw; s.cw[2] = lh * hw; s.cw[3] = lh * lw;
  s.dh[0] = -hw; s.dh[1] = -lw; s.dh[2] = hw; s.dh[3] = lw;    // d cw / d h_im
  s.dw[0] = -hh; s.dw[1] = hh; s.dw[2] = -lh; s.dw[3] = lh;    // d cw / d w_im
  return s;
}

// Generic backward: one thread per (pair, sample), serial over the D channels.  Any shape, any dtype.
template <typename T>
__global__ void __launch_bounds__(kThreads) msda_bwd_generic(const BwdParams p) {
  using A = typename Elem<T>::acc_t;
  const T *__restrict__ value = static_cast<const T *>(p.value);
  const T *__restrict__ loc = static_cast<const T *>(p.loc);
  const T *__restrict__ wgt = static_cast<const T *>(p.weight);
  const T *__restrict__ go = static_cast<const T *>(p.grad_out);
  T *gv = static_cast<T *>(p.grad_value);
  T *gl = static_cast<T *>(p.grad_loc);
  T *gw = static_cast<T *>(p.grad_weight);
  const int LP = p.L * p.P;
  const int64_t n = (int64_t)p.B * p.Q * p.M * LP;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int lp = (int)(i % LP);
    const int64_t pair = i / LP;
    const int l = lp / p.P;
    const int m = (int)(pair % p.M);
    const int64_t b = pair / p.M / p.Q;
    const int H = (int)p.shapes[2 * l], W = (int)p.shapes[2 * l + 1];
    const int64_t lvl = ((int64_t)b * p.S + p.starts[l]) * p.M * p.D + (int64_t)m * p.D;
    const A aw = Elem<T>::to_acc(wgt[i]);
    const BwdSample<A> s = make_bwd_sample<A>(Elem<T>::to_acc(loc[i * 2]), Elem<T>::to_acc(loc[i * 2 + 1]), H, W);
    A g_x = 0, g_y = 0, g_w = 0;
    if (s.inside) {
      for (int c = 0; c < p.D; ++c) {
        const A tg = Elem<T>::to_acc(go[pair * p.D + c]);
        const A tgv = tg * aw;
        A val = 0, gh = 0, gwd = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (!s.ok[j]) continue;
          const int64_t off = lvl + (int64_t)s.idx[j] * p.M * p.D + c;
          const A v = Elem<T>::to_acc(value[off]);
          val += s.cw[j] * v;
          gh += s.dh[j] * v;
          gwd += s.dw[j] * v;
          atomic_add_elem<T>(gv + off, s.cw[j] * tgv);
        }
        g_w += tg * val;
        g_x += (A)W * gwd * tgv;
        g_y += (A)H * gh * tgv;
      }
    }
    gl[i * 2] = Elem<T>::from_acc(g_x);
    gl[i * 2 + 1] = Elem<T>::from_acc(g_y);
    gw[i] = Elem<T>::from_acc(g_w);
  }
}

// 16-byte vector reduction into grad_value (REDG.E.ADD.F16x8 / BF16x8 / F32x4 on sm_100a)
template <typename T>
__device__ __forceinline__ void red_add_row(void *addr, const float *v);
template <>
__device__ __forceinline__ void red_add_row<float>(void *addr, const float *v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
}
template <>
__device__ __forceinline__ void red_add_row<__half>(void *addr, const float *v) {
  asm volatile("red.global.add.noftz.v4.f16x2 [%0], {%1,%2,%3,%4};" ::"l"(addr), "r"(pack_weights<__half>(v[0], v[1])),
               "r"(pack_weights<__half>(v[2], v[3])), "r"(pack_weights<__half>(v[4], v[5])), "r"(pack_weights<__half>(v[6], v[7]))
               : "memory");
}
template <>
__device__ __forceinline__ void red_add_row<__nv_bfloat16>(void *addr, const float *v) {
  asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1,%2,%3,%4};" ::"l"(addr), "r"(pack_weights<__nv_bfloat16>(v[0], v[1])),
               "r"(pack_weights<__nv_bfloat16>(v[2], v[3])), "r"(pack_weights<__nv_bfloat16>(v[4], v[5])),
               "r"(pack_weights<__nv_bfloat16>(v[6], v[7]))
               : "memory");
}

template <typename T, int VEC>
__device__ __forceinline__ void unpack_row(const uint4 &r, float (&f)[VEC]) {
  if constexpr (sizeof(T) == 4) {
    f[0] = __uint_as_float(r.x); f[1] = __uint_as_float(r.y); f[2] = __uint_as_float(r.z); f[3] = __uint_as_float(r.w);
  } else {
    const unsigned w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = unpack2<T>(w[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
}

// Vector backward: lane group of G = D*sizeof(T)/16 lanes per (query, head) pair.
template <typename T, int D>
__global__ void __launch_bounds__(kThreads) msda_bwd_vec(const BwdParams p) {
  constexpr int E = (int)sizeof(T), VEC = 16 / E, G = D / VEC, PAIRS_PER_CTA = kThreads / G;
  const char *__restrict__ value = static_cast<const char *>(p.value);
  const T *__restrict__ loc = static_cast<const T *>(p.loc);
  const T *__restrict__ wgt = static_cast<const T *>(p.weight);
  const char *__restrict__ go = static_cast<const char *>(p.grad_out);
  char *gv = static_cast<char *>(p.grad_value);
  T *gl = static_cast<T *>(p.grad_loc);
  T *gw = static_cast<T *>(p.grad_weight);
  const int LP = p.L * p.P;
  const int sub = threadIdx.x % G;
  const size_t pix_bytes = (size_t)p.M * D * E;
  const int64_t pairs = (int64_t)p.B * p.Q * p.M;
  // every lane of a warp runs the same trip count (dead pairs carry zero grad_out), so the group
  // reductions below can use the full mask
  const int64_t trips = (pairs + (int64_t)gridDim.x * PAIRS_PER_CTA - 1) / ((int64_t)gridDim.x * PAIRS_PER_CTA);
  for (int64_t tr = 0; tr < trips; ++tr) {
    const int64_t pr = (tr * gridDim.x + blockIdx.x) * PAIRS_PER_CTA + threadIdx.x / G;
    const bool live = pr < pairs;
    const int64_t pair = live ? pr : 0;
    const int m = (int)(pair % p.M);
    const int64_t b = pair / p.M / p.Q;
    float g[VEC];
    {
      uint4 raw = make_uint4(0u, 0u, 0u, 0u);
      if (live) raw = ldg128(go + (size_t)pair * D * E + (size_t)sub * 16);
      unpack_row<T, VEC>(raw, g);
    }
    for (int l = 0; l < p.L; ++l) {
      const int H = (int)__ldg(p.shapes + 2 * l), W = (int)__ldg(p.shapes + 2 * l + 1);
      const size_t lvl = (((size_t)b * p.S + (size_t)__ldg(p.starts + l)) * p.M + m) * (size_t)(D * E) + (size_t)sub * 16;
      for (int k = 0; k < p.P; ++k) {
        const int64_t si = pair * LP + (int64_t)l * p.P + k;
        float x, y, aw;
        load_sample_inputs<T>(loc + pair * LP * 2, wgt + pair * LP, l * p.P + k, x, y, aw);
        const BwdSample<float> s = make_bwd_sample<float>(x, y, H, W);
        float t[4];  // per-corner dot product of grad_out with the value row (this lane's channels)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          t[j] = 0.f;
          if (s.ok[j]) {
            const size_t off = lvl + (size_t)(unsigned)s.idx[j] * pix_bytes;
            float v[VEC];
            unpack_row<T, VEC>(ldg128(value + off), v);
            float dv[VEC];
#pragma unroll
            for (int c = 0; c < VEC; ++c) {
              t[j] = fmaf(g[c], v[c], t[j]);
              dv[c] = s.cw[j] * aw * g[c];
            }
            if (live) red_add_row<T>(gv + off, dv);
          }
        }
#pragma unroll
        for (int off = 1; off < G; off <<= 1) {
#pragma unroll
          for (int j = 0; j < 4; ++j) t[j] += __shfl_xor_sync(0xffffffffu, t[j], off);
        }
        if (live && sub == 0) {
          float val = 0.f, gh = 0.f, gwd = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            val = fmaf(s.cw[j], t[j], val);
            gh = fmaf(s.dh[j], t[j], gh);
            gwd = fmaf(s.dw[j], t[j], gwd);
          }
          // samples outside the range test have all-zero t[] (no corner was read): gradients are zero
          gl[si * 2] = Elem<T>::from_acc(s.inside ? (float)W * gwd * aw : 0.f);
          gl[si * 2 + 1] = Elem<T>::from_acc(s.inside ? (float)H * gh * aw : 0.f);
          gw[si] = Elem<T>::from_acc(s.inside ? val : 0.f);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// read-bandwidth probe (roofline denominators: L2 -> SM, HBM -> SM)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) read_probe_kernel(const uint4 *__restrict__ buf, size_t n_vec, int repeats,
                                                              unsigned *sink) {
  unsigned acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int r = 0; r < repeats; ++r) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
      uint4 v;
      // ld.global.cv would bypass L2 too; .cg caches in L2 only, which is what we want to measure
      asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(buf + i));
      acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
  }
  if (acc == 0x9e3779b9u) *sink = acc;  // practically never true; keeps the loads alive
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
size_t elem_size(int dtype) {
  switch (dtype) {
    case MSDA_F32: return 4;
    case MSDA_F16: return 2;
    case MSDA_BF16: return 2;
    case MSDA_F64: return 8;
    default: return 0;
  }
}

const char *dtype_name(int dtype) {
  switch (dtype) {
    case MSDA_F32: return "f32";
    case MSDA_F16: return "f16";
    case MSDA_BF16: return "bf16";
    case MSDA_F64: return "f64";
    default: return "?";
  }
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// Counter pairs for one launch with dynamic unit scheduling (one pair per image), or nullptr when the pool
// cannot be used: symbol lookup failed, more images than a slot run holds, or the stream is being captured.
unsigned *sched_slots(int images, cudaStream_t stream) {
  constexpr int kMaxImages = 64;
  if (images > kMaxImages) return nullptr;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) return nullptr;
  static unsigned *base[64] = {nullptr};
  static std::atomic<unsigned> next{0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!base[dev]) {
    void *ptr = nullptr;
    if (cudaGetSymbolAddress(&ptr, g_sched) != cudaSuccess) return nullptr;
    base[dev] = static_cast<unsigned *>(ptr);
  }
  const unsigned first = next.fetch_add((unsigned)images, std::memory_order_relaxed) % (unsigned)(kSchedSlots - kMaxImages);
  return base[dev] + (size_t)first * 2;
}

int env_int(const char *name, int fallback) {
  const char *v = getenv(name);
  if (!v || !*v) return fallback;
  return atoi(v);
}

template <typename T>
int launch_generic(const MsdaParams &p, cudaStream_t stream) {
  const int64_t n = (int64_t)p.B * p.Q * p.M * p.D;
  int64_t blocks = (n + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)sm_count() * 32;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  msda_fwd_generic<T><<<(unsigned)blocks, kThreads, 0, stream>>>(p);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

// Launch with or without the programmatic-stream-serialization attribute.
template <typename... KArgs, typename... Args>
cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                          Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

struct VecPlan {
  int split;
  int math;
  unsigned grid, grid_y;
  unsigned stage_bytes;  // dynamic shared memory of the TMA staging buffers, 0 = direct loads
  bool pdl;              // launch with programmatic stream serialization (MSDA_FLAG_PDL)
  bool dyn;              // warps draw their units from a device counter (p.sched) instead of striding
};

template <typename T, int D, int P_T, int SPLIT, int MATH>
int launch_vec_inst(const MsdaParams &p, const VecPlan &plan, cudaStream_t stream) {
  constexpr int G = D * (int)sizeof(T) / 16;
  if constexpr (P_T == 4 && SPLIT == 1 && G >= 4) {
    if (p.ref_dim != 0) {
      const cudaError_t fe = launch_kernel(msda_fwd_vec<T, D, P_T, SPLIT, MATH, false, true>, dim3(plan.grid, plan.grid_y, 1),
                                           dim3(kThreads), 0, stream, plan.pdl, p);
      g_launch_count.fetch_add(1, std::memory_order_relaxed);
      return fe != cudaSuccess ? (int)fe : (int)cudaGetLastError();
    }
    if (plan.stage_bytes > 0) {
      const cudaError_t se = launch_kernel(msda_fwd_vec<T, D, P_T, SPLIT, MATH, true, false>, dim3(plan.grid, plan.grid_y, 1),
                                           dim3(kThreads), plan.stage_bytes, stream, plan.pdl, p);
      g_launch_count.fetch_add(1, std::memory_order_relaxed);
      return se != cudaSuccess ? (int)se : (int)cudaGetLastError();
    }
  }
  if constexpr (P_T == 4 && SPLIT == 1) {
    if (plan.dyn && p.sched) {
      const cudaError_t de = launch_kernel(msda_fwd_vec<T, D, P_T, SPLIT, MATH, false, false, true>,
                                           dim3(plan.grid, plan.grid_y, 1), dim3(kThreads), 0, stream, plan.pdl, p);
      g_launch_count.fetch_add(1, std::memory_order_relaxed);
      return de != cudaSuccess ? (int)de : (int)cudaGetLastError();
    }
  }
  const cudaError_t le = launch_kernel(msda_fwd_vec<T, D, P_T, SPLIT, MATH, false, false>, dim3(plan.grid, plan.grid_y, 1),
                                       dim3(kThreads), 0, stream, plan.pdl, p);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return le != cudaSuccess ? (int)le : (int)cudaGetLastError();
}

template <typename T, int D, int MATH>
int launch_vec_d(const MsdaParams &p, const VecPlan &plan, cudaStream_t stream) {
  constexpr int G = D * (int)sizeof(T) / 16;
  if (plan.split == 4) {
    if constexpr (G * 4 <= 32) {
      return p.P == 4 ? launch_vec_inst<T, D, 4, 4, MATH>(p, plan, stream)
                      : launch_vec_inst<T, D, 0, 4, MATH>(p, plan, stream);
    }
  }
  if (plan.split == 2) {
    if constexpr (G * 2 <= 32) {
      if (p.P == 4) return launch_vec_inst<T, D, 4, 2, MATH>(p, plan, stream);
    }
  }
  return p.P == 4 ? launch_vec_inst<T, D, 4, 1, MATH>(p, plan, stream)
                  : launch_vec_inst<T, D, 0, 1, MATH>(p, plan, stream);
}

template <typename T>
int launch_vec_t(const MsdaParams &p, const VecPlan &plan, cudaStream_t stream) {
  constexpr bool k16 = sizeof(T) == 2;
  if (k16 && plan.math == kFhfma) {
    if constexpr (k16) {
      switch (p.D) {
        case 16: return launch_vec_d<T, 16, kFhfma>(p, plan, stream);
        case 32: return launch_vec_d<T, 32, kFhfma>(p, plan, stream);
        case 64: return launch_vec_d<T, 64, kFhfma>(p, plan, stream);
        default: return MSDA_ERR_UNSUPPORTED;
      }
    }
  }
  switch (p.D) {
    case 16: return launch_vec_d<T, 16, kExact>(p, plan, stream);
    case 32: return launch_vec_d<T, 32, kExact>(p, plan, stream);
    case 64: return launch_vec_d<T, 64, kExact>(p, plan, stream);
    default: return MSDA_ERR_UNSUPPORTED;
  }
}

int run_generic(const MsdaParams &p, int dtype, cudaStream_t stream) {
  int rc;
  switch (dtype) {
    case MSDA_F32: rc = launch_generic<float>(p, stream); break;
    case MSDA_F16: rc = launch_generic<__half>(p, stream); break;
    case MSDA_BF16: rc = launch_generic<__nv_bfloat16>(p, stream); break;
    case MSDA_F64: rc = launch_generic<double>(p, stream); break;
    default: return MSDA_ERR_BAD_DTYPE;
  }
  if (rc == 0) snprintf(g_last_variant, sizeof(g_last_variant), "generic<%s>%s", dtype_name(dtype), p.ref_dim ? "/fused" : "");
  return rc;
}

// Warps per CTA of the head-pair kernel: every warp of the cpg CTAs of a head pair walks the query quads with one stride,
// so a call takes ceil(quads / (cpg * nw)) rounds; pick the nw that wastes the least of the last round -- 4,604 quads over
// 37 CTAs: 24 warps -> 6 rounds, 86 % busy; 25 warps -> 5 rounds, 99.5 %.  Fewer warps hide less latency: a configuration
// with half the warps must be 14 % better balanced to win.
int hp_pick_warps(int64_t quads, int cpg) {
  int pick = kHpThreads / 32;
  double best = -1.0;
  for (int nw = kHpThreads / 32; nw >= kHpThreads / 64 && cpg > 0; --nw) {
    const int64_t per_round = (int64_t)cpg * nw;
    const int64_t rounds = (quads + per_round - 1) / per_round;
    const double eff = rounds > 0 ? (double)quads / (double)(rounds * per_round) : 0.0;
    const double score = eff * (0.75 + 0.25 * (double)nw / (double)(kHpThreads / 32));
    if (score > best + 1e-4) {
      best = score;
      pick = nw;
    }
  }
  return pick;
}

bool aligned_to(const void *ptr, size_t a) { return (reinterpret_cast<uintptr_t>(ptr) % a) == 0; }

// Bytes of workspace the packed path needs, or 0 when it does not apply / would not pay off: 16-bit
// types, D = 32, P = 4, and enough gather work per packed row (the pre-pass touches every key once, the
// gather touches Q*L*P*4 rows per head: the decoder's 900 queries would not amortise it).
size_t packed_workspace_bytes(int64_t B, int64_t S, int64_t M, int64_t D, int64_t Q, int64_t L, int64_t P, int dtype) {
  if ((dtype != MSDA_F16 && dtype != MSDA_BF16) || D != 32 || P != 4 || B <= 0 || S <= 0 || M <= 0) return 0;
  if (Q * L * P * 4 < (int64_t)env_int("MSDA_B200_PACKED_RATIO", 16) * S) return 0;
  return (size_t)B * (size_t)S * (size_t)M * 128;
}

int forward_impl(MsdaParams p, int dtype, unsigned flags, void *workspace, size_t workspace_bytes, cudaStream_t stream) {
  const size_t E = elem_size(dtype);
  const bool fused = p.ref_dim != 0;

  // ---- choose the kernel ----
  bool vec_ok = !(flags & MSDA_FLAG_FORCE_GENERIC) && dtype != MSDA_F64 &&
                (p.D == 16 || p.D == 32 || p.D == 64) && p.L <= kMaxLevelsSmem &&
                aligned_to(p.value, 16) && aligned_to(p.out, 16) && ((size_t)p.M * p.D * E) % 16 == 0 &&
                // the fast paths read an (x, y) pair with one load: a contiguous view at an odd element offset takes
                // the element-wise kernel instead of faulting
                aligned_to(fused ? p.offsets : p.loc, 2 * E) &&
                // nothing to sample (L*P == 0, pointers may be NULL): the element-wise kernel writes the zeros
                p.L > 0 && p.P > 0;
  const int G = vec_ok ? (int)(p.D * E / 16) : 1;
  // the fused-producer mode lives on the broadcast path only (P = 4, rows of at least four lanes)
  if (fused && !(p.P == 4 && G >= 4)) vec_ok = false;

  if (!vec_ok) return run_generic(p, dtype, stream);

  VecPlan plan;
  const int sms = sm_count();
  const int64_t pairs = (int64_t)p.B * p.Q * p.M;
  // Small problems only (the decoder's 900 queries: 7,200 pairs = 113 CTAs for 148 SMs): deal the points
  // of a level to 4 (or 2) lane groups so every SM gets work.  As soon as the un-split grid covers the
  // SMs the un-split kernel wins (measured: R50 encoder 608x608 28 us un-split vs 51 us split).
  const int64_t ctas_unsplit = (pairs * G + kThreads - 1) / kThreads;
  plan.split = 1;
  if (ctas_unsplit < env_int("MSDA_B200_SPLIT_MAX_CTAS", sms) && p.P == 4 && !fused) plan.split = (G * 4 <= 32) ? 4 : ((G * 2 <= 32) ? 2 : 1);
  if (!fused) plan.split = env_int("MSDA_B200_SPLIT", plan.split);
  if (!((plan.split == 4 && G * 4 <= 32) || (plan.split == 2 && G * 2 <= 32 && p.P == 4))) plan.split = 1;

  // fp16 defaults to the FHFMA path (combined weights rounded to fp16: measured max-normalised error
  // 4.7e-4 vs 2.7e-4 for fp32 weights at the headline shape, gate 2e-3); bf16 weights would keep only 8
  // bits, so bf16 defaults to fp32 weights
  // Programmatic dependent launch is on by default in its conservative form: this kernel's CTAs may be
  // scheduled while the preceding kernel of the stream drains (hides launch latency and fills its tail),
  // but every thread waits for that kernel to complete before it reads anything.  MSDA_FLAG_PDL additionally
  // reads the two level tables before the wait.  MSDA_B200_PDL=0 turns the attribute off.
  plan.pdl = env_int("MSDA_B200_PDL", 1) != 0;
  p.pdl_early_tables = (plan.pdl && (flags & MSDA_FLAG_PDL)) ? 1 : 0;
  plan.math = (dtype == MSDA_F16) ? kFhfma : kExact;
  if (E == 2) {
    if (flags & MSDA_FLAG_MATH_FHFMA) plan.math = kFhfma;
    if (flags & MSDA_FLAG_MATH_EXACT) plan.math = kExact;
  }
  if (E != 2) plan.math = kExact;

  const int ppw = 32 / (G * plan.split);
  const int pairs_per_pass = kThreads / (G * plan.split);
  auto ceil_log2 = [](int v) { int l = 0; while ((1 << l) < v) ++l; return l; };
  // tile geometry: power-of-two extents, width at least the pairs a warp holds in head-major order
  p.want_tiled = (flags & MSDA_FLAG_LINEAR_ORDER) ? 0 : (p.Q == p.S ? 1 : 0);
  // slot order inside a pass: query-major (a warp holds the heads of one query: its location / weight
  // loads are contiguous) measured 4-5 % faster than head-major (a warp holds one head of neighbouring
  // queries) once the kernel stopped being wavefront-bound; MSDA_FLAG_HEAD_MAJOR / the env knob switch
  p.head_major = env_int("MSDA_B200_HEAD_MAJOR", (flags & MSDA_FLAG_HEAD_MAJOR) ? 1 : 0);
  p.chunked = env_int("MSDA_B200_CHUNKED", 0);  // measured: contiguous runs are 3 % slower (imbalance) than round-robin
  int tile_w = env_int("MSDA_B200_TILE_W", 8);
  int tile_h = env_int("MSDA_B200_TILE_H", p.want_tiled ? 4 : 1);
  if (!p.want_tiled) {
    // linear chunks: about one pass of the CTA per chunk
    tile_w = (pairs_per_pass + p.M - 1) / p.M;
    tile_h = 1;
  }
  if (tile_w < ppw) tile_w = ppw;
  if (tile_h < 1) tile_h = 1;
  p.tile_w_log2 = ceil_log2(tile_w);
  p.tile_h_log2 = ceil_log2(tile_h);
  if (p.tile_w_log2 + p.tile_h_log2 > 12) return MSDA_ERR_UNSUPPORTED;
  const int64_t tile_q = (int64_t)1 << (p.tile_w_log2 + p.tile_h_log2);
  const int64_t passes = (tile_q * p.M + pairs_per_pass - 1) / pairs_per_pass;
  p.passes = (int)passes;
  p.inv_passes = 1.0f / (float)passes;
  p.inv_M = 1.0f / (float)p.M;
  // a pass is a whole number of queries when M divides the pairs per pass (and, head-major, the block
  // of ppw queries divides it too); then it also divides the power-of-two tile width
  p.qpp = 0;
  if (pairs_per_pass % p.M == 0) {
    const int qpp = pairs_per_pass / p.M;
    if ((qpp & (qpp - 1)) == 0 && qpp % ppw == 0 && qpp <= (1 << p.tile_w_log2)) p.qpp = qpp;
  }
  // TMA staging of locations / weights: P == 4 broadcast path, whole-query passes, 16-byte aligned rows
  plan.stage_bytes = 0;
  p.stage_loc_row = p.stage_w_row = 0;
  const bool want_stage = ((flags & MSDA_FLAG_STAGE_TMA) || env_int("MSDA_B200_STAGE", 0)) && !(flags & MSDA_FLAG_NO_STAGING);
  if (p.qpp > 0 && p.P == 4 && plan.split == 1 && G >= 4 && want_stage && !fused) {
    const size_t loc_row = (size_t)p.M * p.L * p.P * 2 * E, w_row = (size_t)p.M * p.L * p.P * E;
    auto pad_row = [](size_t bytes) {  // row pitch = 4 (mod 8) words: the 8 lane groups of a warp hit distinct banks
      size_t r = bytes;
      while ((r / 4) % 8 != 4) r += 16;
      return r;
    };
    if (loc_row % 16 == 0 && w_row % 16 == 0 && aligned_to(p.loc, 16) && aligned_to(p.weight, 16)) {
      p.stage_loc_row = (int)pad_row(loc_row);
      p.stage_w_row = (int)pad_row(w_row);
      const size_t total = 2 * (size_t)p.qpp * (p.stage_loc_row + p.stage_w_row);
      if (total <= 40 * 1024) plan.stage_bytes = (unsigned)total;
    }
  }

  // grid: x strides over the (tile, pass) units of one image, y = image.  The level shapes are
  // device-resident, so the exact 2-D tile count is unknown here; the kernel strides over the real
  // count, so an estimate (interior tiles plus an allowance for the partial tiles on the right/bottom
  // edge of each level) is enough.
  int64_t tiles_est = (p.Q + tile_q - 1) / tile_q;
  if (p.want_tiled) tiles_est += tiles_est / 4 + 4 * p.L;
  // fast_div needs every dividend below 2^22; worst case of the real tile count: one query per tile
  if ((int64_t)p.Q * passes >= ((int64_t)1 << 22) || tile_q * p.M >= ((int64_t)1 << 22) || p.B > 65535) {
    return run_generic(p, dtype, stream);
  }
  // Persistent grid: exactly the resident CTA count (4 per SM) for launches of up to ~32 passes per SM,
  // twice that for larger ones; every CTA strides over the passes, so the per-CTA set-up (level table,
  // barrier) is paid once.  Measured vs one CTA per pass: 51.2 vs 54.7 us at the headline shape, 91.6 vs
  // 100.7 us in fp32 (profiles/r01_sweep_ctas.log).
  int64_t grid = tiles_est * passes;
  const int64_t per_sm = (grid * p.B + sms - 1) / sms;
  const int64_t cap = ((int64_t)sms * env_int("MSDA_B200_CTAS_PER_SM", per_sm <= 32 ? MSDA_MINB : 2 * MSDA_MINB) + p.B - 1) / p.B;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  plan.grid = (unsigned)grid;
  plan.grid_y = (unsigned)p.B;
  // Dynamic unit scheduling (MSDA_B200_DYN=1): warps draw (pass, warp-slice) units from a device counter, which
  // evens out launches whose unit count is a small non-integer multiple of the resident CTA count.
  // L2 prefetch of the value tensor: pays when the whole batch's pyramid is a small part of L2 (1,900-query decoder
  // 13.3 -> 12.3 us, test shape 10.3 -> 9.8, headline 49.05 -> 48.75), costs when it is not (decoder B=8, 75 MB of
  // value for 7,200 queries: 30.4 -> 32.9 us), so: on up to 16 MB.  MSDA_B200_L2_PREFETCH=0/1 forces it.
  const int64_t value_bytes = (int64_t)p.B * p.S * p.M * p.D * E;
  p.l2_prefetch = env_int("MSDA_B200_L2_PREFETCH", value_bytes <= (int64_t)16 * 1024 * 1024 ? 1 : 0);
  plan.dyn = false;
  p.sched = nullptr;
  if (env_int("MSDA_B200_DYN", 0) && plan.split == 1 && p.P == 4 && plan.stage_bytes == 0 && !fused && !p.chunked) {
    p.sched = sched_slots(p.B, stream);
    plan.dyn = p.sched != nullptr;
  }

  // PDL pays when the grid fills the machine (its CTAs can only land where the previous kernel's exit) or on
  // the small-problem kernel; a partial grid placed early piles onto the first SMs that free up and runs
  // unbalanced (measured: 1,900-query decoder 12.6 -> 14.5 us), so it is launched the plain way.
  const bool small_kernel = plan.split == 4 && !fused && p.P == 4 && p.L <= 8 && env_int("MSDA_B200_SMALL", 1);
  if (!small_kernel && (int64_t)plan.grid * plan.grid_y < (int64_t)sms * MSDA_MINB) {
    plan.pdl = false;
    p.pdl_early_tables = 0;
  }

  // ---- small-problem kernel (decoder): 4-way point split, no tiles, no barrier ----
  if (small_kernel) {
    const int64_t lanes = pairs * G * 4;
    const unsigned sgrid = (unsigned)((lanes + kSmallThreads - 1) / kSmallThreads);
    int rc3 = MSDA_ERR_UNSUPPORTED;
    auto launch_small = [&](auto tag_t, auto tag_d) {
      using TT = decltype(tag_t);
      constexpr int DD = decltype(tag_d)::value;
      if constexpr (DD * (int)sizeof(TT) / 16 * 4 <= 32) {
        cudaError_t se = cudaSuccess;
        if (sizeof(TT) == 2 && plan.math == kFhfma) {
          if constexpr (sizeof(TT) == 2) se = launch_kernel(msda_fwd_small<TT, DD, kFhfma>, dim3(sgrid), dim3(kSmallThreads), 0, stream, plan.pdl, p);
        } else {
          se = launch_kernel(msda_fwd_small<TT, DD, kExact>, dim3(sgrid), dim3(kSmallThreads), 0, stream, plan.pdl, p);
        }
        g_launch_count.fetch_add(1, std::memory_order_relaxed);
        rc3 = se != cudaSuccess ? (int)se : (int)cudaGetLastError();
      }
    };
    if (dtype == MSDA_F16 && p.D == 32) launch_small(__half{}, std::integral_constant<int, 32>{});
    else if (dtype == MSDA_F16 && p.D == 64) launch_small(__half{}, std::integral_constant<int, 64>{});
    else if (dtype == MSDA_BF16 && p.D == 32) launch_small(__nv_bfloat16{}, std::integral_constant<int, 32>{});
    else if (dtype == MSDA_BF16 && p.D == 64) launch_small(__nv_bfloat16{}, std::integral_constant<int, 64>{});
    else if (dtype == MSDA_F32 && p.D == 32) launch_small(float{}, std::integral_constant<int, 32>{});
    else if (dtype == MSDA_F32 && p.D == 16) launch_small(float{}, std::integral_constant<int, 16>{});
    if (rc3 != MSDA_ERR_UNSUPPORTED) {
      if (rc3 == 0)
        snprintf(g_last_variant, sizeof(g_last_variant), "small<%s,D%d,P4,split4>/%s", dtype_name(dtype), p.D,
                 plan.math == kFhfma ? "fhfma" : "exact");
      return rc3;
    }
  }

  // ---- head-pair kernel: coarse levels in shared memory (msda_fwd_hp.cuh) ----
  // fp16 / bf16 / fp32, D = 32, P = 4, eight heads, and enough query groups (4 queries per warp for the 16-bit types,
  // 2 for fp32, whose corner rows are whole 128-byte lines read by 8 lanes) to give every warp of every CTA work;
  // MSDA_FLAG_NO_SMEM_LEVELS / MSDA_B200_HP=0 fall back to the all-global kernel below.
  {
    const int NG = p.M / 2;
    const int qpw = E == 4 ? 2 : 4;
    // bf16 without a math flag: FHFMA with each weight as two bf16 terms (msda_fwd_hp.cuh, kFhfmaSplit)
    const bool hp_split = dtype == MSDA_BF16 && plan.math == kExact && !(flags & MSDA_FLAG_MATH_EXACT) && env_int("MSDA_B200_BF16_SPLIT", 1);
    const bool hp_shape = (E == 2 || (E == 4 && env_int("MSDA_B200_HP_F32", 1))) && p.D == 32 && p.P == 4 && p.M == 8 && p.L >= 2 && p.L <= kHpMaxLevels &&
                          NG <= sms && plan.split == 1 && plan.stage_bytes == 0 && !plan.dyn &&
                          // fused producers stay on the vector kernel unless MSDA_B200_HP_FUSED=1: bit-identical here, but the softmax
                          // and location arithmetic (+25 % instructions) cost this kernel more than that one -- 58.2 vs 54.1 us
                          // (fp16), 72.3 vs 67.2 (bf16), 104.7 vs 99.3 (fp32) at the headline shape; reference points are
                          // read as 32-bit words
                          (!fused || ((p.ref_dim == 2 || p.ref_dim == 4) && aligned_to(p.ref, 4) && aligned_to(p.logits, E) && aligned_to(p.offsets, 2 * E) &&
                                      env_int("MSDA_B200_HP_FUSED", 0))) &&
                          // 16-bit calls with MSDA_FLAG_MATH_EXACT stay on the vector kernel (bf16: 57.3 us here vs 58.9 there,
                          // fp16 slower here); MSDA_B200_HP_EXACT=1 overrides
                          (plan.math == kFhfma || E == 4 || hp_split || env_int("MSDA_B200_HP_EXACT", 0)) &&
                          (fused || (aligned_to(p.loc, 2 * E) && aligned_to(p.weight, E))) && !(flags & MSDA_FLAG_NO_SMEM_LEVELS) &&
                          // the kernel addresses an image's locations with 32-bit byte offsets
                          (int64_t)p.Q * p.M * p.L * 8 * E < ((int64_t)1 << 32);
    const int cpg = NG > 0 ? sms / NG : 0;
    const int64_t quads = ((int64_t)p.Q + qpw - 1) / qpw;
    const int64_t hp_min_quads = (int64_t)env_int("MSDA_B200_HP_MIN_QUADS_PER_WARP", 2) * cpg * (kHpThreads / 32);
    // Warps per CTA: every warp of the cpg CTAs of a head pair walks the query quads with the same stride, so the
    // call takes ceil(quads / (cpg * nw)) rounds; pick the nw (of the upper half of what the register budget
    // allows) that wastes the least of the last round -- 4,604 quads over 37 CTAs: 24 warps -> 6 rounds, 86 %
    // busy; 25 warps -> 5 rounds, 99.5 %.
    int hp_warps = hp_pick_warps(quads, cpg);
    {
      hp_warps = env_int("MSDA_B200_HP_WARPS", hp_warps);
      if (hp_warps < 1) hp_warps = 1;
      if (hp_warps > kHpThreads / 32) hp_warps = kHpThreads / 32;
      if (fused && hp_warps > kHpFusedThreads / 32) hp_warps = kHpFusedThreads / 32;
    }
    if (hp_shape && env_int("MSDA_B200_HP", 1) && quads >= hp_min_quads && !(workspace != nullptr && env_int("MSDA_B200_PACKED", 1))) {
      // Cached levels: the copy into shared memory costs every CTA ~1.5 us up front and takes L1 capacity away from
      // the fine levels; measured neutral at 5 rounds per warp (headline, 45.2 vs 44.9 us), -3 % at 20 rounds
      // (strides 4-64: 164.7 vs 170.2 us), +33 % on a 1,900-query decoder call.  So: only for long calls.
      const int64_t hp_rounds = (quads + (int64_t)cpg * hp_warps - 1) / ((int64_t)cpg * hp_warps);
      p.hp_smem_bytes = env_int("MSDA_B200_HP_SMEM", hp_rounds >= 12 ? 148 * 1024 : 0);
      if (p.hp_smem_bytes < 0) p.hp_smem_bytes = 0;
      if (p.hp_smem_bytes > 200 * 1024) p.hp_smem_bytes = 200 * 1024;
      const dim3 hgrid((unsigned)(cpg * NG), (unsigned)p.B, 1);
      auto launch_hp = [&](auto kernel) -> int {
        cudaError_t ae = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, p.hp_smem_bytes);
        if (ae != cudaSuccess) return (int)ae;
        const cudaError_t le = launch_kernel(kernel, hgrid, dim3((unsigned)hp_warps * 32u), (size_t)p.hp_smem_bytes, stream, plan.pdl, p);
        g_launch_count.fetch_add(1, std::memory_order_relaxed);
        return le != cudaSuccess ? (int)le : (int)cudaGetLastError();
      };
      // Dynamic unit scheduling (MSDA_B200_HP_DYN=1, opt-in): warps draw (quad, head pair) units from a device counter
      // to take out the 4 % spread between the fastest and the slowest SM of the static schedule.  Bit-identical, clean
      // code (the loop condition is a warp vote, which ptxas treats as uniform) -- and 16 % SLOWER (headline 51.9 vs
      // 44.5 us, R50 32.0 vs 21.5): 18,416 atomics on one address per launch serialise in one L2 slice at ~3 ns each,
      // which is longer than the kernel.  Drawing chunks would cut the atomics but leave 1.2 chunks of 4 per warp, a
      // worse balance than the static one.  Only without cached levels and outside stream capture.
      bool hp_dyn = false;
      if (p.hp_smem_bytes == 0 && !fused && env_int("MSDA_B200_HP_DYN", 0)) {
        p.sched = sched_slots(p.B, stream);
        hp_dyn = p.sched != nullptr;
        // no rounds to fill evenly any more: as many warps as the register budget allows
        if (hp_dyn) hp_warps = env_int("MSDA_B200_HP_WARPS", kHpThreads / 32);
        if (hp_warps < 1 || hp_warps > kHpThreads / 32) hp_warps = kHpThreads / 32;
      }
      int rch;
      if (fused) {
        if (dtype == MSDA_F32) rch = launch_hp(msda_fwd_hp<float, kExact, 8, false, false, true>);
        else if (dtype == MSDA_F16)
          rch = plan.math == kFhfma ? launch_hp(msda_fwd_hp<__half, kFhfma, 8, false, false, true>)
                                    : launch_hp(msda_fwd_hp<__half, kExact, 8, false, false, true>);
        else if (hp_split) rch = launch_hp(msda_fwd_hp<__nv_bfloat16, kFhfmaSplit, 8, false, false, true>);
        else
          rch = plan.math == kFhfma ? launch_hp(msda_fwd_hp<__nv_bfloat16, kFhfma, 8, false, false, true>)
                                    : launch_hp(msda_fwd_hp<__nv_bfloat16, kExact, 8, false, false, true>);
      } else if (dtype == MSDA_F32) {
        rch = hp_dyn ? launch_hp(msda_fwd_hp<float, kExact, 8, true>) : launch_hp(msda_fwd_hp<float, kExact, 8>);
      } else if (dtype == MSDA_F16) {
        if (hp_dyn) rch = plan.math == kFhfma ? launch_hp(msda_fwd_hp<__half, kFhfma, 8, true>) : launch_hp(msda_fwd_hp<__half, kExact, 8, true>);
        else rch = plan.math == kFhfma ? launch_hp(msda_fwd_hp<__half, kFhfma, 8>) : launch_hp(msda_fwd_hp<__half, kExact, 8>);
      } else if (hp_split && !hp_dyn) {
        rch = launch_hp(msda_fwd_hp<__nv_bfloat16, kFhfmaSplit, 8>);
      } else {
        if (hp_dyn) rch = plan.math == kFhfma ? launch_hp(msda_fwd_hp<__nv_bfloat16, kFhfma, 8, true>) : launch_hp(msda_fwd_hp<__nv_bfloat16, kExact, 8, true>);
        else rch = plan.math == kFhfma ? launch_hp(msda_fwd_hp<__nv_bfloat16, kFhfma, 8>) : launch_hp(msda_fwd_hp<__nv_bfloat16, kExact, 8>);
      }
      if (rch == 0)
        snprintf(g_last_variant, sizeof(g_last_variant), "hp<%s,D32,P4,M%d>/smem%dK/%dwarps/%s%s%s", dtype_name(dtype), p.M,
                 p.hp_smem_bytes / 1024, hp_warps, plan.math == kFhfma ? "fhfma" : (hp_split && !hp_dyn) ? "fhfma-split" : "exact", hp_dyn ? "/dyn" : "",
                 fused ? "/fused-producers" : "");
      return rch;
    }
  }

  // ---- packed path: pixel-pair packed pyramid in the caller's workspace + 256-bit loads ----
  const size_t packed_need = packed_workspace_bytes(p.B, p.S, p.M, p.D, p.Q, p.L, p.P, dtype);
  const bool packed_ok = packed_need > 0 && workspace != nullptr && workspace_bytes >= packed_need && aligned_to(workspace, 128) &&
                         plan.split == 1 && !fused && !(flags & MSDA_FLAG_NO_PACKED) && env_int("MSDA_B200_PACKED", 1) &&
                         aligned_to(p.loc, 4) && aligned_to(p.weight, 2);
  if (packed_ok) {
    p.packed = workspace;
    const long long chunks = (long long)p.B * p.S * p.M * 4;
    long long pgrid = (chunks + kThreads - 1) / kThreads;
    const long long pcap = (long long)sms * 16;
    if (pgrid > pcap) pgrid = pcap;
    int rc2;
    if (dtype == MSDA_F16) {
      msda_pack_value<__half><<<(unsigned)pgrid, kThreads, 0, stream>>>(p);
      if (plan.math == kFhfma) msda_fwd_packed<__half, kFhfma><<<dim3(plan.grid, plan.grid_y, 1), kThreads, 0, stream>>>(p);
      else msda_fwd_packed<__half, kExact><<<dim3(plan.grid, plan.grid_y, 1), kThreads, 0, stream>>>(p);
    } else {
      msda_pack_value<__nv_bfloat16><<<(unsigned)pgrid, kThreads, 0, stream>>>(p);
      if (plan.math == kFhfma) msda_fwd_packed<__nv_bfloat16, kFhfma><<<dim3(plan.grid, plan.grid_y, 1), kThreads, 0, stream>>>(p);
      else msda_fwd_packed<__nv_bfloat16, kExact><<<dim3(plan.grid, plan.grid_y, 1), kThreads, 0, stream>>>(p);
    }
    g_launch_count.fetch_add(2, std::memory_order_relaxed);
    rc2 = (int)cudaGetLastError();
    if (rc2 == 0) {
      snprintf(g_last_variant, sizeof(g_last_variant), "packed<%s,D32,P4>/%s%dx%d/%s/%s+pack-prepass", dtype_name(dtype),
               p.want_tiled ? "tiled" : "linear", 1 << p.tile_w_log2, 1 << p.tile_h_log2,
               p.head_major ? "head-major" : "query-major", plan.math == kFhfma ? "fhfma" : "exact");
    }
    return rc2;
  }

  int rc;
  switch (dtype) {
    case MSDA_F32: rc = launch_vec_t<float>(p, plan, stream); break;
    case MSDA_F16: rc = launch_vec_t<__half>(p, plan, stream); break;
    case MSDA_BF16: rc = launch_vec_t<__nv_bfloat16>(p, plan, stream); break;
    default: return MSDA_ERR_BAD_DTYPE;
  }
  if (rc == 0) {
    snprintf(g_last_variant, sizeof(g_last_variant), "vec<%s,D%d,P%d,split%d>/%s%dx%d/%s/%s%s", dtype_name(dtype), p.D,
             p.P == 4 ? 4 : 0, plan.split, p.want_tiled ? "tiled" : "linear", 1 << p.tile_w_log2, 1 << p.tile_h_log2,
             p.head_major ? "head-major" : "query-major", plan.math == kFhfma ? "fhfma" : "exact",
             fused ? "/fused-producers" : (plan.stage_bytes ? "/tma-staged" : (plan.dyn ? "/dyn" : "")));
  }
  return rc;
}

int validate_common(const void *value, const int64_t *shapes, const int64_t *starts, const void *out, int64_t B,
                    int64_t S, int64_t M, int64_t D, int64_t L, int64_t Q, int64_t P, int dtype) {
  if (elem_size(dtype) == 0) return MSDA_ERR_BAD_DTYPE;
  if (B < 0 || S < 0 || M < 0 || D < 0 || L < 0 || Q < 0 || P < 0) return MSDA_ERR_BAD_SHAPE;
  const int64_t lim = (int64_t)1 << 31;
  if (S >= lim || Q >= lim || M >= 65536 || D >= 65536 || L >= 65536 || P >= 65536 || B >= lim) return MSDA_ERR_UNSUPPORTED;
  if (B * Q * M * D == 0) return MSDA_OK;  // empty output: nothing to do, nothing to check
  if (!out) return MSDA_ERR_NULL_POINTER;
  if (L * P > 0 && S > 0 && (!value || !shapes || !starts)) return MSDA_ERR_NULL_POINTER;
  const size_t E = elem_size(dtype);
  if (!aligned_to(value, E) || !aligned_to(out, E) || !aligned_to(shapes, 8) || !aligned_to(starts, 8)) return MSDA_ERR_MISALIGNED;
  return MSDA_OK;
}

}  // namespace

void msda_detail::set_last_variant(const char *text) { snprintf(g_last_variant, sizeof(g_last_variant), "%s", text); }

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

int msda_b200_abi_version(void) { return MSDA_B200_ABI_VERSION; }

const char *msda_b200_error_string(int code) {
  switch (code) {
    case MSDA_OK: return "success";
    case MSDA_ERR_NULL_POINTER: return "msda_b200: a required pointer is NULL";
    case MSDA_ERR_BAD_SHAPE: return "msda_b200: negative dimension";
    case MSDA_ERR_BAD_DTYPE: return "msda_b200: unknown dtype";
    case MSDA_ERR_BAD_STEP: return "msda_b200: batch must be divisible by min(batch, im2col_step)";
    case MSDA_ERR_MISALIGNED: return "msda_b200: pointer not aligned to its element size";
    case MSDA_ERR_UNSUPPORTED: return "msda_b200: shape outside the supported index range";
    case MSDA_ERR_BAD_FLAGS: return "msda_b200: contradictory flags";
    case MSDA_ERR_WORKSPACE_TOO_SMALL: return "msda_b200: workspace smaller than the matching *_workspace_bytes()";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "msda_b200: unknown error";
}

uint64_t msda_b200_launch_count(void) { return g_launch_count.load(std::memory_order_relaxed); }

const char *msda_b200_last_variant(void) { return g_last_variant; }

uint64_t msda_b200_algorithmic_hbm_bytes(int64_t B, int64_t S, int64_t M, int64_t D, int64_t L, int64_t Q, int64_t P,
                                         int dtype) {
  const uint64_t E = elem_size(dtype);
  // value + locations + weights read once, output written once, plus the level table
  return E * (uint64_t)B * (uint64_t)(S * M * D + Q * M * L * P * 2 + Q * M * L * P + Q * M * D) + 8ull * 3ull * (uint64_t)L;
}

uint64_t msda_b200_algorithmic_gather_bytes(int64_t B, int64_t M, int64_t D, int64_t L, int64_t Q, int64_t P, int dtype) {
  return (uint64_t)elem_size(dtype) * (uint64_t)B * (uint64_t)Q * (uint64_t)M * (uint64_t)L * (uint64_t)P * 4ull * (uint64_t)D;
}

size_t msda_b200_workspace_bytes(int64_t batch, int64_t num_keys, int64_t num_heads, int64_t channels, int64_t num_levels,
                                 int64_t num_queries, int64_t num_points, int dtype) {
  return packed_workspace_bytes(batch, num_keys, num_heads, channels, num_queries, num_levels, num_points, dtype);
}

int msda_b200_forward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                      const void *sampling_loc, const void *attn_weight, void *output, int64_t batch, int64_t num_keys,
                      int64_t num_heads, int64_t channels, int64_t num_levels, int64_t num_queries, int64_t num_points,
                      int64_t im2col_step, int dtype, unsigned flags, void *stream) {
  return msda_b200_forward_ws(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, output, nullptr, 0, batch,
                              num_keys, num_heads, channels, num_levels, num_queries, num_points, im2col_step, dtype, flags,
                              stream);
}

int msda_b200_forward_ws(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                         const void *sampling_loc, const void *attn_weight, void *output, void *workspace,
                         size_t workspace_bytes, int64_t batch, int64_t num_keys, int64_t num_heads, int64_t channels,
                         int64_t num_levels, int64_t num_queries, int64_t num_points, int64_t im2col_step, int dtype,
                         unsigned flags, void *stream) {
  int rc = validate_common(value, spatial_shapes, level_start_index, output, batch, num_keys, num_heads, channels,
                           num_levels, num_queries, num_points, dtype);
  if (rc != MSDA_OK) return rc;
  if ((flags & MSDA_FLAG_MATH_FHFMA) && (flags & MSDA_FLAG_MATH_EXACT)) return MSDA_ERR_BAD_FLAGS;
  // ms_deform_attn.cu:922-926: im2col_step_ = min(batch, im2col_step); batch % im2col_step_ == 0
  if (batch > 0) {
    const int64_t step = batch < im2col_step ? batch : im2col_step;
    if (step <= 0 || batch % step != 0) return MSDA_ERR_BAD_STEP;
  }
  if (batch * num_queries * num_heads * channels == 0) return MSDA_OK;
  if (num_levels * num_points > 0 && (!sampling_loc || !attn_weight)) return MSDA_ERR_NULL_POINTER;
  const size_t E = elem_size(dtype);
  if (!aligned_to(sampling_loc, E) || !aligned_to(attn_weight, E)) return MSDA_ERR_MISALIGNED;

  MsdaParams p;
  memset(&p, 0, sizeof(p));
  p.value = value;
  p.shapes = spatial_shapes;
  p.starts = level_start_index;
  p.loc = sampling_loc;
  p.weight = attn_weight;
  p.out = output;
  p.B = (int)batch; p.S = (int)num_keys; p.M = (int)num_heads; p.D = (int)channels;
  p.L = (int)num_levels; p.Q = (int)num_queries; p.P = (int)num_points;
  p.ref_dim = 0;
  return forward_impl(p, dtype, flags, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int msda_b200_forward_fused(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                            const void *reference_points, const void *sampling_offsets, const void *attn_logits,
                            void *output, int64_t batch, int64_t num_keys, int64_t num_heads, int64_t channels,
                            int64_t num_levels, int64_t num_queries, int64_t num_points, int64_t ref_dim, int dtype,
                            unsigned flags, void *stream) {
  int rc = validate_common(value, spatial_shapes, level_start_index, output, batch, num_keys, num_heads, channels,
                           num_levels, num_queries, num_points, dtype);
  if (rc != MSDA_OK) return rc;
  if (ref_dim != 2 && ref_dim != 4) return MSDA_ERR_BAD_SHAPE;
  if (batch * num_queries * num_heads * channels == 0) return MSDA_OK;
  if (num_levels * num_points > 0 && (!reference_points || !sampling_offsets || !attn_logits)) return MSDA_ERR_NULL_POINTER;
  const size_t E = elem_size(dtype);
  if (!aligned_to(reference_points, E) || !aligned_to(sampling_offsets, E) || !aligned_to(attn_logits, E)) return MSDA_ERR_MISALIGNED;

  MsdaParams p;
  memset(&p, 0, sizeof(p));
  p.value = value;
  p.shapes = spatial_shapes;
  p.starts = level_start_index;
  p.ref = reference_points;
  p.offsets = sampling_offsets;
  p.logits = attn_logits;
  p.out = output;
  p.B = (int)batch; p.S = (int)num_keys; p.M = (int)num_heads; p.D = (int)channels;
  p.L = (int)num_levels; p.Q = (int)num_queries; p.P = (int)num_points;
  p.ref_dim = (int)ref_dim;
  return forward_impl(p, dtype, flags, nullptr, 0, static_cast<cudaStream_t>(stream));
}

static int trt_to_msda_dtype(int trt_dtype) {
  switch (trt_dtype) {  // nvinfer1::DataType values (deformable_attention_plugin.cpp:53-62 accepts kFLOAT, kHALF)
    case 0: return MSDA_F32;
    case 1: return MSDA_F16;
    case 7: return MSDA_BF16;
    default: return -1;
  }
}

size_t msda_b200_plugin_workspace_bytes(const int64_t *value_dims, const int64_t *loc_dims, int trt_dtype) {
  if (!value_dims || !loc_dims) return 0;
  const int dtype = trt_to_msda_dtype(trt_dtype);
  if (dtype < 0) return 0;
  return packed_workspace_bytes(value_dims[0], value_dims[1], value_dims[2], value_dims[3], loc_dims[1], loc_dims[3], loc_dims[4],
                                dtype);
}

int msda_b200_plugin_enqueue(const int64_t *value_dims, const int64_t *loc_dims, int trt_dtype,
                             const void *const *inputs, void *const *outputs, void *workspace, size_t workspace_bytes,
                             int64_t im2col_step, void *stream) {
  if (!value_dims || !loc_dims || !inputs || !outputs) return MSDA_ERR_NULL_POINTER;
  const int dtype = trt_to_msda_dtype(trt_dtype);
  if (dtype < 0) return MSDA_ERR_BAD_DTYPE;
  // deformable_attention_plugin.cpp:305-315
  const int64_t bs = value_dims[0], num_keys = value_dims[1], num_heads = value_dims[2], dim_per_head = value_dims[3];
  const int64_t num_queries = loc_dims[1], num_levels = loc_dims[3], num_points = loc_dims[4];
  if (loc_dims[0] != bs || loc_dims[2] != num_heads || loc_dims[5] != 2) return MSDA_ERR_BAD_SHAPE;
  return msda_b200_forward_ws(inputs[0], static_cast<const int64_t *>(inputs[1]), static_cast<const int64_t *>(inputs[2]),
                              inputs[3], inputs[4], outputs[0], workspace, workspace_bytes, bs, num_keys, num_heads,
                              dim_per_head, num_levels, num_queries, num_points, im2col_step, dtype, MSDA_FLAG_DEFAULT, stream);
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

size_t msda_b200_host_workspace_bytes(int64_t B, int64_t S, int64_t M, int64_t D, int64_t L, int64_t Q, int64_t P,
                                      int dtype) {
  const size_t E = elem_size(dtype);
  if (E == 0) return 0;
  size_t total = 0;
  total += align_up((size_t)B * S * M * D * E, 256);
  total += align_up((size_t)L * 2 * 8, 256);
  total += align_up((size_t)L * 8, 256);
  total += align_up((size_t)B * Q * M * L * P * 2 * E, 256);
  total += align_up((size_t)B * Q * M * L * P * E, 256);
  total += align_up((size_t)B * Q * M * D * E, 256);
  return total;
}

int msda_b200_forward_host(const void *value_host, const int64_t *spatial_shapes_host,
                           const int64_t *level_start_index_host, const void *sampling_loc_host,
                           const void *attn_weight_host, void *output_host, void *workspace_dev, size_t workspace_bytes,
                           int64_t B, int64_t S, int64_t M, int64_t D, int64_t L, int64_t Q, int64_t P,
                           int64_t im2col_step, int dtype, unsigned flags, void *stream_v) {
  const size_t E = elem_size(dtype);
  if (E == 0) return MSDA_ERR_BAD_DTYPE;
  if (B < 0 || S < 0 || M < 0 || D < 0 || L < 0 || Q < 0 || P < 0) return MSDA_ERR_BAD_SHAPE;
  if (!workspace_dev) return MSDA_ERR_NULL_POINTER;
  if (workspace_bytes < msda_b200_host_workspace_bytes(B, S, M, D, L, Q, P, dtype)) return MSDA_ERR_WORKSPACE_TOO_SMALL;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  char *ws = static_cast<char *>(workspace_dev);
  const size_t n_val = (size_t)B * S * M * D * E, n_shp = (size_t)L * 16, n_st = (size_t)L * 8;
  const size_t n_loc = (size_t)B * Q * M * L * P * 2 * E, n_w = (size_t)B * Q * M * L * P * E, n_out = (size_t)B * Q * M * D * E;
  char *d_val = ws; ws += align_up(n_val, 256);
  char *d_shp = ws; ws += align_up(n_shp, 256);
  char *d_st = ws; ws += align_up(n_st, 256);
  char *d_loc = ws; ws += align_up(n_loc, 256);
  char *d_w = ws; ws += align_up(n_w, 256);
  char *d_out = ws;
  cudaError_t e;
#define MSDA_COPY(dst, src, n, kind)                                   \
  if ((n) > 0) {                                                       \
    if (!(src) || !(dst)) return MSDA_ERR_NULL_POINTER;                \
    e = cudaMemcpyAsync((dst), (src), (n), (kind), stream);            \
    if (e != cudaSuccess) return (int)e;                               \
  }
  MSDA_COPY(d_val, value_host, n_val, cudaMemcpyHostToDevice)
  MSDA_COPY(d_shp, spatial_shapes_host, n_shp, cudaMemcpyHostToDevice)
  MSDA_COPY(d_st, level_start_index_host, n_st, cudaMemcpyHostToDevice)
  MSDA_COPY(d_loc, sampling_loc_host, n_loc, cudaMemcpyHostToDevice)
  MSDA_COPY(d_w, attn_weight_host, n_w, cudaMemcpyHostToDevice)
  const int rc = msda_b200_forward(d_val, reinterpret_cast<const int64_t *>(d_shp), reinterpret_cast<const int64_t *>(d_st),
                                   d_loc, d_w, d_out, B, S, M, D, L, Q, P, im2col_step, dtype, flags, stream_v);
  if (rc != 0) return rc;
  MSDA_COPY(output_host, d_out, n_out, cudaMemcpyDeviceToHost)
#undef MSDA_COPY
  return MSDA_OK;
}

int msda_b200_backward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                       const void *sampling_loc, const void *attn_weight, const void *grad_output, void *grad_value,
                       void *grad_sampling_loc, void *grad_attn_weight, int64_t batch, int64_t num_keys, int64_t num_heads,
                       int64_t channels, int64_t num_levels, int64_t num_queries, int64_t num_points, int64_t im2col_step,
                       int dtype, unsigned flags, void *stream_v) {
  const size_t E = elem_size(dtype);
  if (E == 0) return MSDA_ERR_BAD_DTYPE;
  if (batch < 0 || num_keys < 0 || num_heads < 0 || channels < 0 || num_levels < 0 || num_queries < 0 || num_points < 0)
    return MSDA_ERR_BAD_SHAPE;
  const int64_t lim = (int64_t)1 << 31;
  if (num_keys >= lim || num_queries >= lim || batch >= lim || num_heads >= 65536 || channels >= 65536 || num_levels >= 65536 ||
      num_points >= 65536)
    return MSDA_ERR_UNSUPPORTED;
  if (batch > 0) {  // ms_deform_attn.cu:999-1001
    const int64_t step = batch < im2col_step ? batch : im2col_step;
    if (step <= 0 || batch % step != 0) return MSDA_ERR_BAD_STEP;
  }
  const int64_t samples = batch * num_queries * num_heads * num_levels * num_points;
  if (samples == 0) return MSDA_OK;
  if (!value || !spatial_shapes || !level_start_index || !sampling_loc || !attn_weight || !grad_output || !grad_value ||
      !grad_sampling_loc || !grad_attn_weight)
    return MSDA_ERR_NULL_POINTER;
  const void *ptrs[] = {value, sampling_loc, attn_weight, grad_output, grad_value, grad_sampling_loc, grad_attn_weight};
  for (const void *q : ptrs)
    if (!aligned_to(q, E)) return MSDA_ERR_MISALIGNED;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  BwdParams p;
  p.value = value; p.shapes = spatial_shapes; p.starts = level_start_index; p.loc = sampling_loc; p.weight = attn_weight;
  p.grad_out = grad_output; p.grad_value = grad_value; p.grad_loc = grad_sampling_loc; p.grad_weight = grad_attn_weight;
  p.B = (int)batch; p.S = (int)num_keys; p.M = (int)num_heads; p.D = (int)channels; p.L = (int)num_levels;
  p.Q = (int)num_queries; p.P = (int)num_points;
  const int sms = sm_count();
  const bool vec = !(flags & MSDA_FLAG_FORCE_GENERIC) && dtype != MSDA_F64 &&
                   ((E == 2 && channels == 32) || (E == 4 && (channels == 16 || channels == 32))) &&
                   aligned_to(value, 16) && aligned_to(grad_output, 16) && aligned_to(grad_value, 16) &&
                   aligned_to(sampling_loc, 2 * E) && ((size_t)num_heads * channels * E) % 16 == 0;
  if (vec) {
    const int G = (int)(channels * E / 16);
    const int64_t pairs = batch * num_queries * num_heads;
    int64_t grid = (pairs * G + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)sms * 32;
    if (grid > cap) grid = cap;
    if (dtype == MSDA_F16) msda_bwd_vec<__half, 32><<<(unsigned)grid, kThreads, 0, stream>>>(p);
    else if (dtype == MSDA_BF16) msda_bwd_vec<__nv_bfloat16, 32><<<(unsigned)grid, kThreads, 0, stream>>>(p);
    else if (channels == 16) msda_bwd_vec<float, 16><<<(unsigned)grid, kThreads, 0, stream>>>(p);
    else msda_bwd_vec<float, 32><<<(unsigned)grid, kThreads, 0, stream>>>(p);
    snprintf(g_last_variant, sizeof(g_last_variant), "bwd_vec<%s,D%d>", dtype_name(dtype), (int)channels);
  } else {
    int64_t grid = (samples + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)sms * 32;
    if (grid > cap) grid = cap;
    switch (dtype) {
      case MSDA_F32: msda_bwd_generic<float><<<(unsigned)grid, kThreads, 0, stream>>>(p); break;
      case MSDA_F16: msda_bwd_generic<__half><<<(unsigned)grid, kThreads, 0, stream>>>(p); break;
      case MSDA_BF16: msda_bwd_generic<__nv_bfloat16><<<(unsigned)grid, kThreads, 0, stream>>>(p); break;
      default: msda_bwd_generic<double><<<(unsigned)grid, kThreads, 0, stream>>>(p); break;
    }
    snprintf(g_last_variant, sizeof(g_last_variant), "bwd_generic<%s>", dtype_name(dtype));
  }
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

size_t msda_b200_packed_value_bytes(int64_t batch, int64_t num_keys, int64_t num_heads, int64_t channels, int dtype) {
  if ((dtype != MSDA_F16 && dtype != MSDA_BF16) || channels != 32 || batch <= 0 || num_keys <= 0 || num_heads <= 0) return 0;
  return (size_t)batch * (size_t)num_keys * (size_t)num_heads * 128;
}

int msda_b200_pack_value(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index, void *packed,
                         int64_t batch, int64_t num_keys, int64_t num_heads, int64_t channels, int64_t num_levels, int dtype,
                         void *stream_v) {
  if (msda_b200_packed_value_bytes(batch, num_keys, num_heads, channels, dtype) == 0 || num_levels <= 0 || num_levels > kMaxLevelsSmem)
    return MSDA_ERR_UNSUPPORTED;
  if (!value || !spatial_shapes || !level_start_index || !packed) return MSDA_ERR_NULL_POINTER;
  if (!aligned_to(value, 16) || !aligned_to(packed, 32)) return MSDA_ERR_MISALIGNED;
  MsdaParams p;
  memset(&p, 0, sizeof(p));
  p.value = value; p.shapes = spatial_shapes; p.starts = level_start_index; p.packed = packed;
  p.B = (int)batch; p.S = (int)num_keys; p.M = (int)num_heads; p.D = (int)channels; p.L = (int)num_levels;
  p.tile_w_log2 = 3; p.tile_h_log2 = 2; p.want_tiled = 0; p.Q = (int)num_keys; p.P = 4;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const long long chunks = (long long)batch * num_keys * num_heads * 4;
  long long grid = (chunks + kThreads - 1) / kThreads;
  const long long cap = (long long)sm_count() * 16;
  if (grid > cap) grid = cap;
  if (dtype == MSDA_F16) msda_pack_value<__half><<<(unsigned)grid, kThreads, 0, stream>>>(p);
  else msda_pack_value<__nv_bfloat16><<<(unsigned)grid, kThreads, 0, stream>>>(p);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  const int rc = (int)cudaGetLastError();
  if (rc == 0) snprintf(g_last_variant, sizeof(g_last_variant), "pack_value<%s>", dtype_name(dtype));
  return rc;
}

int msda_b200_forward_packed(const void *packed_value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                             const void *sampling_loc, const void *attn_weight, void *output, int64_t batch, int64_t num_keys,
                             int64_t num_heads, int64_t channels, int64_t num_levels, int64_t num_queries, int64_t num_points,
                             int dtype, unsigned flags, void *stream_v) {
  if ((flags & MSDA_FLAG_MATH_FHFMA) && (flags & MSDA_FLAG_MATH_EXACT)) return MSDA_ERR_BAD_FLAGS;
  if (batch < 0 || num_keys < 0 || num_queries < 0) return MSDA_ERR_BAD_SHAPE;
  if (batch * num_queries == 0) return MSDA_OK;
  // what the packed gather is written for; anything else goes through msda_b200_forward on the plain layout
  if ((dtype != MSDA_F16 && dtype != MSDA_BF16) || channels != 32 || num_heads != 8 || num_points != 4 || num_levels < 2 ||
      num_levels > kHpMaxLevels || batch > 65535 || (int64_t)num_queries * num_heads * num_levels * 16 >= ((int64_t)1 << 32) ||
      num_keys >= ((int64_t)1 << 31) / 1024)
    return MSDA_ERR_UNSUPPORTED;
  if (!packed_value || !spatial_shapes || !level_start_index || !sampling_loc || !attn_weight || !output) return MSDA_ERR_NULL_POINTER;
  if (!aligned_to(packed_value, 32) || !aligned_to(output, 16) || !aligned_to(sampling_loc, 4) || !aligned_to(attn_weight, 2) ||
      !aligned_to(spatial_shapes, 8) || !aligned_to(level_start_index, 8))
    return MSDA_ERR_MISALIGNED;
  MsdaParams p;
  memset(&p, 0, sizeof(p));
  p.packed = const_cast<void *>(packed_value);
  p.shapes = spatial_shapes; p.starts = level_start_index; p.loc = sampling_loc; p.weight = attn_weight; p.out = output;
  p.B = (int)batch; p.S = (int)num_keys; p.M = (int)num_heads; p.D = (int)channels; p.L = (int)num_levels;
  p.Q = (int)num_queries; p.P = (int)num_points;
  int math = (dtype == MSDA_F16) ? kFhfma : kExact;
  if (flags & MSDA_FLAG_MATH_FHFMA) math = kFhfma;
  if (flags & MSDA_FLAG_MATH_EXACT) math = kExact;
  const int sms = sm_count(), NG = 4, cpg = sms / NG;
  if (cpg < 1) return MSDA_ERR_UNSUPPORTED;
  const int64_t quads = (num_queries + 3) / 4;
  int warps = env_int("MSDA_B200_HP_WARPS", hp_pick_warps(quads, cpg));
  if (warps < 1 || warps > kHpThreads / 32) warps = kHpThreads / 32;
  const bool pdl = env_int("MSDA_B200_PDL", 1) != 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const dim3 grid((unsigned)(cpg * NG), (unsigned)batch, 1), block((unsigned)warps * 32u);
  cudaError_t le;
  if (dtype == MSDA_F16) {
    le = math == kFhfma ? launch_kernel(msda_fwd_hp<__half, kFhfma, 8, false, true>, grid, block, 0, stream, pdl, p)
                        : launch_kernel(msda_fwd_hp<__half, kExact, 8, false, true>, grid, block, 0, stream, pdl, p);
  } else {
    le = math == kFhfma ? launch_kernel(msda_fwd_hp<__nv_bfloat16, kFhfma, 8, false, true>, grid, block, 0, stream, pdl, p)
                        : launch_kernel(msda_fwd_hp<__nv_bfloat16, kExact, 8, false, true>, grid, block, 0, stream, pdl, p);
  }
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  const int rc = le != cudaSuccess ? (int)le : (int)cudaGetLastError();
  if (rc == 0)
    snprintf(g_last_variant, sizeof(g_last_variant), "hp_packed<%s,D32,P4,M8>/%dwarps/%s", dtype_name(dtype), warps,
             math == kFhfma ? "fhfma" : "exact");
  return rc;
}

int msda_b200_read_probe(const void *buf, size_t bytes, int repeats, void *sink, void *stream) {
  if (!buf || !sink) return MSDA_ERR_NULL_POINTER;
  if (!aligned_to(buf, 16)) return MSDA_ERR_MISALIGNED;
  const size_t n_vec = bytes / 16;
  if (n_vec == 0 || repeats <= 0) return MSDA_OK;
  const unsigned grid = (unsigned)(sm_count() * 8);
  read_probe_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4 *>(buf), n_vec,
                                                                              repeats, static_cast<unsigned *>(sink));
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaGetLastError();
}

}  // extern "C"
