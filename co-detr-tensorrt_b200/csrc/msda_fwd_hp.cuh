// msda_fwd_hp.cuh -- "head-pair" sampling kernel: the coarse pyramid levels live in shared memory.
// Textually included by msda_sm100.cu (same translation unit: it uses that file's helpers).
//
// Why: tools/gather_probe.cu measured what one B200 SM can gather (profiles/r02_gather_probe.jsonl):
//   * 64-byte rows through LDG (any vector width, L1 hit or miss):  1.0 row / clk / SM   -> 31.5 us for the
//     9.2 M live corner rows of the 1152x768 encoder call; round 1's kernel ran at 64 % of that;
//   * the same rows through LDS.128 from shared memory:              1.94 rows / clk / SM when the two lane
//     groups of a quarter-warp sit on opposite halves of the 32 banks, 1.33 when they do not;
//   * TMA (tile::gather4 / cp.async.bulk per row):                   0.30 / 0.13 rows / clk / SM.
// So the only way under the LDG bound is to serve part of the gather from shared memory.  A whole level for
// all 8 heads does not fit (level 2 of the 1152x768 pyramid is 442 KB), but a PAIR of heads does: levels 2-4
// are 1,134 pixels x 2 heads x 64 B = 145 KB, and they receive 52 % of the live corner rows.  Hence:
//   * one CTA per SM, 1,024 threads, owns one head pair (blockIdx.x % (M/2)) of one image (blockIdx.y) and copies
//     the coarsest levels that fit into its dynamic shared memory as [pixel][2 heads][64 B] -- the natural
//     layout, in which head parity IS the bank half;
//   * a warp works on 4 consecutive queries x the 2 heads: lanes 0-3 / 4-7 of every quarter-warp hold heads
//     2k / 2k+1 of the same query, so every LDS.128 quarter is conflict-free whatever pixels are sampled;
//   * fine levels are gathered from global memory exactly like the round-1 kernel (LDG.E.128 per lane).
// Everything else follows msda_fwd_vec: one lane of a group works out the geometry of one point of the level
// and broadcasts index + packed weights by shuffle; "does not contribute" is a weight of exactly zero, which
// predicates the load and its FMAs off; fp16 multiplies with FHFMA; the first sample of the next unit is
// loaded while the current unit is computed.  Which levels are cached is decided on the device from the
// device-resident shapes (the launcher never reads them), against the shared-memory size it was launched with.
//
// Reference semantics: ms_deform_attn.cu:31-77 (bilinear helper), :218-260 (sample loop).

#ifndef MSDA_HP_THREADS
#define MSDA_HP_THREADS 800
#endif
constexpr int kHpThreads = MSDA_HP_THREADS;
#ifndef MSDA_HP_MAXNREG
#define MSDA_HP_MAXNREG 64
#endif
constexpr int kHpMaxLevels = 8;

struct HpLevel {
  int H, W;
  int base;   // first key of the level in the value tensor, or (cached level) its first 128-byte entry in shared memory
  int start;  // first key of the level in the value tensor
};
// threads per CTA of the fused-producer instantiations: softmax statistics of two units, the reference point of the
// step in flight and a second running offset need ~10 more registers than the plain kernel's 72
#ifndef MSDA_HP_FUSED_THREADS
#define MSDA_HP_FUSED_THREADS 800
#endif
constexpr int kHpFusedThreads = MSDA_HP_FUSED_THREADS;

// one 32-byte load straight into two row registers (packed pyramid: a pixel's row and its right-hand neighbour's)
__device__ __forceinline__ void ldg256_pair(uint4 &a, uint4 &b, const void *ptr) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(ptr));
}

__device__ __forceinline__ uint4 lds128(unsigned addr) {
  uint4 r;
  asm("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}

// The four samples of one level for this lane group.  Rows come from `sm_lane` (SMEM: the lane's 32-bit shared address
// inside a cached level, 2 * D * sizeof(T) bytes per pixel) or from `vm` (64-bit global address of the lane's bytes of
// its head in pixel 0, PIXB bytes per pixel).
// PACKED: `vm` points into the pixel-pair packed pyramid (128-byte (pixel, head) entries that also hold the right-hand
// neighbour, chunk-interleaved; PIXB = M * 128): the two corners of an image row arrive with ONE 32-byte load per lane
// (LDG.E.256), one L1 wavefront per lane group instead of two.
// MATH = kFhfmaSplit (bf16): FHFMA with every corner weight as TWO bf16 terms, w = hi + lo (|w - hi - lo| <= 2^-17 |w|):
// the 8-bit bf16 weight alone costs 2^-9 per product, past the one-output-rounding gate, while the exact path spends
// 12 instructions per corner row (8 integer unpacks + 4 FFMA2) and four fp32 weight registers per sample.
constexpr int kFhfmaSplit = 2;

template <typename T, int MATH, bool SMEM, int PIXB, bool PACKED = false>
__device__ __forceinline__ void hp_level_samples(float (&acc)[16 / sizeof(T)], int i00, unsigned pk0, unsigned pk1, unsigned pk2, unsigned pk3,
                                                 const float (&cw)[4], int W, const char *vm, unsigned sm_lane) {
  constexpr bool kPackedW = MATH == kFhfma || MATH == kFhfmaSplit;  // corner weights travel as packed 16-bit pairs
  constexpr unsigned group_mask = 0xffffffffu;
  constexpr int G = 32 * (int)sizeof(T) / 16;   // lanes per corner row: 4 (16-bit), 8 (fp32: a row is a whole 128-byte line)
  constexpr unsigned SPIX = 64u * (unsigned)sizeof(T);  // bytes per pixel of a cached level: two heads x 32 channels
  // Row loads are written MSDA_HP_DEPTH samples (x 4 corner rows) ahead of the FMAs that consume them.  ptxas keeps the
  // order it is given as far as registers allow: depth 1 leaves 4 row loads in flight per lane whatever the register
  // budget; depth 2 / 4 hold 8 / 16 (32 / 64 registers of rows) and need the smaller CTAs of MSDA_HP_THREADS.
#ifndef MSDA_HP_DEPTH
#define MSDA_HP_DEPTH 1
#endif
  struct Slot {
    uint4 r0, r1, r2, r3;
    unsigned p0, p1, p2, p3;
    float w0, w1, w2, w3;
    bool o0, o1, o2, o3;  // corner inside the level and weighted: its row is loaded and accumulated
  };
  auto corner_on = [](const Slot &s, int j) -> bool {
    if constexpr (kPackedW) return ((((j & 2) ? s.p1 : s.p0) >> ((j & 1) * 16)) & 0x7fffu) != 0u;
    else return (j == 0 ? s.w0 : j == 1 ? s.w1 : j == 2 ? s.w2 : s.w3) != 0.f;
  };
  auto issue = [&](Slot &s, int k) {
    const int bi = __shfl_sync(group_mask, i00, k, G);
    s.p0 = s.p1 = s.p2 = s.p3 = 0u;
    s.w0 = s.w1 = s.w2 = s.w3 = 0.f;
    if constexpr (kPackedW) {
      s.p0 = __shfl_sync(group_mask, pk0, k, G);
      s.p1 = __shfl_sync(group_mask, pk1, k, G);
      if constexpr (MATH == kFhfmaSplit) {
        s.p2 = __shfl_sync(group_mask, pk2, k, G);
        s.p3 = __shfl_sync(group_mask, pk3, k, G);
      }
    } else {
      s.w0 = __shfl_sync(group_mask, cw[0], k, G);
      s.w1 = __shfl_sync(group_mask, cw[1], k, G);
      s.w2 = __shfl_sync(group_mask, cw[2], k, G);
      s.w3 = __shfl_sync(group_mask, cw[3], k, G);
    }
    s.o0 = corner_on(s, 0);
    s.o1 = corner_on(s, 1);
    s.o2 = corner_on(s, 2);
    s.o3 = corner_on(s, 3);
    const bool o0 = s.o0, o1 = s.o1, o2 = s.o2, o3 = s.o3;
    // two row addresses per sample (top-left, bottom-left); the right-hand corners are immediate offsets
    if constexpr (PACKED) {
      const char *g0 = vm + (ptrdiff_t)bi * (ptrdiff_t)PIXB;
      const char *g1 = vm + (ptrdiff_t)(bi + W) * (ptrdiff_t)PIXB;
      if (o0 || o1) ldg256_pair(s.r0, s.r1, g0);
      if (o2 || o3) ldg256_pair(s.r2, s.r3, g1);
    } else if constexpr (SMEM) {
      const unsigned s0 = sm_lane + (unsigned)bi * SPIX, s1 = s0 + (unsigned)W * SPIX;
      if (o0) s.r0 = lds128(s0);
      if (o1) s.r1 = lds128(s0 + SPIX);
      if (o2) s.r2 = lds128(s1);
      if (o3) s.r3 = lds128(s1 + SPIX);
    } else {
      // signed: the top-left index is -1 (or -1 - W) when only right-hand / lower corners are inside the level
      const char *g0 = vm + (ptrdiff_t)bi * (ptrdiff_t)PIXB;
      const char *g1 = vm + (ptrdiff_t)(bi + W) * (ptrdiff_t)PIXB;
      if (o0) s.r0 = ldg128(g0);
      if (o1) s.r1 = ldg128(g0 + PIXB);
      if (o2) s.r2 = ldg128(g1);
      if (o3) s.r3 = ldg128(g1 + PIXB);
    }
  };
  auto corner = [&](const Slot &s, int j, bool on, const uint4 &row, float w) {
    if constexpr (kPackedW) {
      const unsigned w16 = (((j & 2) ? s.p1 : s.p0) >> ((j & 1) * 16)) & 0xffffu;
      if (on) RowFma<T, kFhfma>::run(acc, row, 0.f, w16);
      if constexpr (MATH == kFhfmaSplit) {
        // keep the two 8-FMA groups apart: merged into one 16-instruction conditional block they are compiled as a
        // branch instead of predicated FMAs (57.4 instead of 50.6 us at the headline shape)
        asm volatile("");
        const unsigned l16 = (((j & 2) ? s.p3 : s.p2) >> ((j & 1) * 16)) & 0xffffu;
        if (on) RowFma<T, kFhfma>::run(acc, row, 0.f, l16);
      }
    } else {
      if (on) RowFma<T, kExact>::run(acc, row, w, 0u);
    }
  };
  auto consume = [&](const Slot &s) {
    corner(s, 0, s.o0, s.r0, s.w0);
    corner(s, 1, s.o1, s.r1, s.w1);
    corner(s, 2, s.o2, s.r2, s.w2);
    corner(s, 3, s.o3, s.r3, s.w3);
  };
  constexpr int DEPTH = (SMEM || PACKED) ? 1 : MSDA_HP_DEPTH;
  static_assert(DEPTH == 1 || DEPTH == 2 || DEPTH == 4, "MSDA_HP_DEPTH: 1, 2 or 4 samples of row loads ahead");
  if constexpr (DEPTH == 1) {
    Slot a;
    issue(a, 0); consume(a);
    issue(a, 1); consume(a);
    issue(a, 2); consume(a);
    issue(a, 3); consume(a);
  } else if constexpr (DEPTH == 2) {
    Slot a, b;
    issue(a, 0);
    issue(b, 1);
    consume(a);
    issue(a, 2);
    consume(b);
    issue(b, 3);
    consume(a);
    consume(b);
  } else {
    Slot a, b, c, d;
    issue(a, 0);
    issue(b, 1);
    issue(c, 2);
    issue(d, 3);
    consume(a);
    consume(b);
    consume(c);
    consume(d);
  }
}

// Geometry for the packed pyramid: like make_geo(), but the index names the ENTRY that holds the sample's two upper
// corners.  Column -1 (only the right-hand corners are inside): the entry of column 0 is used and its LEFT half is the
// sample's right-hand corner, so the weight pair is (lw, 0) instead of (0, lw) -- the same products in the same order.
__device__ __forceinline__ void make_geo_packed(float x, float y, float aw, int H, int W, int &e_top, float (&cw)[4]) {
  const float w_im = __fmul_rn(x, (float)W) - 0.5f;
  const float h_im = __fmul_rn(y, (float)H) - 0.5f;
  const bool inside = (h_im > -1.f) && (w_im > -1.f) && (h_im < (float)H) && (w_im < (float)W);
  const float hf = floorf(h_im), wf = floorf(w_im);
  const int h_lo = (int)hf, w_lo = (int)wf;
  const float lh = h_im - hf, lw = w_im - wf;
  const float hh = 1.f - lh, hw = 1.f - lw;
  const float wy0 = (inside && h_lo >= 0) ? hh * aw : 0.f;
  const float wy1 = (inside && h_lo < H - 1) ? lh * aw : 0.f;
  const bool neg = w_lo < 0;
  const float wl = inside ? (neg ? lw : hw) : 0.f;
  const float wr = (inside && !neg && w_lo < W - 1) ? lw : 0.f;
  cw[0] = wy0 * wl;
  cw[1] = wy0 * wr;
  cw[2] = wy1 * wl;
  cw[3] = wy1 * wr;
  e_top = h_lo * W + (neg ? 0 : w_lo);
}

// T: __half / __nv_bfloat16, D = 32, P = 4, MT heads (compile time: the neighbour-pixel offset is an immediate).
// DYN: warps draw (query quad, head pair) units from the launch's device counter instead of striding over their own
// head pair's quads (no cached levels in that mode: a CTA is no longer tied to one head pair).
// PACKED: `p.packed` holds the pixel-pair packed pyramid (written by msda_pack_value or by the projection kernel's
// epilogue); `p.value` is not read.  No cached levels in that mode.
// FUSED: `p.offsets` / `p.logits` / `p.ref` in place of locations and weights (msda_b200_forward_fused): the softmax over
// a pair's L*4 logits and the location arithmetic of multi_scale_deformable_attention.py:180-200 run in the geometry
// lanes, rounding where the unfused pipeline rounds -- the same helpers, in the same order, as the vector kernel's
// fused mode, so the two are bit-identical.
template <typename T, int MATH, int MT, bool DYN = false, bool PACKED = false, bool FUSED = false>
__global__ void __launch_bounds__(FUSED ? kHpFusedThreads : kHpThreads, 1) msda_fwd_hp(const MsdaParams p) {
  constexpr int D = 32, E = (int)sizeof(T), VEC = 16 / E;
  constexpr int G = D * E / 16;       // lanes per (query, head) pair = 16-byte pieces of a corner row: 4 (16-bit), 8 (fp32)
  constexpr int QPW = 32 / G / 2;     // queries per warp: 4 (16-bit), 2 (fp32)
  constexpr int SPIX = 2 * D * E;     // bytes per pixel of a cached level (the head pair): 128 / 256
  static_assert(!PACKED || E == 2, "the packed pyramid is a 16-bit layout");
  static_assert(!FUSED || (!DYN && !PACKED), "fused producers: static schedule, plain value tensor");
  extern __shared__ __align__(128) unsigned char hp_rows[];  // cached levels: [pixel][2 heads][D * E bytes]
  __shared__ HpLevel lv[kHpMaxLevels];
  __shared__ int s_first_cached;
  __shared__ float2 lv_rcp[FUSED ? kHpMaxLevels : 1];  // correctly rounded (1 / W, 1 / H): the fused producers' offset normalisation

  pdl_launch_dependents();
  pdl_wait_prior_grid();
  // cold call: the CTAs of an image ask L2 for its whole pyramid up front (one bulk prefetch each, see prefetch_value_l2)
  if constexpr (!PACKED) prefetch_value_l2(p, E);

  constexpr int M = MT;
  const int NG = M >> 1;                      // head pairs
  const int hg = (int)blockIdx.x % NG;        // this CTA's head pair
  const int rank = (int)blockIdx.x / NG;      // this CTA among those of the head pair
  const int cpg = (int)gridDim.x / NG;        // CTAs per head pair (the host launches a multiple of NG)
  const int b = blockIdx.y;
  const unsigned pix_bytes = (unsigned)(M * D * E) * (PACKED ? 2u : 1u);  // packed entries are 128 bytes per head
  const char *__restrict__ value = static_cast<const char *>(PACKED ? p.packed : p.value);
  const T *__restrict__ loc = static_cast<const T *>(FUSED ? p.offsets : p.loc);
  const T *__restrict__ wgt = static_cast<const T *>(FUSED ? p.logits : p.weight);
  T *__restrict__ out = static_cast<T *>(p.out);

  // ---- level table; which levels fit into the shared memory this launch was given ----
  if (threadIdx.x < p.L) {
    const int l = threadIdx.x;
    lv[l].H = (int)__ldg(p.shapes + 2 * l);
    lv[l].W = (int)__ldg(p.shapes + 2 * l + 1);
    lv[l].start = (int)__ldg(p.starts + l);
    lv[l].base = lv[l].start;
    if constexpr (FUSED) {
      lv_rcp[l] = make_float2(__frcp_rn((float)(lv[l].W > 0 ? lv[l].W : 1)), __frcp_rn((float)(lv[l].H > 0 ? lv[l].H : 1)));
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long used = 0;
    int l0 = p.L;
    for (int l = p.L - 1; l >= 0; --l) {
      const long long n = (long long)lv[l].H * lv[l].W;
      const bool ok = !PACKED && lv[l].H > 0 && lv[l].W > 0 && lv[l].start >= 0 && (long long)lv[l].start + n <= (long long)p.S &&
                      used + n * SPIX <= (long long)p.hp_smem_bytes;
      if (!ok) break;
      used += n * SPIX;
      l0 = l;
    }
    int off = 0;
    for (int l = l0; l < p.L; ++l) {
      lv[l].base = off;
      off += lv[l].H * lv[l].W;
    }
    s_first_cached = l0;
  }
  __syncthreads();
  const int l0 = s_first_cached;

  // ---- copy the cached levels: SPIX contiguous bytes (the head pair) per pixel ----
  for (int l = l0; l < p.L; ++l) {
    constexpr int CH = SPIX / 16;
    const int n8 = lv[l].H * lv[l].W * CH;
    const char *src = value + ((size_t)b * p.S + (size_t)lv[l].start) * pix_bytes + (size_t)hg * SPIX;
    unsigned char *dst = hp_rows + (size_t)lv[l].base * SPIX;
#pragma unroll 4
    for (int i = threadIdx.x; i < n8; i += (int)blockDim.x) {
      const int pix = i / CH, c = i % CH;
      *reinterpret_cast<uint4 *>(dst + (size_t)pix * SPIX + c * 16) = ldg128(src + (size_t)pix * pix_bytes + c * 16);
    }
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = lane / G, sub = lane % G;
  const int qi = grp >> 1, hh = grp & 1;  // query of the warp's QPW, head of the pair
  const int ks = sub & 3;                 // the sampling point this lane prepares (fp32: lanes 4..7 repeat 0..3)
  const int m = hg * 2 + hh;
  const int LP = p.L * 4;
  const int units = (p.Q + QPW - 1) / QPW;
  const int nw = (int)blockDim.x >> 5;  // warps per CTA: chosen by the host so that the units divide evenly
  const int stride = cpg * nw;
  constexpr int kHeadBytes = D * E * (PACKED ? 2 : 1), kLaneBytes = PACKED ? 32 : 16;
  const char *vm_img = value + (size_t)b * p.S * M * (size_t)kHeadBytes + (size_t)sub * kLaneBytes;  // head 0 of this image, this lane's bytes
  const unsigned sm_lane = smem_u32(hp_rows) + (unsigned)(hh * (D * E) + sub * 16);

  // Per-image bases are uniform; inside an image this lane's next sample is addressed by ONE running 32-bit byte
  // offset into the locations (its weight sits at half that offset: both arrays are [Q, M, L*4] with 4- and 2-byte
  // entries).  The host takes this kernel only when an image's locations are below 4 GB.
  const char *loc_b = reinterpret_cast<const char *>(loc) + (size_t)b * p.Q * M * LP * (2 * E);
  const char *wgt_b = reinterpret_cast<const char *>(wgt) + (size_t)b * p.Q * M * LP * E;
  // A unit index names a query quad and a head pair.  Static schedule: the CTA's own pair, index = quad.  Dynamic:
  // index = quad * NG + pair, so warps that draw consecutive indices read adjacent pieces of one query's inputs.
  auto unit_quad = [&](int idx) { return DYN ? idx / NG : idx; };
  auto unit_head = [&](int idx) { return (DYN ? idx % NG : hg) * 2 + hh; };
  const int total = DYN ? units * NG : units;
  auto is_live = [&](int idx) { return idx < total && QPW * unit_quad(idx) + qi < p.Q; };
  // offset of point `sub` of level 0 of the pair that unit idx gives this lane group; padding slots (tail of the
  // last quad, units past the end) read pair 0 of the image and contribute / store nothing
  auto unit_offset = [&](int idx) -> unsigned {
    return (is_live(idx) ? (unsigned)((QPW * unit_quad(idx) + qi) * M + unit_head(idx)) * (unsigned)(LP * 2 * E) : 0u) + (unsigned)(ks * 2 * E);
  };
  constexpr unsigned kLevelStep = 4u * 2u * (unsigned)E;  // bytes of locations per (pair, level): four points
  // fused producers: the query's reference point of the step's level travels with the step's inputs (one 32-bit
  // running offset of its own: `ref` is [Q, L, ref_dim]), the softmax statistics of a pair are worked out one unit ahead
  struct RefRaw {
    unsigned r[4];  // 16-bit: r[0] = (x, y), r[1] = (w, h); fp32: x, y, w, h
  };
  struct Stats {
    float mx, inv;
  };
  const char *ref_b = FUSED ? static_cast<const char *>(p.ref) + (size_t)b * p.Q * p.L * p.ref_dim * E : nullptr;
  const unsigned ref_step = FUSED ? (unsigned)(p.ref_dim * E) : 0u;
  auto ref_offset = [&](int idx) -> unsigned {
    return is_live(idx) ? (unsigned)((QPW * unit_quad(idx) + qi) * p.L) * ref_step : 0u;
  };
  auto load_ref = [&](unsigned o) -> RefRaw {
    RefRaw r;
    r.r[0] = r.r[1] = r.r[2] = r.r[3] = 0u;
    if constexpr (FUSED) {
      const unsigned *q = reinterpret_cast<const unsigned *>(ref_b + (size_t)o);
      if constexpr (E == 2) {
        r.r[0] = __ldg(q);
        if (p.ref_dim == 4) r.r[1] = __ldg(q + 1);
      } else {
        r.r[0] = __ldg(q);
        r.r[1] = __ldg(q + 1);
        if (p.ref_dim == 4) {
          r.r[2] = __ldg(q + 2);
          r.r[3] = __ldg(q + 3);
        }
      }
    }
    return r;
  };
  // maximum and 1 / sum of exp over the pair's L*4 logits: lane ks holds point ks of every level, the four point
  // lanes complete both with two butterflies (same order as the vector kernel's fused mode)
  auto softmax_stats = [&](int idx) -> Stats {
    Stats st{0.f, 1.f};
    if constexpr (FUSED) {
      const T *lg = reinterpret_cast<const T *>(wgt_b + (size_t)(unit_offset(idx) >> 1));
      float mx = -INFINITY;
      for (int l = 0; l < p.L; ++l) mx = fmaxf(mx, Elem<T>::to_acc(lg[l * 4]));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      float sum = 0.f;
      for (int l = 0; l < p.L; ++l) sum += fused_exp<T>(Elem<T>::to_acc(lg[l * 4]) - mx);
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      st.mx = mx;
      st.inv = 1.f / sum;
    }
    return st;
  };
  auto load_sample = [&](unsigned o) -> RawSample {
    RawSample r;
    if constexpr (E == 2) {
      r.a = ld_stream_u32(loc_b + (size_t)o);
      r.b = 0u;
      r.w = (unsigned)ld_stream_u16(wgt_b + (size_t)(o >> 1));
    } else {
      const float2 xy = __ldg(reinterpret_cast<const float2 *>(loc_b + (size_t)o));
      r.a = __float_as_uint(xy.x);
      r.b = __float_as_uint(xy.y);
      r.w = __float_as_uint(__ldg(reinterpret_cast<const float *>(wgt_b + (size_t)(o >> 1))));
    }
    return r;
  };

  // Software pipeline over (unit, level) steps, two deep: while the four samples of step t are gathered and
  // accumulated, the geometry of step t+1 (this lane's point of the next level, or of level 0 of the next
  // unit) is worked out from inputs loaded during step t-1, and the inputs of step t+2 are requested.  The
  // geometry chain (un-normalise, floor, weights, pack: ~15 dependent instructions) and the location / weight
  // loads therefore never sit between a warp and its next row loads.  Needs L >= 2 (host-checked).
  struct Geo {
    int i00, W;
    unsigned pk0, pk1, pk2, pk3;
    float cw[4];
  };
  auto geometry = [&](const RawSample &rw, const RefRaw &rr, const Stats &st, int l, bool lv_live) -> Geo {
    Geo g;
    const int H = lv[l].H;
    g.W = lv[l].W;
    float x, y, aw;
    decode_raw<T>(rw, x, y, aw);
    if constexpr (FUSED) {
      // x, y are the raw offsets, aw the logit
      T rf[4];
      if constexpr (E == 2) {
        const unsigned short h[4] = {(unsigned short)(rr.r[0] & 0xffffu), (unsigned short)(rr.r[0] >> 16), (unsigned short)(rr.r[1] & 0xffffu),
                                     (unsigned short)(rr.r[1] >> 16)};
#pragma unroll
        for (int i = 0; i < 4; ++i) rf[i] = *reinterpret_cast<const T *>(&h[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) rf[i] = __uint_as_float(rr.r[i]);
      }
      const float ox = x, oy = y;
      fused_location_fast<T>(rf, p.ref_dim, ox, oy, (float)g.W, (float)H, lv_rcp[l].x, lv_rcp[l].y, x, y);
      aw = round_like<T, float>(fused_exp<T>(aw - st.mx) * st.inv);
    }
    aw = lv_live ? aw : 0.f;
    if constexpr (PACKED) make_geo_packed(x, y, aw, H, g.W, g.i00, g.cw);
    else make_geo(x, y, aw, H, g.W, g.i00, g.cw);
    g.i00 += lv[l].base;
    g.pk0 = g.pk1 = g.pk2 = g.pk3 = 0u;
    if constexpr (MATH == kFhfma || MATH == kFhfmaSplit) {
      g.pk0 = pack_weights<T>(g.cw[0], g.cw[1]);
      g.pk1 = pack_weights<T>(g.cw[2], g.cw[3]);
    }
    if constexpr (MATH == kFhfmaSplit) {  // residuals of the rounding above (exact in fp32: Sterbenz-like, same binade)
      const float2 h0 = unpack2<T>(g.pk0), h1 = unpack2<T>(g.pk1);
      g.pk2 = pack_weights<T>(g.cw[0] - h0.x, g.cw[1] - h0.y);
      g.pk3 = pack_weights<T>(g.cw[2] - h1.x, g.cw[3] - h1.y);
    }
    return g;
  };
  // ---- unit schedule ----
  // Static: this warp's quads rank*nw + warp, + stride, ... with a trip count that is uniform by construction (that
  // of warp 0 of the CTA).  ptxas cannot prove that a bound depending on threadIdx.x >> 5 is warp-uniform, and with a
  // possibly divergent loop around the shuffles it emitted divergence fall-backs plus a register copy per predicated
  // load / FMA (46.9 M instead of 21 M instructions per call); a warp whose last unit lies past the end runs it dead.
  // Dynamic: indices drawn two ahead from the launch's counter (lane 0 draws, the warp shares the value); the loop
  // condition is a warp vote, which ptxas does treat as uniform.
  unsigned *const sched = DYN ? p.sched + (size_t)blockIdx.y * 2 : nullptr;
  auto draw = [&]() -> int {
    unsigned v = 0;
    if (lane == 0) v = atomicAdd(sched, 1u);
    return (int)__shfl_sync(0xffffffffu, v, 0);
  };
  const int first = rank * nw;
  const int iters = first < units ? (units - first + stride - 1) / stride : 0;
  int cur, nxt;
  if constexpr (DYN) {
    cur = draw();
    nxt = draw();
  } else {
    cur = first + warp;
    nxt = cur + stride;
  }
  const char *vm = vm_img + (size_t)unit_head(cur) * (size_t)kHeadBytes;
  asm volatile("" : "+l"(vm));  // one opaque 64-bit base: every corner address is a single IMAD.WIDE

  unsigned off = unit_offset(cur);          // offset of the inputs held in `raw`
  unsigned roff = ref_offset(cur);
  bool live = is_live(cur);
  Stats st_c = softmax_stats(cur);
  Geo geo = geometry(load_sample(off), load_ref(roff), st_c, 0, live);  // step 0
  off += kLevelStep;
  roff += ref_step;
  RawSample raw = load_sample(off);         // inputs of step 1 (level 1 of the first unit)
  RefRaw rraw = load_ref(roff);

  int it = 0;
#pragma unroll 1
  while (DYN ? __any_sync(0xffffffffu, cur < total) : it < iters) {
    int after = 0;
    if constexpr (DYN) after = draw();      // the unit after next: the atomic's latency hides behind this unit
    else after = nxt + stride;
    const bool live_n = is_live(nxt);
    const Stats st_n = softmax_stats(nxt);  // used by the last step of this unit (geometry of the next unit's level 0)
    float acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f;

    auto step = [&](int l, auto smem_tag) {
      constexpr bool kSmem = decltype(smem_tag)::value;
      // geometry of step t+1 from the inputs in `raw`; request the inputs of step t+2
      const bool wrap = l + 1 >= p.L;
      const Geo next = geometry(raw, rraw, wrap ? st_n : st_c, wrap ? 0 : l + 1, wrap ? live_n : live);
      off = (l + 2 == p.L) ? unit_offset(nxt) : off + kLevelStep;
      raw = load_sample(off);
      if constexpr (FUSED) {
        roff = (l + 2 == p.L) ? ref_offset(nxt) : roff + ref_step;
        rraw = load_ref(roff);
      }
      hp_level_samples<T, MATH, kSmem, MT * D * E * (PACKED ? 2 : 1), PACKED>(acc, geo.i00, geo.pk0, geo.pk1, geo.pk2, geo.pk3, geo.cw, geo.W, vm, sm_lane);
      geo = next;
    };
    // fine levels from global memory, then the cached coarse levels from shared memory
#pragma unroll 1
    for (int l = 0; l < l0; ++l) step(l, std::false_type{});
#pragma unroll 1
    for (int l = l0; l < p.L; ++l) step(l, std::true_type{});

    if (live) store_row<T, VEC>(out + ((size_t)b * p.Q * M + (size_t)(QPW * unit_quad(cur) + qi) * M + unit_head(cur)) * D + sub * VEC, acc);
    live = live_n;
    st_c = st_n;
    cur = nxt;
    nxt = after;
    if constexpr (DYN) {
      vm = vm_img + (size_t)unit_head(cur) * (size_t)kHeadBytes;
      asm volatile("" : "+l"(vm));
    }
    ++it;
  }
  if constexpr (DYN) {
    // this warp has drawn its last unit; the last warp of the image's grid row to get here re-arms the counters
    if (lane == 0) {
      const unsigned finished = atomicAdd(sched + 1, 1u);
      if (finished == gridDim.x * (unsigned)nw - 1u) {
        sched[0] = 0u;
        sched[1] = 0u;
        __threadfence();
      }
    }
  }
}
