/*
 * msda_oracle.c -- CPU restatement of the reference's multi-scale deformable
 * attention forward, in plain C.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (the package
 * co-detr-tensorrt_b200/, include/, the C-ABI library) may link, import or call
 * this file.  It is used by tests/, by __graft_entry__.smoke() as the checker,
 * and by bench.py's cpu_baseline / --impl reference legs.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against
 * the .npz fixtures under tests/golden/, which were produced in the build container by running the
 * reference's own multi_scale_deformable_attention_pytorch
 * (/root/reference/codetr/ops.py:129-186) -- see tests/golden/make_golden.py.
 *
 * What it follows (all citations relative to /root/reference):
 *   - one bilinear sample of one channel, with per-corner zero padding:
 *       codetr/csrc/ms_deform_attn.cu:31-77  (ms_deform_attn_im2col_bilinear)
 *   - the per-output-element loop over levels and points, the
 *     "x*W - 0.5 / y*H - 0.5" un-normalisation and the whole-sample range test:
 *       codetr/csrc/ms_deform_attn.cu:218-260 (ms_deformable_im2col_gpu_kernel)
 *   - the tensor layouts (value [B,S,M,D], loc [B,Q,M,L,P,2] as (x,y),
 *     weight [B,Q,M,L,P], out [B,Q,M*D]; shapes [L,2] as (H,W); starts [L]):
 *       codetr/csrc/ms_deform_attn.cu:899-956 and codetr/ops.py:129-151
 *
 * The arithmetic is carried out in the element type (float or double), in the
 * same order as the reference kernel: the four corner products are summed
 * first, then multiplied by the attention weight and added to the running
 * total.  This is a restatement, not a copy: the loops are organised per
 * (image, query, head) with the channel loop innermost so that the sample
 * geometry is computed once per sample rather than once per channel.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define MSDA_ORACLE_VERSION 1

int msda_oracle_version(void) { return MSDA_ORACLE_VERSION; }

int msda_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void msda_oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* The body is written once and instantiated for float and double. */
#define DEFINE_ORACLE(NAME, T, FLOORFN)                                                          \
  void NAME(const T *value, const int64_t *shapes, const int64_t *starts, const T *loc,          \
            const T *weight, T *out, int64_t B, int64_t S, int64_t M, int64_t D, int64_t L,      \
            int64_t Q, int64_t P) {                                                              \
    const int64_t pix_stride = M * D; /* ms_deform_attn.cu:44 (w_stride) */                      \
    _Pragma("omp parallel for collapse(2) schedule(static)")                                     \
    for (int64_t b = 0; b < B; ++b) {                                                            \
      for (int64_t q = 0; q < Q; ++q) {                                                          \
        for (int64_t m = 0; m < M; ++m) {                                                        \
          const int64_t qm = (b * Q + q) * M + m;                                                \
          const T *loc_qm = loc + qm * L * P * 2;                                                \
          const T *w_qm = weight + qm * L * P;                                                   \
          T *out_qm = out + qm * D;                                                              \
          for (int64_t c = 0; c < D; ++c) out_qm[c] = (T)0;                                      \
          for (int64_t l = 0; l < L; ++l) {                                                      \
            const int H = (int)shapes[2 * l];      /* ms_deform_attn.cu:237-239 */               \
            const int W = (int)shapes[2 * l + 1];                                                \
            const T *base = value + (b * S + starts[l]) * pix_stride + m * D;                    \
            for (int64_t p = 0; p < P; ++p) {                                                    \
              const T x = loc_qm[(l * P + p) * 2];                                               \
              const T y = loc_qm[(l * P + p) * 2 + 1];                                           \
              const T aw = w_qm[l * P + p];                                                      \
              const T h_im = y * (T)H - (T)0.5; /* ms_deform_attn.cu:246-247 */                  \
              const T w_im = x * (T)W - (T)0.5;                                                  \
              if (!(h_im > (T)-1 && w_im > (T)-1 && h_im < (T)H && w_im < (T)W)) continue;       \
              const int h_lo = (int)FLOORFN(h_im); /* ms_deform_attn.cu:35-42 */                 \
              const int w_lo = (int)FLOORFN(w_im);                                               \
              const int h_hi = h_lo + 1, w_hi = w_lo + 1;                                        \
              const T lh = h_im - (T)h_lo, lw = w_im - (T)w_lo;                                  \
              const T hh = (T)1 - lh, hw = (T)1 - lw;                                            \
              const int ok1 = (h_lo >= 0 && w_lo >= 0);          /* :53-71 */                    \
              const int ok2 = (h_lo >= 0 && w_hi <= W - 1);                                      \
              const int ok3 = (h_hi <= H - 1 && w_lo >= 0);                                      \
              const int ok4 = (h_hi <= H - 1 && w_hi <= W - 1);                                  \
              const T *p1 = base + ((int64_t)h_lo * W + w_lo) * pix_stride;                      \
              const T *p2 = base + ((int64_t)h_lo * W + w_hi) * pix_stride;                      \
              const T *p3 = base + ((int64_t)h_hi * W + w_lo) * pix_stride;                      \
              const T *p4 = base + ((int64_t)h_hi * W + w_hi) * pix_stride;                      \
              const T w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw; /* :73 */          \
              for (int64_t c = 0; c < D; ++c) {                                                  \
                const T v1 = ok1 ? p1[c] : (T)0;                                                 \
                const T v2 = ok2 ? p2[c] : (T)0;                                                 \
                const T v3 = ok3 ? p3[c] : (T)0;                                                 \
                const T v4 = ok4 ? p4[c] : (T)0;                                                 \
                const T val = (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4); /* :75 */                 \
                out_qm[c] += val * aw;                                 /* :250-252 */            \
              }                                                                                  \
            }                                                                                    \
          }                                                                                      \
        }                                                                                        \
      }                                                                                          \
    }                                                                                            \
  }

DEFINE_ORACLE(msda_oracle_forward_f32, float, floorf)
DEFINE_ORACLE(msda_oracle_forward_f64, double, floor)

/*
 * Producer-fused variant used as the checker for the opt-in fused entry point
 * (SURVEY.md section 8(f).1): softmax over the L*P logits of each (query, head)
 * and the sampling-location arithmetic of the calling module
 * (codetr/multi_scale_deformable_attention.py:180-200), followed by the
 * forward above.  ref_dim is 2 (encoder: loc = ref + off / (W,H)) or
 * 4 (decoder: loc = ref_xy + off / P * ref_wh * 0.5).
 * Scratch buffers loc_tmp [B,Q,M,L,P,2] and w_tmp [B,Q,M,L,P] are caller owned.
 */
#define DEFINE_PRODUCERS(NAME, T, EXPFN)                                                         \
  void NAME(const int64_t *shapes, const T *ref, const T *offsets, const T *logits, T *loc_tmp,  \
            T *w_tmp, int64_t B, int64_t M, int64_t L, int64_t Q, int64_t P, int64_t ref_dim) {  \
    _Pragma("omp parallel for collapse(2) schedule(static)")                                     \
    for (int64_t b = 0; b < B; ++b) {                                                            \
      for (int64_t q = 0; q < Q; ++q) {                                                          \
        for (int64_t m = 0; m < M; ++m) {                                                        \
          const int64_t qm = (b * Q + q) * M + m;                                                \
          const T *lg = logits + qm * L * P;                                                     \
          T *wo = w_tmp + qm * L * P;                                                            \
          T mx = lg[0];                                                                          \
          for (int64_t i = 1; i < L * P; ++i) mx = lg[i] > mx ? lg[i] : mx;                      \
          T sum = (T)0;                                                                          \
          for (int64_t i = 0; i < L * P; ++i) { wo[i] = EXPFN(lg[i] - mx); sum += wo[i]; }       \
          for (int64_t i = 0; i < L * P; ++i) wo[i] = wo[i] / sum;                               \
          for (int64_t l = 0; l < L; ++l) {                                                      \
            const T *r = ref + ((b * Q + q) * L + l) * ref_dim;                                  \
            const T Hn = (T)shapes[2 * l], Wn = (T)shapes[2 * l + 1];                            \
            for (int64_t p = 0; p < P; ++p) {                                                    \
              const T ox = offsets[(qm * L * P + l * P + p) * 2];                                \
              const T oy = offsets[(qm * L * P + l * P + p) * 2 + 1];                            \
              T *dst = loc_tmp + (qm * L * P + l * P + p) * 2;                                   \
              if (ref_dim == 2) {                                                                \
                dst[0] = r[0] + ox / Wn;                                                         \
                dst[1] = r[1] + oy / Hn;                                                         \
              } else {                                                                           \
                dst[0] = r[0] + ox / (T)P * r[2] * (T)0.5;                                       \
                dst[1] = r[1] + oy / (T)P * r[3] * (T)0.5;                                       \
              }                                                                                  \
            }                                                                                    \
          }                                                                                      \
        }                                                                                        \
      }                                                                                          \
    }                                                                                            \
  }

DEFINE_PRODUCERS(msda_oracle_producers_f32, float, expf)
DEFINE_PRODUCERS(msda_oracle_producers_f64, double, exp)

/*
 * Backward of the forward above: gradients w.r.t. value (scatter-add), the sampling locations and the
 * attention weights, given grad_out [B,Q,M*D].  Restates the reference's per-sample derivative
 *   codetr/csrc/ms_deform_attn.cu:79-141   (ms_deform_attn_col2im_bilinear)
 * and its caller loop :263-345 (one of six variants that differ only in how the per-channel partial
 * sums are reduced); the channel reduction here is a plain serial sum.  grad_value must be zeroed by the
 * caller (the reference's autograd glue passes zeros, codetr/ops.py:94-96); grad_loc and grad_weight are
 * fully overwritten.  Serial over (b, q, m) so the scatter-add needs no atomics: test infrastructure,
 * not a performance path.
 */
#define DEFINE_ORACLE_BWD(NAME, T, FLOORFN)                                                      \
  void NAME(const T *value, const int64_t *shapes, const int64_t *starts, const T *loc,          \
            const T *weight, const T *grad_out, T *grad_value, T *grad_loc, T *grad_weight,      \
            int64_t B, int64_t S, int64_t M, int64_t D, int64_t L, int64_t Q, int64_t P) {       \
    const int64_t pix_stride = M * D;                                                            \
    for (int64_t b = 0; b < B; ++b)                                                              \
      for (int64_t q = 0; q < Q; ++q)                                                            \
        for (int64_t m = 0; m < M; ++m) {                                                        \
          const int64_t qm = (b * Q + q) * M + m;                                                \
          const T *go = grad_out + qm * D;                                                       \
          for (int64_t l = 0; l < L; ++l) {                                                      \
            const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];                        \
            const int64_t lvl_off = (b * S + starts[l]) * pix_stride + m * D;                    \
            for (int64_t p = 0; p < P; ++p) {                                                    \
              const int64_t si = qm * L * P + l * P + p;                                         \
              const T x = loc[si * 2], y = loc[si * 2 + 1], aw = weight[si];                     \
              T g_x = (T)0, g_y = (T)0, g_w = (T)0;                                              \
              const T h_im = y * (T)H - (T)0.5, w_im = x * (T)W - (T)0.5;                        \
              if (h_im > (T)-1 && w_im > (T)-1 && h_im < (T)H && w_im < (T)W) {                  \
                const int h_lo = (int)FLOORFN(h_im), w_lo = (int)FLOORFN(w_im);                  \
                const int h_hi = h_lo + 1, w_hi = w_lo + 1;                                      \
                const T lh = h_im - (T)h_lo, lw = w_im - (T)w_lo;                                \
                const T hh = (T)1 - lh, hw = (T)1 - lw;                                          \
                const int ok[4] = {h_lo >= 0 && w_lo >= 0, h_lo >= 0 && w_hi <= W - 1,           \
                                   h_hi <= H - 1 && w_lo >= 0, h_hi <= H - 1 && w_hi <= W - 1};  \
                const int64_t off[4] = {lvl_off + ((int64_t)h_lo * W + w_lo) * pix_stride,       \
                                        lvl_off + ((int64_t)h_lo * W + w_hi) * pix_stride,       \
                                        lvl_off + ((int64_t)h_hi * W + w_lo) * pix_stride,       \
                                        lvl_off + ((int64_t)h_hi * W + w_hi) * pix_stride};      \
                const T cw[4] = {hh * hw, hh * lw, lh * hw, lh * lw};       /* :103 */           \
                const T dh[4] = {-hw, -lw, hw, lw};   /* d(cw)/d(h_im), :111,:118,:125,:132 */   \
                const T dw[4] = {-hh, hh, -lh, lh};   /* d(cw)/d(w_im), :112,:119,:126,:133 */   \
                for (int64_t c = 0; c < D; ++c) {                                                \
                  const T tg = go[c];                                                            \
                  const T tgv = tg * aw;              /* :104 */                                 \
                  T val = (T)0, gh = (T)0, gw = (T)0;                                            \
                  for (int j = 0; j < 4; ++j) {                                                  \
                    if (!ok[j]) continue;                                                        \
                    const T v = value[off[j] + c];                                               \
                    val += cw[j] * v;                                                            \
                    gh += dh[j] * v;                                                             \
                    gw += dw[j] * v;                                                             \
                    grad_value[off[j] + c] += cw[j] * tgv;                  /* :113 */           \
                  }                                                                              \
                  g_w += tg * val;                                          /* :138 */           \
                  g_x += (T)W * gw * tgv;                                   /* :139 */           \
                  g_y += (T)H * gh * tgv;                                   /* :140 */           \
                }                                                                                \
              }                                                                                  \
              grad_loc[si * 2] = g_x;                                                            \
              grad_loc[si * 2 + 1] = g_y;                                                        \
              grad_weight[si] = g_w;                                                             \
            }                                                                                    \
          }                                                                                      \
        }                                                                                        \
  }

DEFINE_ORACLE_BWD(msda_oracle_backward_f32, float, floorf)
DEFINE_ORACLE_BWD(msda_oracle_backward_f64, double, floor)
