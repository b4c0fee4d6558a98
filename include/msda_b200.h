/*
 * msda_b200.h -- C ABI of the B200-native (sm_100a) multi-scale deformable
 * attention forward core.
 *
 * This is the drop-in boundary of the repo: a plain-C shared library
 * (libmsda_b200.so) with no torch / ATen / TensorRT types in its signatures.
 * Every entry point below names the reference interface it stands behind
 * (citations are file:line under the reference checkout, anenbergb/Co-DETR-TensorRT).
 *
 * Conventions shared by all launch entry points
 *   - all data pointers are DEVICE pointers unless the name says "_host";
 *   - tensors are dense row-major ("contiguous" in the reference's asserts,
 *     codetr/csrc/ms_deform_attn.cu:902-912):
 *         value              [B, S, M, D]          element type = dtype
 *         spatial_shapes     [L, 2]  int64, (H, W) per level          (device)
 *         level_start_index  [L]     int64, first key of each level   (device)
 *         sampling_loc       [B, Q, M, L, P, 2]    (x, y) in [0,1] units
 *         attn_weight        [B, Q, M, L, P]       already soft-maxed
 *         output             [B, Q, M*D]           fully overwritten
 *   - the launcher never allocates, never synchronises, never reads device
 *     memory on the host, and enqueues work only on `stream` (a cudaStream_t
 *     passed as void*; NULL = legacy default stream).  It is therefore safe
 *     under CUDA-graph capture and re-entrant from several host threads, which
 *     is what TensorRT's enqueue contract needs
 *     (codetr/csrc/deformable_attention_plugin.cpp:285-355, README.md:193);
 *   - `output` does not have to be zeroed by the caller (the reference zero
 *     fills twice, ms_deform_attn.cu:936 and :968, then overwrites);
 *   - return value: 0 on success, a negative MSDA_ERR_* for a rejected call,
 *     or a positive cudaError_t if the launch itself failed (the reference
 *     only printf()s launch errors, ms_deform_attn.cu:775-778).
 */
#ifndef MSDA_B200_H_
#define MSDA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSDA_B200_ABI_VERSION 1

/* Element type of value / sampling_loc / attn_weight / output.  The reference
 * dispatches double, float and half (AT_DISPATCH_FLOATING_TYPES_AND_HALF,
 * ms_deform_attn.cu:946); bf16 is new here. */
enum msda_dtype {
  MSDA_F32 = 0,
  MSDA_F16 = 1,
  MSDA_BF16 = 2,
  MSDA_F64 = 3
};

enum msda_error {
  MSDA_OK = 0,
  MSDA_ERR_NULL_POINTER = -1,   /* a required pointer is NULL                         */
  MSDA_ERR_BAD_SHAPE = -2,      /* a dimension is negative / zero where not allowed   */
  MSDA_ERR_BAD_DTYPE = -3,      /* dtype is not one of msda_dtype                     */
  MSDA_ERR_BAD_STEP = -4,       /* B % min(B, im2col_step) != 0 (ms_deform_attn.cu:924-926) */
  MSDA_ERR_MISALIGNED = -5,     /* a pointer is not aligned to its element size       */
  MSDA_ERR_UNSUPPORTED = -6,    /* shape outside what the kernels index (e.g. > 2^31 keys) */
  MSDA_ERR_BAD_FLAGS = -7,
  MSDA_ERR_WORKSPACE_TOO_SMALL = -8 /* a caller-provided workspace is smaller than the matching *_workspace_bytes() */
};

/* Launch flags (bit field).  0 = library defaults. */
enum msda_flags {
  MSDA_FLAG_DEFAULT = 0,
  MSDA_FLAG_FORCE_GENERIC = 1 << 0, /* use the shape-agnostic scalar kernel               */
  MSDA_FLAG_LINEAR_ORDER = 1 << 1,  /* do not re-tile queries spatially (encoder shapes)  */
  MSDA_FLAG_MATH_FHFMA = 1 << 2,    /* fp16/bf16: Blackwell FHFMA, combined weights rounded to the 16-bit type
                                       (default for fp16).  bf16 without a math flag: fp32 weights, except on the
                                       head-pair kernel's shapes, where each weight is carried as two bf16 terms
                                       (hi + lo, 2^-17 relative: same max error as fp32 weights, 14 % faster) */
  MSDA_FLAG_MATH_EXACT = 1 << 3,    /* fp16/bf16: fp32 weights, convert + FFMA, on every shape */
  MSDA_FLAG_NO_STAGING = 1 << 4,    /* never stage loc/weights through shared memory (the default) */
  MSDA_FLAG_STAGE_TMA = 1 << 5,     /* stage loc/weights of each pass with TMA bulk copies (opt-in:
                                       measured slightly slower than direct loads) */
  MSDA_FLAG_NO_PACKED = 1 << 6,     /* ignore the workspace: never use the packed-pyramid path */
  MSDA_FLAG_HEAD_MAJOR = 1 << 7,    /* a warp holds one head of neighbouring queries (default: query-major) */
  MSDA_FLAG_NO_SMEM_LEVELS = 1 << 9, /* never use the head-pair kernel that keeps the coarse pyramid levels in
                                       shared memory (msda_fwd_hp); every level is gathered from global memory */
  MSDA_FLAG_PDL = 1 << 8            /* The launches always carry the programmatic-stream-serialization attribute
                                       (CTAs may be scheduled while the preceding kernel of the stream drains; all
                                       reads wait for that kernel -- semantics are plain stream order).  This flag
                                       additionally lets the kernel read spatial_shapes / level_start_index BEFORE
                                       that wait: only set it if the two level tables are not produced by the
                                       immediately preceding kernel (they are constants in every known caller). */
};

/*
 * msda_b200_forward -- the forward core.
 *
 * Stands behind  codetr::ms_deform_attn_forward_reference
 *   (codetr/csrc/ms_deform_attn.cu:899-956; declared extern by the TensorRT
 *   plugin, codetr/csrc/deformable_attention_plugin.cpp:64-69, called :351)
 * and, with a caller-allocated output, behind  codetr::ms_deform_attn_forward
 *   (ms_deform_attn.cu:958-973; bound to the torch op
 *   codetr::multi_scale_deformable_attention in
 *   codetr/csrc/deformable_attention_torch.cpp:17-19, :28-31).
 *
 * im2col_step keeps the reference's argument: it must satisfy
 * B % min(B, im2col_step) == 0, otherwise MSDA_ERR_BAD_STEP.  It does not
 * change the result and (unlike the reference, which launches once per chunk)
 * does not change the number of launches: one kernel covers the whole batch.
 */
int msda_b200_forward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                      const void *sampling_loc, const void *attn_weight, void *output, int64_t batch,
                      int64_t num_keys, int64_t num_heads, int64_t channels, int64_t num_levels,
                      int64_t num_queries, int64_t num_points, int64_t im2col_step, int dtype,
                      unsigned flags, void *stream);

/*
 * msda_b200_forward_ws -- msda_b200_forward with an optional caller-owned device WORKSPACE.
 *
 * With a workspace of at least msda_b200_workspace_bytes(...) bytes (128-byte aligned) the library may
 * take the packed path: a pre-pass re-lays the value pyramid out in the workspace so that one 128-byte
 * line holds a (pixel, head) row and its right-hand neighbour, and the gather kernel fetches both
 * horizontal corners of a sample with one 256-bit load (two L1 wavefronts per sample instead of four).
 * Results are the same bits as msda_b200_forward's for the same math mode.  The workspace is scratch: it
 * carries no state between calls.  msda_b200_workspace_bytes returns 0 when the packed path does not
 * apply to the shape (then any workspace is ignored).  Stands behind the same reference interface as
 * msda_b200_forward; the workspace corresponds to TensorRT's plugin workspace
 * (IPluginV3OneRuntime::getWorkspaceSize, deformable_attention_plugin.cpp:371-374 returns 0 today).
 */
size_t msda_b200_workspace_bytes(int64_t batch, int64_t num_keys, int64_t num_heads, int64_t channels,
                                 int64_t num_levels, int64_t num_queries, int64_t num_points, int dtype);
int msda_b200_forward_ws(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                         const void *sampling_loc, const void *attn_weight, void *output, void *workspace,
                         size_t workspace_bytes, int64_t batch, int64_t num_keys, int64_t num_heads,
                         int64_t channels, int64_t num_levels, int64_t num_queries, int64_t num_points,
                         int64_t im2col_step, int dtype, unsigned flags, void *stream);

/*
 * msda_b200_plugin_enqueue -- the TensorRT IPluginV3OneRuntime::enqueue body
 * without libtorch.
 *
 * Stands behind  DeformableAttentionPlugin::enqueue
 *   (codetr/csrc/deformable_attention_plugin.cpp:285-355): `inputs` is the
 *   plugin's five device pointers in the plugin's order {value,
 *   spatial_shapes, level_start_index, sampling_loc, attn_weight}, `outputs[0]`
 *   the output buffer; value_dims = inputDesc[0].dims.d (4 entries),
 *   loc_dims = inputDesc[3].dims.d (6 entries); trt_dtype is the integer value
 *   of nvinfer1::DataType of input 0 (kFLOAT = 0, kHALF = 1, kBF16 = 7).
 * Returns 0 on success and non-zero on failure, like enqueue (:320-325).
 * `workspace` / `workspace_bytes` are TensorRT's plugin workspace (may be NULL / 0); a plugin that wants the
 * packed path returns msda_b200_plugin_workspace_bytes(...) from getWorkspaceSize.
 */
size_t msda_b200_plugin_workspace_bytes(const int64_t *value_dims, const int64_t *loc_dims, int trt_dtype);
int msda_b200_plugin_enqueue(const int64_t *value_dims, const int64_t *loc_dims, int trt_dtype,
                             const void *const *inputs, void *const *outputs, void *workspace,
                             size_t workspace_bytes, int64_t im2col_step, void *stream);

/*
 * msda_b200_forward_fused -- opt-in producer-fused mode (not part of the
 * reference's operator API).
 *
 * Replaces, for callers that opt in, the softmax over L*P and the sampling
 * location arithmetic that the reference module performs before calling the
 * op (codetr/multi_scale_deformable_attention.py:180-200), so the [B,Q,M,L,P]
 * weights and [B,Q,M,L,P,2] locations never round-trip through HBM:
 *     attn_logits      [B, Q, M, L*P]     pre-softmax
 *     sampling_offsets [B, Q, M, L, P, 2] raw Linear output
 *     reference_points [B, Q, L, ref_dim] ref_dim = 2 (x,y) or 4 (cx,cy,w,h)
 */
int msda_b200_forward_fused(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                            const void *reference_points, const void *sampling_offsets,
                            const void *attn_logits, void *output, int64_t batch, int64_t num_keys,
                            int64_t num_heads, int64_t channels, int64_t num_levels, int64_t num_queries,
                            int64_t num_points, int64_t ref_dim, int dtype, unsigned flags, void *stream);

/*
 * msda_b200_forward_host -- the same call for HOST buffers: copies the five
 * inputs host->device into a caller-provided device workspace, launches, and
 * copies the output back, all on `stream` (asynchronous when the host buffers
 * are pinned).  This is the end-to-end path a caller without device-resident
 * tensors takes (what bench.py reports as "e2e").  workspace_bytes must be at
 * least msda_b200_host_workspace_bytes(...).  The caller synchronises `stream`.
 */
size_t msda_b200_host_workspace_bytes(int64_t batch, int64_t num_keys, int64_t num_heads, int64_t channels,
                                      int64_t num_levels, int64_t num_queries, int64_t num_points, int dtype);
int msda_b200_forward_host(const void *value_host, const int64_t *spatial_shapes_host,
                           const int64_t *level_start_index_host, const void *sampling_loc_host,
                           const void *attn_weight_host, void *output_host, void *workspace_dev,
                           size_t workspace_bytes, int64_t batch, int64_t num_keys, int64_t num_heads,
                           int64_t channels, int64_t num_levels, int64_t num_queries, int64_t num_points,
                           int64_t im2col_step, int dtype, unsigned flags, void *stream);

/*
 * msda_b200_backward -- gradients of the forward.
 *
 * Stands behind  codetr::ms_deform_attn_backward  (codetr/csrc/ms_deform_attn.cu:975-1028; bound to the torch
 * op codetr::multi_scale_deformable_attention_backward, deformable_attention_torch.cpp:20-23, :30; called
 * by the autograd glue codetr/ops.py:90-126).  grad_output [B,Q,M*D]; grad_value [B,S,M,D] is ACCUMULATED
 * into (atomics) and must be zero-initialised by the caller, as the reference's glue does (ops.py:94-96);
 * grad_sampling_loc [B,Q,M,L,P,2] and grad_attn_weight [B,Q,M,L,P] are fully overwritten.  Per-sample
 * derivative follows ms_deform_attn.cu:79-141 (arithmetic in fp32 for 16-bit types, fp64 for double).
 */
int msda_b200_backward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                       const void *sampling_loc, const void *attn_weight, const void *grad_output,
                       void *grad_value, void *grad_sampling_loc, void *grad_attn_weight, int64_t batch,
                       int64_t num_keys, int64_t num_heads, int64_t channels, int64_t num_levels,
                       int64_t num_queries, int64_t num_points, int64_t im2col_step, int dtype,
                       unsigned flags, void *stream);

/*
 * Packed value pyramid ("emit the layout the kernel prefers").  No reference counterpart: the
 * reference's module hands the op the plain [B, S, M, D] tensor (multi_scale_deformable_attention.py:173-176).
 *
 * Layout: one 128-byte, line-aligned entry per (key s, head m) at byte offset ((b*S + s)*M + m)*128, holding the key's
 * 64-byte row AND the row of the key to its right in the flattened pyramid (s + 1), interleaved in 16-byte chunks:
 *     [ v(s)[0:8] | v(s+1)[0:8] | v(s)[8:16] | v(s+1)[8:16] | v(s)[16:24] | v(s+1)[16:24] | v(s)[24:32] | v(s+1)[24:32] ]
 * so the two horizontal corners of a bilinear sample arrive with one 32-byte load per lane and one L1 wavefront per lane
 * group instead of two.  The right-hand half of the last key of an image row is never used (its weight is zero), so a
 * producer may leave anything there.  16-bit types, channels == 32 only.
 *
 *   msda_b200_packed_value_bytes   size of the packed buffer, or 0 when the shape / dtype has no packed form
 *   msda_b200_pack_value           plain value tensor -> packed buffer (a pre-pass; the projection kernel below can write
 *                                  the packed layout directly instead: msda_b200_value_proj with `packed_out`)
 *   msda_b200_forward_packed       the operator on a packed value buffer: same sampling_loc / attn_weight / output
 *                                  contract and the same bits as msda_b200_forward on the plain tensor.  num_heads == 8,
 *                                  num_points == 4, 2 <= num_levels <= 8, else MSDA_ERR_UNSUPPORTED (the caller keeps
 *                                  the plain path).  `packed_value` 32-byte aligned.
 */
size_t msda_b200_packed_value_bytes(int64_t batch, int64_t num_keys, int64_t num_heads, int64_t channels, int dtype);
int msda_b200_pack_value(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index, void *packed,
                         int64_t batch, int64_t num_keys, int64_t num_heads, int64_t channels, int64_t num_levels, int dtype,
                         void *stream);
int msda_b200_forward_packed(const void *packed_value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                             const void *sampling_loc, const void *attn_weight, void *output, int64_t batch,
                             int64_t num_keys, int64_t num_heads, int64_t channels, int64_t num_levels, int64_t num_queries,
                             int64_t num_points, int dtype, unsigned flags, void *stream);

/* ---- introspection / measurement helpers (no reference counterpart) ---- */

/* ABI version of the loaded library (== MSDA_B200_ABI_VERSION). */
int msda_b200_abi_version(void);

/* Human-readable text for a return value of the functions above. */
const char *msda_b200_error_string(int code);

/* Number of kernels this library has launched in this process (all threads). */
uint64_t msda_b200_launch_count(void);

/* Name of the kernel variant the last successful msda_b200_forward* call on
 * this host thread selected, e.g. "vec<f16,D32,P4>/tiled/exact". */
const char *msda_b200_last_variant(void);

/* Algorithmic byte counts of one forward call:
 * compulsory HBM bytes (every input read once, output written once) and the
 * no-reuse gather bytes (every corner row fetched separately). */
uint64_t msda_b200_algorithmic_hbm_bytes(int64_t batch, int64_t num_keys, int64_t num_heads, int64_t channels,
                                         int64_t num_levels, int64_t num_queries, int64_t num_points, int dtype);
uint64_t msda_b200_algorithmic_gather_bytes(int64_t batch, int64_t num_heads, int64_t channels,
                                            int64_t num_levels, int64_t num_queries, int64_t num_points,
                                            int dtype);

/* Producer of the `value` input: nn.Linear + masked_fill + head split of the
 * calling module (/root/reference/codetr/multi_scale_deformable_attention.py:173-176) in one tensor-core
 * kernel (TMA-staged operands, tcgen05.mma into tensor memory, bias / mask / rounding in the epilogue):
 *
 *     value[r, :] = key_padding_mask[r] ? 0 : x[r, :] @ weight^T + bias,        r in [0, rows = B*S)
 *
 * x [rows, in_features], weight [out_features, in_features] (nn.Linear layout), bias [out_features] or NULL,
 * key_padding_mask [rows] bytes (non-zero = padded key, torch.bool layout) or NULL, value [rows, out_features]
 * = the [B, S, M, D] tensor msda_b200_forward reads.  x / weight / bias / value share `dtype` (MSDA_F16 or
 * MSDA_BF16), are contiguous and 16-byte aligned.  Supported: in_features and out_features multiples of 64,
 * at most 256 (msda_b200_value_proj_supported says so without a GPU); anything else returns
 * MSDA_ERR_UNSUPPORTED and the caller keeps its own GEMM.  Same conventions as the other entry points: no
 * allocation, no synchronisation, launches only on `stream`, capture-safe. */
int msda_b200_value_proj_supported(int64_t in_features, int64_t out_features, int dtype);
/* The consumer of the op's output, same kernel with a residual epilogue: output_proj + (inference-mode) dropout +
 * residual of the calling module (/root/reference/codetr/multi_scale_deformable_attention.py:212-218):
 *
 *     out[r, :] = round(attended[r, :] @ weight^T + bias) + residual[r, :]
 *
 * (the Linear result is rounded to the element type before the residual is added, like the two PyTorch ops it
 * replaces).  attended [rows, in_features] = the [B, Q, M*D] output of msda_b200_forward, residual and out
 * [rows, out_features]; `out` may alias `residual`.  Same shape / dtype limits as msda_b200_value_proj. */
int msda_b200_output_proj(const void *attended, const void *weight, const void *bias, const void *residual, void *out,
                          int64_t rows, int64_t in_features, int64_t out_features, int dtype, unsigned flags, void *stream);
int msda_b200_value_proj(const void *x, const void *weight, const void *bias, const unsigned char *key_padding_mask,
                         void *value, int64_t rows, int64_t in_features, int64_t out_features, int dtype, unsigned flags,
                         void *stream);

/* Read-bandwidth probe used for the roofline denominators that
 * MEASURED_PEAKS.json does not carry (L2): every thread block streams
 * `bytes` of `buf` (16-byte loads) `repeats` times; a working set below the
 * L2 capacity measures L2->SM bandwidth, a larger one HBM.  `sink` receives a
 * checksum so the loads cannot be elided (>= 4 bytes, device). */
int msda_b200_read_probe(const void *buf, size_t bytes, int repeats, void *sink, void *stream);

#ifdef __cplusplus
}
#endif

#endif /* MSDA_B200_H_ */
