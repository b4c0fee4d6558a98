// codetr_aten_adapter.cpp -- thin C++ adapter from the reference's ATen-typed entry points to the
// C ABI of libmsda_b200.so (include/msda_b200.h).
//
// The reference's two unchanged native callers link against C++-mangled ATen signatures:
//   * the TensorRT plugin declares  codetr::ms_deform_attn_forward_reference  extern
//     (codetr/csrc/deformable_attention_plugin.cpp:64-69) and calls it from enqueue (:351);
//   * the torch binding declares  codetr::ms_deform_attn_forward  and  codetr::ms_deform_attn_backward
//     (codetr/csrc/deformable_attention_torch.cpp:7-14) and registers them (:28-31).
// Compiling this file in place of codetr/csrc/ms_deform_attn.cu (and linking libmsda_b200.so) gives
// both of them the B200 kernel with no source change on their side -- see INTEGRATION.md.
//
// Behaviour mirrored from ms_deform_attn.cu:899-973: contiguity / CUDA asserts raise c10::Error, the
// im2col_step divisibility rule is enforced (by the C ABI), the launch goes to ATen's *current* stream
// (the plugin makes TensorRT's stream current before calling, plugin.cpp:330-333).  Differences, all
// deliberate: the output is not zero-filled first (every element is written by the kernel), one launch
// covers the whole batch, bf16 is accepted, and launch errors are raised instead of printf'ed (:775-778).
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include "msda_b200.h"

namespace codetr {

namespace {
int to_msda_dtype(at::ScalarType t) {
  switch (t) {
  case at::kFloat:
    return MSDA_F32;
  case at::kHalf:
    return MSDA_F16;
  case at::kBFloat16:
    return MSDA_BF16;
  case at::kDouble:
    return MSDA_F64;
  default:
    TORCH_CHECK(false, "ms_deform_attn_forward: unsupported dtype ", t);
  }
}

// Extent / device agreement between the five inputs: what the reference's fake kernel asserts at trace time
// (codetr/ops.py:59-84).  The C ABI trusts its caller's sizes, so a mismatch here would become an out-of-bounds
// device read instead of an error.
void check_extents(const at::Tensor &value, const at::Tensor &spatial_shapes, const at::Tensor &level_start_index,
                   const at::Tensor &sampling_loc, const at::Tensor &attn_weight) {
  TORCH_CHECK(level_start_index.dim() == 1 && attn_weight.dim() == 5, "unexpected tensor ranks");
  TORCH_CHECK(spatial_shapes.size(1) == 2, "spatial_shapes must be [num_levels, 2]");
  TORCH_CHECK(level_start_index.size(0) == spatial_shapes.size(0), "level_start_index must have one entry per level");
  TORCH_CHECK(sampling_loc.size(0) == value.size(0) && sampling_loc.size(2) == value.size(2) &&
                  sampling_loc.size(3) == spatial_shapes.size(0) && sampling_loc.size(5) == 2,
              "sampling_loc must be [batch, num_query, num_heads, num_levels, num_points, 2]");
  for (int axis = 0; axis < 5; ++axis)
    TORCH_CHECK(attn_weight.size(axis) == sampling_loc.size(axis), "attn_weight must match sampling_loc on axis ", axis);
  const auto dev = value.device();
  TORCH_CHECK(spatial_shapes.device() == dev && level_start_index.device() == dev && sampling_loc.device() == dev &&
                  attn_weight.device() == dev,
              "all tensors must be on the same CUDA device");
}
} // namespace

void ms_deform_attn_forward_reference(const at::Tensor &value, const at::Tensor &spatial_shapes,
                                      const at::Tensor &level_start_index, const at::Tensor &sampling_loc,
                                      const at::Tensor &attn_weight, at::Tensor &output, const int64_t im2col_step) {
  TORCH_CHECK(value.is_contiguous(), "value tensor has to be contiguous");
  TORCH_CHECK(spatial_shapes.is_contiguous(), "spatial_shapes tensor has to be contiguous");
  TORCH_CHECK(level_start_index.is_contiguous(), "level_start_index tensor has to be contiguous");
  TORCH_CHECK(sampling_loc.is_contiguous(), "sampling_loc tensor has to be contiguous");
  TORCH_CHECK(attn_weight.is_contiguous(), "attn_weight tensor has to be contiguous");
  TORCH_CHECK(output.is_contiguous(), "output tensor has to be contiguous");
  TORCH_CHECK(value.is_cuda() && spatial_shapes.is_cuda() && level_start_index.is_cuda() && sampling_loc.is_cuda() &&
                  attn_weight.is_cuda() && output.is_cuda(),
              "all tensors must be CUDA tensors");
  TORCH_CHECK(value.dim() == 4 && sampling_loc.dim() == 6 && attn_weight.dim() == 5 && spatial_shapes.dim() == 2,
              "unexpected tensor ranks");
  TORCH_CHECK(spatial_shapes.scalar_type() == at::kLong && level_start_index.scalar_type() == at::kLong,
              "spatial_shapes / level_start_index must be int64");
  TORCH_CHECK(sampling_loc.scalar_type() == value.scalar_type() && attn_weight.scalar_type() == value.scalar_type() &&
                  output.scalar_type() == value.scalar_type(),
              "value, sampling_loc, attn_weight and output must share one dtype");
  check_extents(value, spatial_shapes, level_start_index, sampling_loc, attn_weight);
  TORCH_CHECK(output.device() == value.device(), "output must be on the same CUDA device");

  const int64_t batch = value.size(0), num_keys = value.size(1), num_heads = value.size(2), channels = value.size(3);
  const int64_t num_levels = spatial_shapes.size(0);
  const int64_t num_query = sampling_loc.size(1), num_point = sampling_loc.size(4);
  TORCH_CHECK(output.size(0) == batch, "output size(0) must be equal to batch");
  TORCH_CHECK(output.size(1) == num_query, "output size(1) must be equal to num_query");
  TORCH_CHECK(output.size(2) == num_heads * channels, "output size(2) must be equal to num_heads * channels");

  const c10::cuda::CUDAGuard device_guard(value.device());
  cudaStream_t stream = at::cuda::getCurrentCUDAStream();
  const int rc = msda_b200_forward(value.data_ptr(), spatial_shapes.data_ptr<int64_t>(), level_start_index.data_ptr<int64_t>(),
                                   sampling_loc.data_ptr(), attn_weight.data_ptr(), output.data_ptr(), batch, num_keys,
                                   num_heads, channels, num_levels, num_query, num_point, im2col_step,
                                   to_msda_dtype(value.scalar_type()), MSDA_FLAG_DEFAULT, stream);
  TORCH_CHECK(rc == 0, "ms_deform_attn_forward: ", msda_b200_error_string(rc), " (batch=", batch,
              ", im2col_step=", im2col_step, ")");
}

at::Tensor ms_deform_attn_forward(const at::Tensor &value, const at::Tensor &spatial_shapes,
                                  const at::Tensor &level_start_index, const at::Tensor &sampling_loc,
                                  const at::Tensor &attn_weight, const int64_t im2col_step) {
  TORCH_CHECK(value.dim() == 4 && sampling_loc.dim() == 6, "unexpected tensor ranks");
  auto output = at::empty({value.size(0), sampling_loc.size(1), value.size(2) * value.size(3)}, value.options());
  ms_deform_attn_forward_reference(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, output, im2col_step);
  return output;
}

// Backward (ms_deform_attn.cu:975-1028): same asserts, then the C ABI.  grad_value is accumulated into
// (the reference's autograd glue passes zeros, codetr/ops.py:94-96), the other two are overwritten.
void ms_deform_attn_backward(const at::Tensor &value, const at::Tensor &spatial_shapes, const at::Tensor &level_start_index,
                             const at::Tensor &sampling_loc, const at::Tensor &attn_weight, const at::Tensor &grad_output,
                             at::Tensor &grad_value, at::Tensor &grad_sampling_loc, at::Tensor &grad_attn_weight,
                             const int64_t im2col_step) {
  TORCH_CHECK(value.is_contiguous(), "value tensor has to be contiguous");
  TORCH_CHECK(spatial_shapes.is_contiguous(), "spatial_shapes tensor has to be contiguous");
  TORCH_CHECK(level_start_index.is_contiguous(), "level_start_index tensor has to be contiguous");
  TORCH_CHECK(sampling_loc.is_contiguous(), "sampling_loc tensor has to be contiguous");
  TORCH_CHECK(attn_weight.is_contiguous(), "attn_weight tensor has to be contiguous");
  TORCH_CHECK(grad_output.is_contiguous(), "grad_output tensor has to be contiguous");
  TORCH_CHECK(grad_value.is_contiguous() && grad_sampling_loc.is_contiguous() && grad_attn_weight.is_contiguous(),
              "gradient tensors have to be contiguous");
  TORCH_CHECK(value.is_cuda() && spatial_shapes.is_cuda() && level_start_index.is_cuda() && sampling_loc.is_cuda() &&
                  attn_weight.is_cuda() && grad_output.is_cuda() && grad_value.is_cuda() && grad_sampling_loc.is_cuda() &&
                  grad_attn_weight.is_cuda(),
              "all tensors must be CUDA tensors");
  TORCH_CHECK(value.dim() == 4 && sampling_loc.dim() == 6 && spatial_shapes.dim() == 2, "unexpected tensor ranks");
  const auto st = value.scalar_type();
  TORCH_CHECK(sampling_loc.scalar_type() == st && attn_weight.scalar_type() == st && grad_output.scalar_type() == st &&
                  grad_value.scalar_type() == st && grad_sampling_loc.scalar_type() == st && grad_attn_weight.scalar_type() == st,
              "all floating tensors must share one dtype");
  TORCH_CHECK(grad_value.sizes() == value.sizes() && grad_sampling_loc.sizes() == sampling_loc.sizes() &&
                  grad_attn_weight.sizes() == attn_weight.sizes(),
              "gradient shapes must match their tensors");
  check_extents(value, spatial_shapes, level_start_index, sampling_loc, attn_weight);
  TORCH_CHECK(grad_output.dim() == 3 && grad_output.size(0) == value.size(0) && grad_output.size(1) == sampling_loc.size(1) &&
                  grad_output.size(2) == value.size(2) * value.size(3),
              "grad_output must be [batch, num_query, num_heads * channels]");
  const c10::cuda::CUDAGuard device_guard(value.device());
  cudaStream_t stream = at::cuda::getCurrentCUDAStream();
  const int rc = msda_b200_backward(value.data_ptr(), spatial_shapes.data_ptr<int64_t>(), level_start_index.data_ptr<int64_t>(),
                                    sampling_loc.data_ptr(), attn_weight.data_ptr(), grad_output.data_ptr(), grad_value.data_ptr(),
                                    grad_sampling_loc.data_ptr(), grad_attn_weight.data_ptr(), value.size(0), value.size(1),
                                    value.size(2), value.size(3), spatial_shapes.size(0), sampling_loc.size(1),
                                    sampling_loc.size(4), im2col_step, to_msda_dtype(st), MSDA_FLAG_DEFAULT, stream);
  TORCH_CHECK(rc == 0, "ms_deform_attn_backward: ", msda_b200_error_string(rc));
}

} // namespace codetr
