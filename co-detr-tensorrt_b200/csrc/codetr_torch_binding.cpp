// codetr_torch_binding.cpp -- optional native registration of the two `codetr::` operators.
//
// The package registers the operators from Python by default-capable code (ops.py); that path costs ~26 us
// of host time per call (dispatcher -> Python -> ctypes).  This translation unit registers the SAME schemas
// (the reference's, codetr/csrc/deformable_attention_torch.cpp:17-23) straight onto the ATen adapter
// (codetr_aten_adapter.cpp -> C ABI), which costs ~7 us per call -- less than the reference's own extension
// (~18 us: at::zeros + zero_() + launch).  When the built library is present the package loads it instead of
// registering from Python; fake kernel and autograd glue stay in Python either way.
#include <ATen/ATen.h>
#include <torch/library.h>

namespace codetr {
// codetr_aten_adapter.cpp
at::Tensor ms_deform_attn_forward(const at::Tensor &value, const at::Tensor &spatial_shapes,
                                  const at::Tensor &level_start_index, const at::Tensor &sampling_loc,
                                  const at::Tensor &attn_weight, const int64_t im2col_step);
void ms_deform_attn_backward(const at::Tensor &value, const at::Tensor &spatial_shapes, const at::Tensor &level_start_index,
                             const at::Tensor &sampling_loc, const at::Tensor &attn_weight, const at::Tensor &grad_output,
                             at::Tensor &grad_value, at::Tensor &grad_sampling_loc, at::Tensor &grad_attn_weight,
                             const int64_t im2col_step);
} // namespace codetr

namespace {

constexpr const char *kForwardSchema =
    "multi_scale_deformable_attention(Tensor value, Tensor spatial_shapes, Tensor level_start_index, "
    "Tensor sampling_loc, Tensor attn_weight, int im2col_step) -> Tensor";
constexpr const char *kBackwardSchema =
    "multi_scale_deformable_attention_backward(Tensor value, Tensor spatial_shapes, Tensor level_start_index, "
    "Tensor sampling_loc, Tensor attn_weight, Tensor grad_output, Tensor(a!) grad_value, "
    "Tensor(b!) grad_sampling_loc, Tensor(c!) grad_attn_weight, int im2col_step) -> ()";

void backward_entry(const at::Tensor &value, const at::Tensor &shapes, const at::Tensor &starts, const at::Tensor &loc,
                    const at::Tensor &weight, const at::Tensor &grad_out, at::Tensor grad_value, at::Tensor grad_loc,
                    at::Tensor grad_weight, int64_t im2col_step) {
  codetr::ms_deform_attn_backward(value, shapes, starts, loc, weight, grad_out, grad_value, grad_loc, grad_weight, im2col_step);
}

} // namespace

TORCH_LIBRARY(codetr, ops) {
  ops.def(kForwardSchema);
  ops.def(kBackwardSchema);
}

TORCH_LIBRARY_IMPL(codetr, CUDA, ops) {
  ops.impl("multi_scale_deformable_attention", &codetr::ms_deform_attn_forward);
  ops.impl("multi_scale_deformable_attention_backward", &backward_entry);
}
