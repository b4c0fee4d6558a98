// Binding used ONLY to build the reference's own CUDA op as a second oracle / the "kernel to beat"
// (oracle/_ref/msda_ref_cuda.so; test infrastructure, never part of the product).
//
// The reference's kernel file, /root/reference/codetr/csrc/ms_deform_attn.cu, is compiled from where
// it lies (see build_ref.py); this file only registers its forward and backward entry points
// (codetr::ms_deform_attn_forward, ms_deform_attn.cu:958-973; codetr::ms_deform_attn_backward, :975-1028) under a *different* torch library
// namespace, `codetr_ref`, so that it can be loaded next to this repo's own
// `codetr::multi_scale_deformable_attention`.
#include <ATen/ATen.h>
#include <torch/library.h>

namespace codetr {
at::Tensor ms_deform_attn_forward(const at::Tensor &value, const at::Tensor &spatial_shapes,
                                  const at::Tensor &level_start_index, const at::Tensor &sampling_loc,
                                  const at::Tensor &attn_weight, const int64_t im2col_step);
void ms_deform_attn_backward(const at::Tensor &value, const at::Tensor &spatial_shapes, const at::Tensor &level_start_index,
                             const at::Tensor &sampling_loc, const at::Tensor &attn_weight, const at::Tensor &grad_output,
                             at::Tensor &grad_value, at::Tensor &grad_sampling_loc, at::Tensor &grad_attn_weight,
                             const int64_t im2col_step);
}

TORCH_LIBRARY(codetr_ref, m) {
  m.def("multi_scale_deformable_attention(Tensor value, Tensor spatial_shapes, Tensor level_start_index, "
        "Tensor sampling_loc, Tensor attn_weight, int im2col_step) -> Tensor");
  m.def("multi_scale_deformable_attention_backward(Tensor value, Tensor spatial_shapes, Tensor level_start_index, "
        "Tensor sampling_loc, Tensor attn_weight, Tensor grad_output, Tensor(a!) grad_value, Tensor(b!) grad_sampling_loc, "
        "Tensor(c!) grad_attn_weight, int im2col_step) -> ()");
}

TORCH_LIBRARY_IMPL(codetr_ref, CUDA, m) {
  m.impl("multi_scale_deformable_attention", &codetr::ms_deform_attn_forward);
  m.impl("multi_scale_deformable_attention_backward", &codetr::ms_deform_attn_backward);
}
