// Test infrastructure only (never part of the product): exposes the REFERENCE's own CUDA implementation,
// compiled from /root/reference/codetr/csrc/ms_deform_attn.cu where it lies (see build_ref.py), to Python as
//     torch.ops.codetr_ref.msda_forward / torch.ops.codetr_ref.msda_backward
// so that it can be loaded next to this repo's `codetr::multi_scale_deformable_attention` and used as the
// "kernel to beat" and as a second, CUDA-side oracle.  The two entry points it wraps are the reference's
// codetr::ms_deform_attn_forward (ms_deform_attn.cu:958-973) and codetr::ms_deform_attn_backward (:975-1028).
#include <ATen/ATen.h>
#include <torch/library.h>

namespace codetr {
// defined in the reference's ms_deform_attn.cu
at::Tensor ms_deform_attn_forward(const at::Tensor &, const at::Tensor &, const at::Tensor &, const at::Tensor &,
                                  const at::Tensor &, const int64_t);
void ms_deform_attn_backward(const at::Tensor &, const at::Tensor &, const at::Tensor &, const at::Tensor &,
                             const at::Tensor &, const at::Tensor &, at::Tensor &, at::Tensor &, at::Tensor &,
                             const int64_t);
} // namespace codetr

namespace {

at::Tensor ref_forward(const at::Tensor &pyramid, const at::Tensor &level_hw, const at::Tensor &level_first,
                       const at::Tensor &xy, const at::Tensor &weights, int64_t step) {
  return codetr::ms_deform_attn_forward(pyramid, level_hw, level_first, xy, weights, step);
}

void ref_backward(const at::Tensor &pyramid, const at::Tensor &level_hw, const at::Tensor &level_first,
                  const at::Tensor &xy, const at::Tensor &weights, const at::Tensor &d_out, at::Tensor d_pyramid,
                  at::Tensor d_xy, at::Tensor d_weights, int64_t step) {
  codetr::ms_deform_attn_backward(pyramid, level_hw, level_first, xy, weights, d_out, d_pyramid, d_xy, d_weights, step);
}

} // namespace

TORCH_LIBRARY(codetr_ref, lib) {
  lib.def("msda_forward(Tensor pyramid, Tensor level_hw, Tensor level_first, Tensor xy, Tensor weights, int step) -> Tensor");
  lib.def("msda_backward(Tensor pyramid, Tensor level_hw, Tensor level_first, Tensor xy, Tensor weights, Tensor d_out, "
          "Tensor(a!) d_pyramid, Tensor(b!) d_xy, Tensor(c!) d_weights, int step) -> ()");
}

TORCH_LIBRARY_IMPL(codetr_ref, CUDA, lib) {
  lib.impl("msda_forward", &ref_forward);
  lib.impl("msda_backward", &ref_backward);
}
