"""B200-native (sm_100a) multi-scale deformable attention forward for Co-DETR.

Import name: ``codetr_b200`` (the directory keeps the repo-mandated name
``co-detr-tensorrt_b200``, which is not a Python identifier; ``codetr_b200.py`` at the repo
root loads it under the importable name).

Importing the package loads ``csrc/libmsda_b200.so`` and registers
``torch.ops.codetr.multi_scale_deformable_attention`` with the reference's schema.  It raises if
the library has not been built: there is no CPU or PyTorch fallback.
"""
from . import _native
from ._native import (FLAG_FORCE_GENERIC, FLAG_LINEAR_ORDER, FLAG_MATH_EXACT, FLAG_HEAD_MAJOR, FLAG_MATH_FHFMA, FLAG_PDL, FLAG_NO_PACKED, FLAG_NO_SMEM_LEVELS, FLAG_NO_STAGING, FLAG_STAGE_TMA,
                      NativeLibraryError, build_native, last_variant, launch_count)
from . import workloads
from . import sharding
from . import ops
from .module import MultiScaleDeformableAttention
from .ops import (HostForward, HostPipeline, PreparedForward, backward_into, forward_fused, forward_into, forward_packed, pack_value, multi_scale_deformable_attention, plugin_enqueue,
                  read_bandwidth_probe, set_default_flags, set_use_workspace, output_proj, value_proj, value_proj_supported,
                  workspace_bytes)

__all__ = [
    "MultiScaleDeformableAttention", "multi_scale_deformable_attention", "forward_into", "backward_into", "forward_fused", "plugin_enqueue", "HostForward", "HostPipeline", "PreparedForward",
    "set_default_flags", "read_bandwidth_probe", "build_native", "launch_count", "last_variant", "workloads", "sharding",
    "FLAG_FORCE_GENERIC", "FLAG_LINEAR_ORDER", "FLAG_MATH_EXACT", "FLAG_MATH_FHFMA", "FLAG_NO_STAGING", "FLAG_STAGE_TMA", "FLAG_NO_PACKED", "FLAG_NO_SMEM_LEVELS", "FLAG_HEAD_MAJOR", "FLAG_PDL", "workspace_bytes", "set_use_workspace",
    "NativeLibraryError", "value_proj", "value_proj_supported", "output_proj", "pack_value", "forward_packed",
]
