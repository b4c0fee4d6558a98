"""The drop-in boundary: the reference's UNCHANGED torch binding (deformable_attention_torch.cpp) linked
against this repo's ATen adapter + libmsda_b200.so (tools/build_dropin.py, build container only).

CPU tier: the built library exports the three C++ symbols the reference's native callers link against.
GPU tier: loaded in a fresh process (it registers the same `codetr` namespace as the Python registration,
so the two cannot share a process), the op reproduces the golden vectors."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "co-detr-tensorrt_b200", "csrc", "_dropin", "codetr_cpp_extension.so")
needs_dropin = pytest.mark.skipif(not os.path.isfile(DROPIN), reason="drop-in not built (tools/build_dropin.py needs /root/reference)")


@needs_dropin
def test_dropin_exports_reference_symbols():
    nm = subprocess.run(["nm", "-D", "--defined-only", "-C", DROPIN], capture_output=True, text=True).stdout
    # deformable_attention_plugin.cpp:64-69 and deformable_attention_torch.cpp:7-14
    assert "codetr::ms_deform_attn_forward_reference(at::Tensor const&, at::Tensor const&, at::Tensor const&, at::Tensor const&, at::Tensor const&, at::Tensor&, long)" in nm
    assert "codetr::ms_deform_attn_forward(at::Tensor const&, at::Tensor const&, at::Tensor const&, at::Tensor const&, at::Tensor const&, long)" in nm
    assert "codetr::ms_deform_attn_backward(" in nm
    ldd = subprocess.run(["ldd", DROPIN], capture_output=True, text=True).stdout
    assert "libmsda_b200.so" in ldd


_CHILD = r"""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
torch.ops.load_library(DROPIN)     # what codetr/__init__.py:8-11 does with codetr_cpp_extension.so
assert "codetr_b200" not in sys.modules
op = torch.ops.codetr.multi_scale_deformable_attention
worst = 0.0
for name in ("ref_seed3", "ref_forward_small", "edge_borders", "codino_enc_tiny", "codino_dec_tiny", "odd_dims"):
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    dev = lambda k, dt=None: (torch.from_numpy(z[k]).cuda() if dt is None else torch.from_numpy(z[k]).cuda().to(dt))
    out = op(dev("value"), dev("spatial_shapes"), dev("level_start_index"), dev("sampling_loc"), dev("attn_weight"), 64)
    ref = z["out_f32"]
    err = float(np.linalg.norm(out.cpu().numpy() - ref) / np.linalg.norm(ref))
    worst = max(worst, err)
    out16 = op(dev("value", torch.float16), dev("spatial_shapes"), dev("level_start_index"), dev("sampling_loc", torch.float16),
               dev("attn_weight", torch.float16), 64)
    assert out16.dtype == torch.float16
try:
    op(dev("value").transpose(2, 3), dev("spatial_shapes"), dev("level_start_index"), dev("sampling_loc"), dev("attn_weight"), 64)
    raise SystemExit("non-contiguous input was accepted")
except RuntimeError as e:
    assert "contiguous" in str(e)
# backward through the reference's own binding + this repo's adapter
z = np.load(os.path.join(ROOT, "tests", "golden", "codino_dec_tiny.npz"))
v, lc, aw = dev("value").double(), dev("sampling_loc").double(), dev("attn_weight").double()
gv, gl, gw = torch.zeros_like(v), torch.zeros_like(lc), torch.zeros_like(aw)
torch.ops.codetr.multi_scale_deformable_attention_backward(v, dev("spatial_shapes"), dev("level_start_index"), lc, aw,
                                                            torch.from_numpy(z["grad_out"]).cuda(), gv, gl, gw, 64)
assert float((gv.cpu() - torch.from_numpy(z["grad_value"])).abs().max()) < 1e-10
assert float((gw.cpu() - torch.from_numpy(z["grad_weight"])).abs().max()) < 1e-10
print("DROPIN_OK", worst)
assert worst <= 1e-5
"""


@needs_dropin
@pytest.mark.gpu
def test_dropin_op_matches_golden_in_fresh_process(cuda_device):
    code = f"ROOT = {ROOT!r}\nDROPIN = {DROPIN!r}\n" + _CHILD
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "DROPIN_OK" in out.stdout
