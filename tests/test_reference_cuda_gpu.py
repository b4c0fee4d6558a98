"""GPU tier: parity against the REFERENCE ITSELF run here -- its own CUDA kernel (ms_deformable_im2col_gpu_kernel,
/root/reference/codetr/csrc/ms_deform_attn.cu:211-261, launched by ms_deform_attn_forward :958-973), compiled
unchanged for sm_100a by oracle/build_ref.py into oracle/_ref/msda_ref_cuda.so (built in the build container, shipped
to the GPU box with the snapshot).  fp32 on all five BASELINE.json configurations at full size: rel-L2 <= 1e-5
(north_star).  In fp16 the reference kernel computes positions, weights and the 20-term sum in half precision
(:40-42, :246-252), so it is compared through the fp32 truth: this kernel must be at least as close to it.
"""
import os

import numpy as np
import pytest
import torch

import codetr_b200 as cb
from codetr_b200 import workloads as W
from oracle import build_ref
from parity import FP32_REL_L2, HALF_MAX_REL

pytestmark = pytest.mark.gpu

KEYS = ("value", "spatial_shapes", "level_start_index", "sampling_loc", "attn_weight")
# BASELINE.json configs[0..4] (name, per-call batch)
CONFIGS = [("ref_test_mid_fp32", 1), ("r50_enc_608", 1), ("swinl_enc_1152x768", 1), ("swinl_dec_1152x768", 2), ("swinl_enc_1920x1280", 2)]


@pytest.fixture(scope="module")
def ref_op():
    if not build_ref.load_if_built():
        pytest.skip("oracle/_ref/msda_ref_cuda.so not built (needs /root/reference at build time)")
    return torch.ops.codetr_ref.msda_forward


def _dev(name, batch, dt, device):
    inp = W.make_inputs(W.CONFIGS[name], batch=batch, out_of_range_frac=0.05)
    d = {}
    for k in KEYS:
        t = torch.from_numpy(getattr(inp, k))
        d[k] = t.to(device) if t.dtype == torch.int64 else t.to(device=device, dtype=dt)
    return d


@pytest.mark.parametrize("name,batch", CONFIGS)
def test_fp32_matches_the_reference_cuda_kernel(name, batch, ref_op, cuda_device):
    d = _dev(name, batch, torch.float32, cuda_device)
    want = ref_op(*(d[k] for k in KEYS), 64)
    for flags in (0, cb.FLAG_FORCE_GENERIC):
        got = cb.multi_scale_deformable_attention(*(d[k] for k in KEYS), 64, flags=flags)
        torch.cuda.synchronize()
        err = float(torch.linalg.vector_norm((got - want).double()) / torch.linalg.vector_norm(want.double()))
        assert err <= FP32_REL_L2, f"{name} b{batch} flags={flags} {cb.last_variant()}: rel-L2 {err:.3e} vs the reference CUDA kernel"


@pytest.mark.parametrize("name,batch", CONFIGS[1:])
def test_fp16_at_least_as_close_to_fp32_truth_as_the_reference_cuda_kernel(name, batch, ref_op, cuda_device):
    d = _dev(name, batch, torch.float16, cuda_device)
    up = {k: (v.float() if v.dtype == torch.float16 else v) for k, v in d.items()}
    truth = ref_op(*(up[k] for k in KEYS), 64)                      # the reference kernel in fp32 on the rounded inputs
    theirs = ref_op(*(d[k] for k in KEYS), 64).float()
    ours = cb.multi_scale_deformable_attention(*(d[k] for k in KEYS), 64).float()
    torch.cuda.synchronize()
    scale = float(truth.abs().max())
    e_ours, e_theirs = float((ours - truth).abs().max()) / scale, float((theirs - truth).abs().max()) / scale
    assert e_ours <= HALF_MAX_REL, f"{name}: {e_ours:.3e} ({cb.last_variant()})"
    assert e_ours <= e_theirs + 1e-6, f"{name}: ours {e_ours:.3e} vs reference fp16 kernel {e_theirs:.3e}"
