"""GPU tier: the backward kernels through the C ABI and torch autograd, against the
fp64 golden gradients (autograd through the reference's own function) and the C oracle."""
import os

import numpy as np
import pytest
import torch

import codetr_b200 as cb
import oracle
from codetr_b200 import workloads as W
from golden_cases import ARRAY_KEYS, cases
from parity import max_abs, max_rel, rel_l2
from test_oracle import GRAD_CASES, kink_mask, load_case

pytestmark = pytest.mark.gpu
TORCH_DT = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16, "f64": torch.float64}


def run_backward(arrs, grad_out, dtype, device, flags=0):
    d = {}
    for k in ARRAY_KEYS:
        t = torch.from_numpy(np.ascontiguousarray(arrs[k]))
        d[k] = t.to(device) if t.dtype == torch.int64 else t.to(device=device, dtype=dtype)
    go = torch.from_numpy(np.ascontiguousarray(grad_out)).to(device=device, dtype=dtype)
    gv = torch.zeros_like(d["value"])
    gl = torch.full_like(d["sampling_loc"], float("nan"))
    gw = torch.full_like(d["attn_weight"], float("nan"))
    cb.backward_into(*(d[k] for k in ARRAY_KEYS), go, gv, gl, gw, flags=flags)
    torch.cuda.synchronize()
    return d, go, gv, gl, gw


@pytest.mark.parametrize("case", GRAD_CASES, ids=[c.name for c in GRAD_CASES])
def test_backward_fp64_matches_golden(case, cuda_device):
    arrs, z = load_case(case)
    _, _, gv, gl, gw = run_backward(arrs, z["grad_out"], torch.float64, cuda_device)
    assert cb.last_variant() == "bwd_generic<f64>"
    assert max_rel(gv.cpu().numpy(), z["grad_value"]) < 1e-12   # atomics: summation order varies
    assert max_rel(gw.cpu().numpy(), z["grad_weight"]) < 1e-13
    smooth = ~kink_mask(arrs)
    assert max_abs(gl.cpu().numpy()[smooth], z["grad_loc"][smooth]) < 1e-12 * np.abs(z["grad_loc"]).max()
    # on the kinks this implementation follows the reference's CUDA kernel, i.e. the C oracle
    o_gv, o_gl, o_gw = oracle.backward_c(arrs["value"].astype(np.float64), arrs["spatial_shapes"], arrs["level_start_index"],
                                         arrs["sampling_loc"].astype(np.float64), arrs["attn_weight"].astype(np.float64), z["grad_out"])
    assert max_rel(gl.cpu().numpy(), o_gl) < 1e-12


@pytest.mark.parametrize("flagset", [0, cb.FLAG_FORCE_GENERIC], ids=["vector", "generic"])
@pytest.mark.parametrize("dt", ["f32", "f16", "bf16"])
@pytest.mark.parametrize("case", GRAD_CASES, ids=[c.name for c in GRAD_CASES])
def test_backward_lower_precisions(case, dt, flagset, cuda_device):
    arrs, z = load_case(case)
    d, go, gv, gl, gw = run_backward(arrs, z["grad_out"], TORCH_DT[dt], cuda_device, flagset)
    f = lambda t: t.float().cpu().numpy()
    o_gv, o_gl, o_gw = oracle.backward_c(f(d["value"]), arrs["spatial_shapes"], arrs["level_start_index"], f(d["sampling_loc"]),
                                         f(d["attn_weight"]), f(go))
    assert not torch.isnan(gl).any() and not torch.isnan(gw).any()
    if dt == "f32":
        assert rel_l2(f(gv), o_gv) < 1e-5 and rel_l2(f(gl), o_gl) < 1e-5 and rel_l2(f(gw), o_gw) < 1e-5
    else:
        # 16-bit: gradients w.r.t. loc / weight are rounded once; grad_value is accumulated by 16-bit atomics
        # (as in the reference, whose grad_value is a half tensor), so its error grows with the fan-in
        tol = 2e-3 if dt == "f16" else 2.0 ** -7
        assert max_rel(f(gl), o_gl) < tol and max_rel(f(gw), o_gw) < tol
        assert max_rel(f(gv), o_gv) < 8 * tol


def test_backward_vector_kernel_is_used_for_codino_shapes(cuda_device):
    wl = W.Workload(name="t", shapes=tuple(W.pyramid_shapes(128, 192)), num_queries=0, batch=2, kind="encoder", seed=8)
    inp = W.make_inputs(wl, out_of_range_frac=0.05)
    arrs = {k: getattr(inp, k) for k in ARRAY_KEYS}
    g = np.random.default_rng(3).standard_normal((2, wl.Q, 256)).astype(np.float32)
    for dt in ("f32", "f16"):
        d, go, gv, gl, gw = run_backward(arrs, g, TORCH_DT[dt], cuda_device)
        assert cb.last_variant().startswith("bwd_vec<")
        f = lambda t: t.float().cpu().numpy()
        o_gv, o_gl, o_gw = oracle.backward_c(f(d["value"]), arrs["spatial_shapes"], arrs["level_start_index"], f(d["sampling_loc"]),
                                             f(d["attn_weight"]), f(go))
        tol = 1e-5 if dt == "f32" else 2e-3
        metric = rel_l2 if dt == "f32" else max_rel
        assert metric(f(gl), o_gl) < tol and metric(f(gw), o_gw) < tol
        assert metric(f(gv), o_gv) < (tol if dt == "f32" else 2e-2)


@pytest.mark.parametrize("channels", [4, 30, 32, 64, 71])
def test_gradcheck_through_registered_autograd(channels, cuda_device):
    """The reference's gradcheck (tests/test_multi_scale_deformable_attention.py:367-414: fp64, eps 1e-6,
    atol 1e-2, D in {4, 30, 32, 64, 71, 1025}) through torch.ops.codetr + the registered autograd glue."""
    torch.manual_seed(3)
    N, M, Lq, L, P = 1, 2, 2, 2, 2
    shapes = torch.as_tensor([(3, 2), (2, 1)], dtype=torch.long, device=cuda_device)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    S = int(shapes.prod(1).sum())
    value = (torch.rand(N, S, M, channels, device=cuda_device, dtype=torch.float64) * 0.01).requires_grad_(True)
    loc = torch.rand(N, Lq, M, L, P, 2, device=cuda_device, dtype=torch.float64).requires_grad_(True)
    w = torch.rand(N, Lq, M, L, P, device=cuda_device, dtype=torch.float64) + 1e-5
    w = (w / w.sum(-1, keepdim=True).sum(-2, keepdim=True)).requires_grad_(True)
    fn = lambda v, lc, aw: torch.ops.codetr.multi_scale_deformable_attention(v, shapes, lsi, lc, aw, 2)
    assert torch.autograd.gradcheck(fn, (value, loc, w), eps=1e-6, atol=1e-2, nondet_tol=1e-12)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_reference_backward_smoke(dtype, cuda_device):
    """tests/test_multi_scale_deformable_attention.py:65-119: loss.backward() gives non-zero grads of the
    right shapes for value, sampling_loc and attn_weight."""
    torch.manual_seed(0)
    bs, heads, queries, dim, levels, points, h, w = 2, 4, 8, 16, 3, 4, 32, 32
    shapes = torch.tensor([[h, w], [h // 2, w // 2], [h // 4, w // 4]], device=cuda_device, dtype=torch.int64)
    lsi = torch.tensor([0, h * w, h * w + (h // 2) * (w // 2)], device=cuda_device, dtype=torch.int64)
    value = torch.rand(bs, int((shapes[:, 0] * shapes[:, 1]).sum()), heads, dim, device=cuda_device, dtype=dtype, requires_grad=True)
    loc = torch.rand(bs, queries, heads, levels, points, 2, device=cuda_device, dtype=dtype, requires_grad=True)
    aw = torch.rand(bs, queries, heads, levels, points, device=cuda_device, dtype=dtype, requires_grad=True)
    out = torch.ops.codetr.multi_scale_deformable_attention(value, shapes, lsi, loc, aw, 2)
    out.float().sum().backward()
    for t in (value, loc, aw):
        assert t.grad is not None and t.grad.shape == t.shape and t.grad.dtype == dtype
        assert torch.count_nonzero(t.grad) > 0 and torch.isfinite(t.grad).all()
    o_gv, o_gl, o_gw = oracle.backward_c(value.detach().float().cpu().numpy(), shapes.cpu().numpy(), lsi.cpu().numpy(),
                                         loc.detach().float().cpu().numpy(), aw.detach().float().cpu().numpy(),
                                         np.ones((bs, queries, heads * dim), np.float32))
    tol = 1e-5 if dtype == torch.float32 else 4e-3
    assert max_rel(loc.grad.float().cpu().numpy(), o_gl) < tol
    assert max_rel(aw.grad.float().cpu().numpy(), o_gw) < tol
    assert max_rel(value.grad.float().cpu().numpy(), o_gv) < (tol if dtype == torch.float32 else 2e-2)
