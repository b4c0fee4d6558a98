"""GPU tier (run on the B200 box with ``-m gpu``): parity of the CUDA path, called through the C ABI
(``libmsda_b200.so``), against the CPU oracle and the committed golden vectors.

Gates (BASELINE.json north_star):
  fp32      ||out - ref||_2 / ||ref||_2 <= 1e-5 against the fp32 reference on identical inputs
  fp16/bf16 max|out - ref32| / max|ref32| <= 2e-3, ref32 = fp32 reference on the upcast inputs
plus the reference's own per-test thresholds (tests/test_multi_scale_deformable_attention.py).
"""
import os

import numpy as np
import pytest
import torch

import codetr_b200 as cb
import oracle
from codetr_b200 import workloads as W
from golden_cases import ARRAY_KEYS, cases
from parity import BF16_MAX_REL, FP32_REL_L2, HALF_MAX_REL, bf16_ulp_errors, max_abs, max_rel, rel_l2

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = cases()
TORCH_DT = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16, "f64": torch.float64}
FLAG_SETS = {
    "default": 0,
    "generic": cb.FLAG_FORCE_GENERIC,
    "linear": cb.FLAG_LINEAR_ORDER,
    "fhfma": cb.FLAG_MATH_FHFMA,
    "exact": cb.FLAG_MATH_EXACT,
    "staged": cb.FLAG_STAGE_TMA,
    "staged+fhfma": cb.FLAG_STAGE_TMA | cb.FLAG_MATH_FHFMA,
    "head-major": cb.FLAG_HEAD_MAJOR,
}


def load_case(case):
    z = np.load(os.path.join(GOLDEN, case.name + ".npz"))
    arrs = {k: z[k] for k in ARRAY_KEYS} if case.store_inputs else case.build()
    return arrs, z


def to_dev(arrs, dtype, device):
    t = {k: torch.from_numpy(np.ascontiguousarray(arrs[k])) for k in ARRAY_KEYS}
    out = {}
    for k, v in t.items():
        out[k] = v.to(device) if v.dtype == torch.int64 else v.to(device=device, dtype=dtype)
    return out


def run_op(arrs, dtype, device, flags=0, step=64):
    d = to_dev(arrs, dtype, device)
    out = cb.multi_scale_deformable_attention(d["value"], d["spatial_shapes"], d["level_start_index"], d["sampling_loc"],
                                              d["attn_weight"], step, flags=flags)
    torch.cuda.synchronize()
    return out, d


def ref32_of(d):
    """fp32 reference on the (possibly 16-bit-rounded) inputs actually given to the kernel."""
    return oracle.forward_c(d["value"].float().cpu().numpy(), d["spatial_shapes"].cpu().numpy(),
                            d["level_start_index"].cpu().numpy(), d["sampling_loc"].float().cpu().numpy(),
                            d["attn_weight"].float().cpu().numpy())


# ----------------------------------------------------------------------------------------------
# golden vectors
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("flagset", ["default", "generic", "linear", "staged", "head-major"])
@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_fp32_matches_golden(case, flagset, cuda_device):
    arrs, z = load_case(case)
    out, _ = run_op(arrs, torch.float32, cuda_device, FLAG_SETS[flagset])
    assert out.dtype == torch.float32 and tuple(out.shape) == z["out_f32"].shape
    assert rel_l2(out.cpu().numpy(), z["out_f32"]) <= FP32_REL_L2
    assert rel_l2(out.cpu().numpy(), z["out_f64"]) <= FP32_REL_L2


@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_fp64_matches_golden(case, cuda_device):
    arrs, z = load_case(case)
    out, _ = run_op(arrs, torch.float64, cuda_device)
    assert out.dtype == torch.float64
    assert max_rel(out.cpu().numpy(), z["out_f64"]) < 1e-13


@pytest.mark.parametrize("flagset", ["default", "generic", "linear", "fhfma", "exact", "staged", "staged+fhfma", "head-major"])
@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_half_matches_fp32_reference(case, dt, flagset, cuda_device):
    arrs, _ = load_case(case)
    out, d = run_op(arrs, TORCH_DT[dt], cuda_device, FLAG_SETS[flagset])
    ref = ref32_of(d)
    assert out.dtype == TORCH_DT[dt]
    got = out.float().cpu().numpy()
    err = max_rel(got, ref)
    if dt == "f16":
        assert err <= HALF_MAX_REL, f"{case.name} {dt} {flagset}: {err:.3e}"
    elif "fhfma" not in flagset:
        # bf16: one output rounding (see parity.BF16_MAX_REL) -- fp32 arithmetic inside, so every element
        # is the correctly rounded bf16 value or its neighbour
        assert err <= BF16_MAX_REL, f"{case.name} {dt} {flagset}: {err:.3e}"
        assert bf16_ulp_errors(got, ref) <= 1.0 + 1e-3 or max_abs(got, ref) <= 1e-6
    else:
        # bf16 + FHFMA also rounds the combined weights to 8 bits; explicit opt-in only
        assert err <= 3 * BF16_MAX_REL, f"{case.name} {dt} {flagset}: {err:.3e}"


# ----------------------------------------------------------------------------------------------
# the reference's own tests, restated against this implementation
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_reference_forward_test(dtype, cuda_device):
    """tests/test_multi_scale_deformable_attention.py:14-62 (random inputs, opcheck, rtol 1e-2 / atol 1e-3)."""
    torch.manual_seed(0)
    bs, heads, queries, dim, levels, points, h, w = 2, 4, 8, 16, 3, 4, 32, 32
    shapes = torch.tensor([[h, w], [h // 2, w // 2], [h // 4, w // 4]], device=cuda_device, dtype=torch.int64)
    lsi = torch.tensor([0, h * w, h * w + (h // 2) * (w // 2)], device=cuda_device, dtype=torch.int64)
    value = torch.rand(bs, int((shapes[:, 0] * shapes[:, 1]).sum()), heads, dim, device=cuda_device, dtype=dtype)
    loc = torch.rand(bs, queries, heads, levels, points, 2, device=cuda_device, dtype=dtype)
    aw = torch.rand(bs, queries, heads, levels, points, device=cuda_device, dtype=dtype)
    args = (value, shapes, lsi, loc, aw, 2)
    # the reference runs the full default opcheck suite on its op (tests:44): schema, autograd registration,
    # fake tensor, AOT dispatch
    torch.library.opcheck(torch.ops.codetr.multi_scale_deformable_attention.default, args)
    if dtype == torch.float32:
        g_args = (value.clone().requires_grad_(True), shapes, lsi, loc.clone().requires_grad_(True), aw.clone().requires_grad_(True), 2)
        torch.library.opcheck(torch.ops.codetr.multi_scale_deformable_attention.default, g_args)
    out = torch.ops.codetr.multi_scale_deformable_attention(*args)
    ref = oracle.forward_grid_sample(value.cpu().float(), shapes.cpu(), loc.cpu().float(), aw.cpu().float())
    assert out.shape == (bs, queries, heads * dim)
    assert not torch.all(out == 0)
    assert out.device == cuda_device
    torch.testing.assert_close(out.cpu().float(), ref, rtol=1e-2, atol=1e-3)


def _seed3(device, dtype):
    z = np.load(os.path.join(GOLDEN, "ref_seed3.npz"))
    d = to_dev({k: z[k] for k in ARRAY_KEYS}, dtype, device)
    return d, z


def test_reference_equal_with_pytorch_double(cuda_device):
    """tests:246-283.  The reference asserts abs < 1e-18 and rel < 1e-15 between its kernel and
    grid_sample; abs < 1e-18 is below one fp64 ulp of these 5e-3-sized outputs (8.7e-19), so it only
    holds when both sides round identically -- we assert rel < 1e-15 and abs <= 2 ulp."""
    d, z = _seed3(cuda_device, torch.float64)
    out = torch.ops.codetr.multi_scale_deformable_attention(d["value"], d["spatial_shapes"], d["level_start_index"],
                                                            d["sampling_loc"], d["attn_weight"], 2).cpu().numpy()
    ref = z["out_f64"]
    assert (np.abs(out - ref) / np.abs(ref)).max() < 1e-15
    assert np.abs(out - ref).max() < 2e-18


def test_reference_equal_with_pytorch_float(cuda_device):
    """tests:286-320: abs < 1e-9, per-element rel < 1e-6."""
    d, z = _seed3(cuda_device, torch.float32)
    out = torch.ops.codetr.multi_scale_deformable_attention(d["value"], d["spatial_shapes"], d["level_start_index"],
                                                            d["sampling_loc"], d["attn_weight"], 2).cpu().numpy()
    ref = z["out_f32"]
    assert np.abs(out - ref).max() < 1e-9
    assert (np.abs(out - ref) / np.abs(ref)).max() < 1e-6


def test_reference_equal_with_pytorch_half(cuda_device):
    """tests:323-364: both sides fp16 there (abs < 1e-5, rel < 1e-2); here against the fp32 reference."""
    d, _ = _seed3(cuda_device, torch.float16)
    out = torch.ops.codetr.multi_scale_deformable_attention(d["value"], d["spatial_shapes"], d["level_start_index"],
                                                            d["sampling_loc"], d["attn_weight"], 2).float().cpu().numpy()
    ref = ref32_of(d)
    assert np.abs(out - ref).max() < 1e-5
    assert (np.abs(out - ref) / np.abs(ref)).max() < 1e-2


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_reference_mid_shape(dtype, cuda_device):
    """tests:417-501 (BASELINE configs[0]): rtol/atol 1e-5/1e-6 in fp32, 1e-2/1e-3 in fp16."""
    case = [c for c in CASES if c.name == "ref_mid"][0]
    arrs, z = load_case(case)
    out, d = run_op(arrs, dtype, cuda_device, step=2)
    if dtype == torch.float32:
        torch.testing.assert_close(out.cpu(), torch.from_numpy(z["out_f32"]), rtol=1e-5, atol=1e-6)
    else:
        torch.testing.assert_close(out.float().cpu(), torch.from_numpy(ref32_of(d)), rtol=1e-2, atol=1e-3)


# ----------------------------------------------------------------------------------------------
# boundary behaviour of the operator API
# ----------------------------------------------------------------------------------------------
def _small(device, dtype=torch.float32, bs=2):
    wl = W.Workload(name="t", shapes=tuple(W.pyramid_shapes(64, 96)), num_queries=0, batch=bs, kind="encoder", seed=77)
    inp = W.make_inputs(wl, out_of_range_frac=0.03)
    arrs = dict(value=inp.value, spatial_shapes=inp.spatial_shapes, level_start_index=inp.level_start_index,
                sampling_loc=inp.sampling_loc, attn_weight=inp.attn_weight)
    return to_dev(arrs, dtype, device), inp


def test_rejects_non_contiguous_and_bad_step(cuda_device):
    d, _ = _small(cuda_device, bs=4)
    with pytest.raises(RuntimeError, match="contiguous"):
        torch.ops.codetr.multi_scale_deformable_attention(d["value"].transpose(2, 3), d["spatial_shapes"],
                                                          d["level_start_index"], d["sampling_loc"], d["attn_weight"], 64)
    with pytest.raises(RuntimeError, match="im2col_step"):
        torch.ops.codetr.multi_scale_deformable_attention(d["value"], d["spatial_shapes"], d["level_start_index"],
                                                          d["sampling_loc"], d["attn_weight"], 3)
    with pytest.raises(RuntimeError):
        torch.ops.codetr.multi_scale_deformable_attention(d["value"], d["spatial_shapes"], d["level_start_index"],
                                                          d["sampling_loc"].half(), d["attn_weight"], 64)


def test_cpu_tensors_have_no_kernel(cuda_device):
    """Like the reference's library, only the CUDA key is implemented: no CPU fallback."""
    d, _ = _small(cuda_device)
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.codetr.multi_scale_deformable_attention(*(d[k].cpu() for k in ARRAY_KEYS), 64)


def test_batch_larger_than_im2col_step(cuda_device):
    d, _ = _small(cuda_device, bs=4)
    a = torch.ops.codetr.multi_scale_deformable_attention(*(d[k] for k in ARRAY_KEYS), 2)
    b = torch.ops.codetr.multi_scale_deformable_attention(*(d[k] for k in ARRAY_KEYS), 64)
    assert torch.equal(a, b)
    assert rel_l2(a.cpu().numpy(), ref32_of(d)) <= FP32_REL_L2


def test_empty_inputs(cuda_device):
    d, _ = _small(cuda_device)
    out = torch.ops.codetr.multi_scale_deformable_attention(d["value"], d["spatial_shapes"], d["level_start_index"],
                                                            d["sampling_loc"][:, :0].contiguous(),
                                                            d["attn_weight"][:, :0].contiguous(), 64)
    assert tuple(out.shape) == (2, 0, 256)
    out = torch.ops.codetr.multi_scale_deformable_attention(d["value"][:0].contiguous(), d["spatial_shapes"],
                                                            d["level_start_index"], d["sampling_loc"][:0].contiguous(),
                                                            d["attn_weight"][:0].contiguous(), 64)
    assert tuple(out.shape) == (0, 129, 256)


def test_output_fully_overwritten_and_deterministic(cuda_device):
    d, _ = _small(cuda_device, torch.float16)
    outs = []
    for fill in (float("nan"), 123.0):
        out = torch.full((2, 129, 256), fill, dtype=torch.float16, device=cuda_device)
        cb.forward_into(*(d[k] for k in ARRAY_KEYS), out)
        outs.append(out)
    torch.cuda.synchronize()
    assert not torch.isnan(outs[0]).any()
    assert torch.equal(outs[0], outs[1])
    lin = cb.multi_scale_deformable_attention(*(d[k] for k in ARRAY_KEYS), flags=cb.FLAG_LINEAR_ORDER)
    assert torch.equal(outs[0], lin)  # query order is a scheduling choice only: bit-identical results


def test_non_finite_locations_are_skipped(cuda_device):
    d, _ = _small(cuda_device)
    loc = d["sampling_loc"].clone()
    loc[0, 0, 0, 0, 0, 0] = float("nan")
    loc[0, 1, 1, 1, 1, 1] = float("inf")
    loc[1, 2, 2, 2, 2, 0] = -1e30
    out = cb.multi_scale_deformable_attention(d["value"], d["spatial_shapes"], d["level_start_index"], loc, d["attn_weight"])
    ref = oracle.forward_c(d["value"].cpu().numpy(), d["spatial_shapes"].cpu().numpy(), d["level_start_index"].cpu().numpy(),
                           loc.cpu().numpy(), d["attn_weight"].cpu().numpy())
    assert torch.isfinite(out).all()
    assert rel_l2(out.cpu().numpy(), ref) <= FP32_REL_L2


def test_plugin_enqueue_external_stream(cuda_device):
    """Emulates DeformableAttentionPlugin::enqueue (deformable_attention_plugin.cpp:285-355): raw
    pointers, device int64 shapes, caller-owned un-zeroed output, external non-default stream."""
    for dtype, trt in ((torch.float32, cb.ops.TRT_FLOAT), (torch.float16, cb.ops.TRT_HALF), (torch.bfloat16, cb.ops.TRT_BF16)):
        d, _ = _small(cuda_device, dtype)
        out = torch.full((2, 129, 256), float("nan"), dtype=dtype, device=cuda_device)
        stream = torch.cuda.Stream(device=cuda_device)
        stream.wait_stream(torch.cuda.current_stream(cuda_device))
        rc = cb.plugin_enqueue(d["value"].shape, d["sampling_loc"].shape, trt, [d[k].data_ptr() for k in ARRAY_KEYS],
                               out.data_ptr(), stream.cuda_stream)
        assert rc == 0
        stream.synchronize()
        ref = ref32_of(d)
        if dtype == torch.float32:
            assert rel_l2(out.cpu().numpy(), ref) <= FP32_REL_L2
        else:
            assert max_rel(out.float().cpu().numpy(), ref) <= (HALF_MAX_REL if dtype == torch.float16 else BF16_MAX_REL)
    # unsupported TensorRT dtype (kINT8 = 2) -> non-zero status, like enqueue's `return 1`
    assert cb.plugin_enqueue(d["value"].shape, d["sampling_loc"].shape, 2, [d[k].data_ptr() for k in ARRAY_KEYS],
                             out.data_ptr(), 0) != 0


def test_cuda_graph_capture(cuda_device):
    """The launcher neither synchronises nor allocates nor reads device shapes on the host, so it can
    be stream-captured (trtexec --useCudaGraph, README.md:193)."""
    d, _ = _small(cuda_device, torch.float16)
    out = torch.zeros((2, 129, 256), dtype=torch.float16, device=cuda_device)
    eager = cb.multi_scale_deformable_attention(*(d[k] for k in ARRAY_KEYS))
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(device=cuda_device)
    s.wait_stream(torch.cuda.current_stream(cuda_device))
    with torch.cuda.stream(s):
        cb.forward_into(*(d[k] for k in ARRAY_KEYS), out)  # warm-up outside capture
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            cb.forward_into(*(d[k] for k in ARRAY_KEYS), out)
    out.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, eager)


def test_host_forward_end_to_end(cuda_device):
    _, inp = _small(cuda_device)
    host = {k: torch.from_numpy(getattr(inp, k)).pin_memory() for k in ARRAY_KEYS}
    hf = cb.HostForward(cuda_device)
    out = hf(*(host[k] for k in ARRAY_KEYS))
    ref = oracle.forward_c(inp.value, inp.spatial_shapes, inp.level_start_index, inp.sampling_loc, inp.attn_weight)
    assert not out.is_cuda
    assert rel_l2(out.numpy(), ref) <= FP32_REL_L2
    h2d, d2h = hf.bytes_moved(*(host[k] for k in ARRAY_KEYS))
    assert d2h == out.numel() * 4 and h2d > d2h


def _module_producers(spatial_shapes, reference_points, sampling_offsets, attn_logits, num_points):
    """The calling module's own PyTorch ops, in the tensors' dtype
    (codetr/multi_scale_deformable_attention.py:180-200)."""
    bs, nq, heads, levels, points, _ = sampling_offsets.shape
    w = attn_logits.reshape(bs, nq, heads, levels * points).softmax(-1).view(bs, nq, heads, levels, points)
    if reference_points.shape[-1] == 2:
        norm = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1)
        loc = reference_points[:, :, None, :, None, :] + sampling_offsets / norm[None, None, None, :, None, :]
    else:
        loc = reference_points[:, :, None, :, None, :2] + sampling_offsets / num_points * reference_points[:, :, None, :, None, 2:] * 0.5
    return loc.contiguous(), w.contiguous()


@pytest.mark.parametrize("flags", [0, cb.FLAG_FORCE_GENERIC], ids=["vector", "generic"])
@pytest.mark.parametrize("kind,q", [("encoder", 0), ("decoder", 37)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_fused_producers(kind, q, dtype, flags, cuda_device):
    """Opt-in fused entry (softmax + sampling-location arithmetic inside the kernel) against the unfused
    pipeline: the module's PyTorch ops in the tensor dtype, then the fp32 reference of the op."""
    wl = W.Workload(name="t", shapes=tuple(W.pyramid_shapes(128, 192)), num_queries=q, batch=2, kind=kind, seed=5)
    inp = W.make_inputs(wl)
    dev = lambda a: torch.from_numpy(a).to(cuda_device)
    cast = lambda a: dev(a).to(dtype)
    ref_pts, off, lg = cast(inp.reference_points), cast(inp.sampling_offsets), cast(inp.attn_logits)
    value, shapes, starts = cast(inp.value), dev(inp.spatial_shapes), dev(inp.level_start_index)
    out = cb.forward_fused(value, shapes, starts, ref_pts, off, lg, flags=flags)
    assert ("fused" in cb.last_variant()) and cb.last_variant().startswith("generic" if flags else "vec<")
    loc, w = _module_producers(shapes, ref_pts, off, lg, wl.num_points)
    assert loc.dtype == dtype and w.dtype == dtype
    ref = oracle.forward_c(value.float().cpu().numpy(), inp.spatial_shapes, inp.level_start_index,
                           loc.float().cpu().numpy(), w.float().cpu().numpy())
    got = out.float().cpu().numpy()
    if dtype == torch.float32:
        assert rel_l2(got, ref) <= FP32_REL_L2
    else:
        # the kernel rounds its intermediates where the PyTorch ops do; what is left is an occasional
        # 1-ulp difference of a softmax weight or a location, well inside the gate
        assert max_rel(got, ref) <= (HALF_MAX_REL if dtype == torch.float16 else 2 * BF16_MAX_REL)
    # and it agrees with this library's own unfused op on those PyTorch-made inputs
    unfused = cb.multi_scale_deformable_attention(value, shapes, starts, loc, w, flags=flags)
    assert max_rel(got, unfused.float().cpu().numpy()) <= (1e-5 if dtype == torch.float32 else 2 * (HALF_MAX_REL if dtype == torch.float16 else BF16_MAX_REL))


# ----------------------------------------------------------------------------------------------
# BASELINE.json configurations at full size
# ----------------------------------------------------------------------------------------------
FULL = ["r50_enc_608", "swinl_enc_1152x768", "swinl_dec_1152x768", "swinl_enc_1920x1280", "swinl_dec_1900q"]


_FULL_INPUTS = {}


def _full_inputs(name, batch):
    """Full-size synthetic inputs, generated once per (workload, batch) for the whole session."""
    key = (name, batch)
    if key not in _FULL_INPUTS:
        if len(_FULL_INPUTS) > 3:
            _FULL_INPUTS.clear()
        inp = W.make_inputs(W.CONFIGS[name], batch=batch)
        _FULL_INPUTS[key] = {k: getattr(inp, k) for k in ARRAY_KEYS}
    return _FULL_INPUTS[key]


@pytest.mark.parametrize("dt", ["f16", "bf16", "f32"])
@pytest.mark.parametrize("name", FULL)
def test_full_size_configs_against_oracle(name, dt, cuda_device):
    wl = W.CONFIGS[name]
    batch = 1
    arrs = _full_inputs(name, batch)
    before = cb.launch_count()
    out, d = run_op(arrs, TORCH_DT[dt], cuda_device)
    assert cb.launch_count() == before + 1          # one kernel per call, whatever the batch
    assert cb.last_variant().startswith(("hp<", "vec<", "small<"))     # the Co-DINO shapes take the fast kernels
    ref = ref32_of(d)
    if dt == "f32":
        assert rel_l2(out.cpu().numpy(), ref) <= FP32_REL_L2
    else:
        assert max_rel(out.float().cpu().numpy(), ref) <= (HALF_MAX_REL if dt == "f16" else BF16_MAX_REL)
    # schedule independence: linear query order and TMA-staged inputs give the same bits
    lin = cb.multi_scale_deformable_attention(*(d[k] for k in ARRAY_KEYS), flags=cb.FLAG_LINEAR_ORDER)
    assert torch.equal(out, lin)
    staged = cb.multi_scale_deformable_attention(*(d[k] for k in ARRAY_KEYS), flags=cb.FLAG_STAGE_TMA)
    if dt == "bf16":
        # bf16's default on the big shapes carries each weight as two bf16 terms (head-pair kernel, 2^-17 relative);
        # the staged path uses fp32 weights: identical to MATH_EXACT, and within one output rounding of the default
        exact = cb.multi_scale_deformable_attention(*(d[k] for k in ARRAY_KEYS), flags=cb.FLAG_MATH_EXACT)
        assert torch.equal(exact, staged)
        assert max_rel(out.float().cpu().numpy(), staged.float().cpu().numpy()) <= BF16_MAX_REL
    else:
        assert torch.equal(out, staged)
    if dt != "f32":
        # both 16-bit math modes meet the gate at full size
        for fl in (cb.FLAG_MATH_EXACT, cb.FLAG_MATH_FHFMA):
            alt = cb.multi_scale_deformable_attention(*(d[k] for k in ARRAY_KEYS), flags=fl)
            bound = HALF_MAX_REL if dt == "f16" else (BF16_MAX_REL if fl == cb.FLAG_MATH_EXACT else 3 * BF16_MAX_REL)
            assert max_rel(alt.float().cpu().numpy(), ref) <= bound


@pytest.mark.parametrize("loc_mode", ["uniform", "encoder"])
def test_full_size_properties(loc_mode, cuda_device):
    """Size-independent properties at the headline shape (no oracle needed):
    linearity in value, a constant field reproduces the in-range weight mass, zero weights give zero."""
    wl = W.CONFIGS[W.HEADLINE]
    inp = W.make_inputs(wl, loc_mode=loc_mode, pad_frac=0.0)
    d = to_dev({k: getattr(inp, k) for k in ARRAY_KEYS}, torch.float32, cuda_device)
    f = lambda v, w=d["attn_weight"]: cb.multi_scale_deformable_attention(v, d["spatial_shapes"], d["level_start_index"],
                                                                          d["sampling_loc"], w)
    v1 = d["value"]
    v2 = torch.randn_like(v1)
    lhs = f(2.5 * v1 - v2)
    rhs = 2.5 * f(v1) - f(v2)
    assert rel_l2(lhs.cpu().numpy(), rhs.cpu().numpy()) < 1e-5
    assert torch.count_nonzero(f(v1, torch.zeros_like(d["attn_weight"]))) == 0
    # constant field: each sample returns c * (bilinear mass inside the level), so with locations kept
    # half a pixel inside every level the output equals c * sum of weights = c
    loc = d["sampling_loc"].clamp(0.0, 1.0)
    shp = d["spatial_shapes"].float()
    lo = (0.5 / torch.stack([shp[:, 1], shp[:, 0]], -1))[None, None, None, :, None, :]
    loc = torch.minimum(torch.maximum(loc, lo), 1.0 - lo).contiguous()
    ones = cb.multi_scale_deformable_attention(torch.full_like(v1, 3.0), d["spatial_shapes"], d["level_start_index"], loc,
                                               d["attn_weight"])
    assert (ones - 3.0).abs().max().item() < 1e-4


def test_config5_plugin_path_batch2(cuda_device):
    """BASELINE configs[4]: Swin-L encoder at 1920x1280 (S=Q=51,150), two images per GPU, launched the way
    TensorRT's enqueue launches it: raw pointers, device int64 shapes, external stream, un-zeroed output."""
    wl = W.CONFIGS["swinl_enc_1920x1280"]
    inp = W.make_inputs(wl, batch=2)
    d = to_dev({k: getattr(inp, k) for k in ARRAY_KEYS}, torch.float16, cuda_device)
    out = torch.full((2, wl.Q, 256), float("nan"), dtype=torch.float16, device=cuda_device)
    stream = torch.cuda.Stream(device=cuda_device)
    stream.wait_stream(torch.cuda.current_stream(cuda_device))
    rc = cb.plugin_enqueue(d["value"].shape, d["sampling_loc"].shape, cb.ops.TRT_HALF, [d[k].data_ptr() for k in ARRAY_KEYS],
                           out.data_ptr(), stream.cuda_stream)
    assert rc == 0
    stream.synchronize()
    assert max_rel(out.float().cpu().numpy(), ref32_of(d)) <= HALF_MAX_REL


def test_host_pipeline_matches_synchronous_calls(cuda_device):
    _, inp = _small(cuda_device)
    host = {k: torch.from_numpy(getattr(inp, k)).pin_memory() for k in ARRAY_KEYS}
    want = cb.HostForward(cuda_device)(*(host[k] for k in ARRAY_KEYS)).clone()
    pipe = cb.HostPipeline(cuda_device, depth=3)
    tickets = [pipe.submit(*(host[k] for k in ARRAY_KEYS)) for _ in range(3)]
    for t in tickets:
        assert torch.equal(pipe.result(t), want)
    for _ in range(7):  # slots are reused safely
        t = pipe.submit(*(host[k] for k in ARRAY_KEYS))
    pipe.drain()
    assert torch.equal(pipe.result(t), want)


# ----------------------------------------------------------------------------------------------
# packed-pyramid path (workspace + 256-bit loads)
# ----------------------------------------------------------------------------------------------
def _with_workspace(d, flags=0):
    need = cb.workspace_bytes(d["value"], d["sampling_loc"])
    assert need > 0
    ws = torch.empty(need, dtype=torch.uint8, device=d["value"].device)
    out = torch.full((d["value"].shape[0], d["sampling_loc"].shape[1], d["value"].shape[2] * d["value"].shape[3]), float("nan"),
                     dtype=d["value"].dtype, device=d["value"].device)
    cb.forward_into(*(d[k] for k in ARRAY_KEYS), out, flags=flags, workspace=ws)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("case", [c for c in CASES if c.name in ("edge_borders", "codino_enc_tiny", "codino_dec_tiny")],
                         ids=lambda c: c.name)
def test_packed_path_matches_direct_path(case, dt, cuda_device, monkeypatch):
    """Same math mode -> same bits as the direct path (the corner order of the accumulation is identical),
    including the column -1 / last-column / top / bottom border cases of edge_borders."""
    monkeypatch.setenv("MSDA_B200_PACKED_RATIO", "0")  # small fixtures would not amortise the pre-pass ...
    monkeypatch.setenv("MSDA_B200_SPLIT", "1")         # ... and would take the split-points variant
    arrs, _ = load_case(case)
    d = to_dev(arrs, TORCH_DT[dt], cuda_device)
    for fl in (cb.FLAG_MATH_EXACT, cb.FLAG_MATH_FHFMA):
        before = cb.launch_count()
        packed = _with_workspace(d, fl)
        assert cb.launch_count() == before + 2 and cb.last_variant().startswith("packed<")
        direct = cb.multi_scale_deformable_attention(*(d[k] for k in ARRAY_KEYS), flags=fl | cb.FLAG_NO_PACKED)
        assert cb.last_variant().startswith("vec<")
        assert torch.equal(packed, direct)
    ref = ref32_of(d)
    exact = _with_workspace(d, cb.FLAG_MATH_EXACT)
    assert max_rel(exact.float().cpu().numpy(), ref) <= (HALF_MAX_REL if dt == "f16" else BF16_MAX_REL)


def test_packed_path_exotic_level_layout_falls_back(cuda_device, monkeypatch):
    """Overlapping levels (both start at key 0) have no packed pyramid: the packed kernel must detect it on
    the device and still return the reference's result."""
    monkeypatch.setenv("MSDA_B200_PACKED_RATIO", "0")
    monkeypatch.setenv("MSDA_B200_SPLIT", "1")
    rng = np.random.default_rng(9)
    shapes = np.array([[6, 5], [3, 4]], dtype=np.int64)
    starts = np.array([0, 0], dtype=np.int64)
    arrs = dict(value=rng.standard_normal((2, 30, 8, 32), dtype=np.float32), spatial_shapes=shapes, level_start_index=starts,
                sampling_loc=rng.uniform(-0.1, 1.1, (2, 9, 8, 2, 4, 2)).astype(np.float32),
                attn_weight=rng.random((2, 9, 8, 2, 4), dtype=np.float32))
    d = to_dev(arrs, torch.float16, cuda_device)
    out = _with_workspace(d)
    assert cb.last_variant().startswith("packed<")
    assert max_rel(out.float().cpu().numpy(), ref32_of(d)) <= HALF_MAX_REL


def test_packed_path_full_size_and_torch_op(cuda_device):
    """Headline shape: with workspaces enabled the package hands the library a scratch buffer and gets the packed path;
    bits equal the direct path; the plugin-style call with a TensorRT workspace does the same."""
    arrs = _full_inputs(W.HEADLINE, 1)
    d = to_dev(arrs, torch.float16, cuda_device)
    old = cb.set_use_workspace(True)   # honoured by the package's functional API and its Python-registered op
    try:
        out = cb.multi_scale_deformable_attention(*(d[k] for k in ARRAY_KEYS), 64)
        variant = cb.last_variant()
    finally:
        cb.set_use_workspace(old)
    direct = cb.multi_scale_deformable_attention(*(d[k] for k in ARRAY_KEYS), flags=cb.FLAG_NO_PACKED)
    assert torch.equal(out, direct)
    assert variant.startswith("packed<")
    need = cb._native.load().msda_b200_plugin_workspace_bytes(
        (__import__("ctypes").c_int64 * 4)(*d["value"].shape), (__import__("ctypes").c_int64 * 6)(*d["sampling_loc"].shape), cb.ops.TRT_HALF)
    assert need == cb.workspace_bytes(d["value"], d["sampling_loc"]) == 18414 * 8 * 128
    ws = torch.empty(need, dtype=torch.uint8, device=cuda_device)
    out2 = torch.full_like(out, float("nan"))
    s = torch.cuda.Stream(device=cuda_device)
    s.wait_stream(torch.cuda.current_stream(cuda_device))
    assert cb.plugin_enqueue(d["value"].shape, d["sampling_loc"].shape, cb.ops.TRT_HALF, [d[k].data_ptr() for k in ARRAY_KEYS],
                             out2.data_ptr(), s.cuda_stream, workspace_ptr=ws.data_ptr(), workspace_bytes=need) == 0
    s.synchronize()
    assert torch.equal(out2, direct)


@pytest.mark.parametrize("name", ["swinl_enc_1152x768", "swinl_dec_1152x768"])
def test_full_size_adversarial_locations(name, cuda_device):
    """~5 % of the locations outside [0, 1] at full size (boundary parity)."""
    wl = W.CONFIGS[name]
    inp = W.make_inputs(wl, batch=1, loc_mode="adversarial")
    assert ((inp.sampling_loc < 0) | (inp.sampling_loc > 1)).mean() > 0.03
    arrs = {k: getattr(inp, k) for k in ARRAY_KEYS}
    for dt, gate, metric in (("f32", FP32_REL_L2, rel_l2), ("f16", HALF_MAX_REL, max_rel)):
        out, d = run_op(arrs, TORCH_DT[dt], cuda_device)
        assert metric(out.float().cpu().numpy(), ref32_of(d)) <= gate


@pytest.mark.parametrize("dt", ["f32", "f16", "bf16"])
def test_small_problem_kernel_matches_general_kernel(dt, cuda_device, monkeypatch):
    """The decoder-sized launch takes msda_fwd_small (4-way point split, no tiles); forcing the general
    kernel on the same tensors must give the same values up to the order of the split reduction."""
    arrs = _full_inputs("swinl_dec_1152x768", 1)
    if dt == "f32":  # fp32 rows are 8 lanes wide: 225 un-split CTAs already cover the SMs, so force the split
        monkeypatch.setenv("MSDA_B200_SPLIT_MAX_CTAS", "100000")
    out, d = run_op(arrs, TORCH_DT[dt], cuda_device)
    assert cb.last_variant().startswith("small<")
    monkeypatch.setenv("MSDA_B200_SMALL", "0")
    monkeypatch.setenv("MSDA_B200_SPLIT", "1")
    gen = cb.multi_scale_deformable_attention(*(d[k] for k in ARRAY_KEYS))
    assert cb.last_variant().startswith("vec<")
    ref = ref32_of(d)
    gate = {"f32": 1e-5, "f16": HALF_MAX_REL, "bf16": BF16_MAX_REL}[dt]
    metric = rel_l2 if dt == "f32" else max_rel
    assert metric(out.float().cpu().numpy(), ref) <= gate
    assert metric(gen.float().cpu().numpy(), ref) <= gate


@pytest.mark.parametrize("name,batch,dt", [("swinl_enc_1152x768", 1, "f16"), ("swinl_enc_1152x768", 2, "bf16"),
                                            ("r50_enc_608", 3, "f32"), ("swinl_dec_1900q", 2, "f16")])
def test_dynamic_unit_scheduling_is_bit_identical(name, batch, dt, cuda_device, monkeypatch):
    """MSDA_B200_DYN=1: warps draw their (pass, warp-slice) units from a self-resetting device counter.  Which
    warp computes a pair does not change the arithmetic of the pair, so the output must be bit-identical to the
    strided schedule; 5,000 launches walk the 4,096-slot counter pool round more than once, and a launch under
    stream capture must fall back to the strided schedule."""
    monkeypatch.setenv("MSDA_B200_SPLIT", "1")
    d = to_dev(_full_inputs(name, batch), TORCH_DT[dt], cuda_device)
    # bf16: pin the fp32-weight arithmetic (its default on this shape is the head-pair kernel's split-weight FHFMA,
    # which the dynamically scheduled vector kernel does not implement)
    fl = cb.FLAG_MATH_EXACT if dt == "bf16" else 0
    want = cb.multi_scale_deformable_attention(*(d[k] for k in ARRAY_KEYS), flags=fl)
    assert not cb.last_variant().endswith("/dyn")
    monkeypatch.setenv("MSDA_B200_DYN", "1")
    got = cb.multi_scale_deformable_attention(*(d[k] for k in ARRAY_KEYS), flags=fl)
    assert cb.last_variant().endswith("/dyn")
    assert torch.equal(got, want)
    if name == "swinl_dec_1900q":
        call = cb.PreparedForward(*(d[k] for k in ARRAY_KEYS))
        for _ in range(5000 // batch + 1):
            call()
        assert torch.equal(call(), want)
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=cuda_device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                captured = call()
                assert not cb.last_variant().endswith("/dyn")
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(captured, want)


@pytest.mark.parametrize("queries", [5, 20000])
def test_degenerate_shapes_give_zeros(queries, cuda_device):
    """No levels or no points: the sum is empty, the output is all zeros (and still fully written) -- also at a query
    count that would otherwise take the persistent fast kernels, whose input pre-loads must not run on empty tensors."""
    v = torch.randn(2, 10, 8, 32, device=cuda_device, dtype=torch.float16)
    for levels, points in ((0, 4), (2, 0)):
        shapes = torch.tensor([[2, 3], [2, 2]][:levels], dtype=torch.int64, device=cuda_device).reshape(levels, 2)
        lsi = torch.tensor([0, 6][:levels], dtype=torch.int64, device=cuda_device)
        loc = torch.rand(2, queries, 8, levels, points, 2, device=cuda_device, dtype=torch.float16)
        w = torch.rand(2, queries, 8, levels, points, device=cuda_device, dtype=torch.float16)
        out = torch.full((2, queries, 256), float("nan"), device=cuda_device, dtype=torch.float16)
        cb.forward_into(v, shapes, lsi, loc, w, out)
        torch.cuda.synchronize()
        assert cb.last_variant().startswith("generic<")
        assert torch.count_nonzero(out) == 0 and not torch.isnan(out).any()


@pytest.mark.parametrize("dt", ["f16", "bf16", "f32"])
@pytest.mark.parametrize("shape", ["small", "head_pair_sized"])
def test_locations_at_an_odd_element_offset(shape, dt, cuda_device):
    """A contiguous `sampling_loc` that is a view one element into a larger buffer is aligned for its elements but not
    for the (x, y) pair loads of the fast kernels: it must take the element-wise kernel and give the same result as
    an aligned copy -- not a misaligned-address fault, which would poison the CUDA context for the whole process."""
    shapes = W.pyramid_shapes(64, 96) if shape == "small" else W.pyramid_shapes(384, 256)
    wl = W.Workload(name="odd", shapes=tuple(shapes), num_queries=0, batch=1, kind="encoder", seed=23)
    inp = W.make_inputs(wl, out_of_range_frac=0.05)
    d = to_dev({k: getattr(inp, k) for k in ARRAY_KEYS}, TORCH_DT[dt], cuda_device)
    want = cb.multi_scale_deformable_attention(*(d[k] for k in ARRAY_KEYS), flags=cb.FLAG_MATH_EXACT)
    assert not cb.last_variant().startswith("generic<")
    loc = d["sampling_loc"]
    buf = torch.empty(loc.numel() + 1, dtype=loc.dtype, device=cuda_device)
    view = buf[1:].view(loc.shape)
    view.copy_(loc)
    assert view.is_contiguous() and view.data_ptr() % (2 * loc.element_size()) != 0
    got = cb.multi_scale_deformable_attention(d["value"], d["spatial_shapes"], d["level_start_index"], view, d["attn_weight"],
                                              flags=cb.FLAG_MATH_EXACT)
    torch.cuda.synchronize()
    assert cb.last_variant().startswith("generic<"), cb.last_variant()
    ref = ref32_of(d)
    if dt == "f32":
        assert rel_l2(got.cpu().numpy(), ref) <= FP32_REL_L2
    else:
        assert max_rel(got.float().cpu().numpy(), ref) <= (HALF_MAX_REL if dt == "f16" else BF16_MAX_REL)
        assert max_rel(got.float().cpu().numpy(), want.float().cpu().numpy()) <= (HALF_MAX_REL if dt == "f16" else BF16_MAX_REL)


def test_concurrent_streams_and_threads(cuda_device):
    """The launcher is stateless and re-entrant: two host threads, each on its own stream, interleave calls
    (TensorRT may call enqueue from its own threads / several contexts, plugin.cpp:367)."""
    import threading

    d, _ = _small(cuda_device, torch.float16, bs=2)
    want = cb.multi_scale_deformable_attention(*(d[k] for k in ARRAY_KEYS))
    torch.cuda.synchronize()
    results, errors = {}, []

    def worker(i):
        try:
            s = torch.cuda.Stream(device=cuda_device)
            outs = []
            with torch.cuda.stream(s):
                for _ in range(50):
                    outs.append(cb.multi_scale_deformable_attention(*(d[k] for k in ARRAY_KEYS)))
            s.synchronize()
            results[i] = all(torch.equal(o, want) for o in outs)
        except Exception as exc:  # pragma: no cover
            errors.append(exc)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors and all(results.get(i) for i in range(4))


@pytest.mark.parametrize("name", ["swinl_enc_1152x768", "swinl_dec_1152x768"])
def test_programmatic_dependent_launch_keeps_stream_order(name, cuda_device):
    """MSDA_FLAG_PDL lets a call start while the previous kernel drains, but it must still see everything the
    previous kernel wrote: chain calls through global memory (each call's value is the previous call's output)
    and compare with the fully serialised chain."""
    arrs = _full_inputs(name, 1)
    d = to_dev(arrs, torch.float16, cuda_device)
    wl = W.CONFIGS[name]

    def chain(flags):
        v = d["value"]
        outs = []
        for _ in range(4):
            o = cb.multi_scale_deformable_attention(v, d["spatial_shapes"], d["level_start_index"], d["sampling_loc"],
                                                    d["attn_weight"], flags=flags)
            outs.append(o)
            if wl.Q == wl.S:  # encoder: the output has the value's shape, feed it back
                v = (o * 0.5).view(1, wl.S, 8, 32).contiguous()
        torch.cuda.synchronize()
        return outs

    plain = chain(0)
    pdl = chain(cb.FLAG_PDL)
    for a, b in zip(plain, pdl):
        assert torch.equal(a, b)


def test_python_registered_op_path_in_fresh_process(cuda_device):
    """The operators can be registered natively (csrc/_torch/codetr_b200_torch.so, the default when built) or
    from Python (MSDA_B200_PYTHON_OP=1).  The rest of this file runs on whichever is the default; this test
    runs the core operator checks on the Python registration in a fresh process."""
    import subprocess, sys

    code = """
import os, sys, numpy as np, torch
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import codetr_b200 as cb, oracle
assert cb.ops.op_registration == "python", cb.ops.op_registration
z = np.load(os.path.join(ROOT, "tests", "golden", "codino_dec_tiny.npz"))
dev = lambda k, dt=None: (torch.from_numpy(z[k]).cuda() if dt is None else torch.from_numpy(z[k]).cuda().to(dt))
args = (dev("value"), dev("spatial_shapes"), dev("level_start_index"), dev("sampling_loc"), dev("attn_weight"), 64)
torch.library.opcheck(torch.ops.codetr.multi_scale_deformable_attention.default, args)
out = torch.ops.codetr.multi_scale_deformable_attention(*args)
ref = z["out_f32"]
assert np.linalg.norm(out.cpu().numpy() - ref) / np.linalg.norm(ref) <= 1e-5
v = dev("value").double().requires_grad_(True)
o = torch.ops.codetr.multi_scale_deformable_attention(v, args[1], args[2], dev("sampling_loc").double(), dev("attn_weight").double(), 64)
o.backward(torch.from_numpy(z["grad_out"]).cuda())
assert float((v.grad.cpu() - torch.from_numpy(z["grad_value"])).abs().max()) < 1e-10
print("PYTHON_OP_OK")
"""
    env = dict(os.environ, MSDA_B200_PYTHON_OP="1")
    out = subprocess.run([sys.executable, "-c", f"ROOT = {os.path.dirname(GOLDEN)[:-6]!r}\n" + code], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0 and "PYTHON_OP_OK" in out.stdout, out.stdout + out.stderr


# ----------------------------------------------------------------------------------------------
# seeded random shapes across every dispatch boundary
# ----------------------------------------------------------------------------------------------
def _random_problem(seed):
    """Heads, channels, levels, points, level sizes, query count and batch drawn from a seed: shapes the fixed fixtures do
    not name (odd channel counts, a single level, 1-pixel levels, 8 points, 16 heads, query counts around the kernels'
    group sizes) so that every kernel family and every fallback edge sees inputs it was not tuned for."""
    rng = np.random.default_rng(1000 + seed)
    M = int(rng.choice([1, 2, 3, 4, 8, 16]))
    D = int(rng.choice([8, 16, 24, 32, 32, 32, 64]))
    L = int(rng.integers(1, 7))
    P = int(rng.choice([1, 2, 4, 4, 4, 8]))
    shapes = [(int(rng.integers(1, 40)), int(rng.integers(1, 40))) for _ in range(L)]
    S = sum(h * w for h, w in shapes)
    Q = int(rng.choice([1, 3, 4, 5, 31, 32, 33, 127, 900, 2500, 7001]))
    B = int(rng.integers(1, 4))
    starts = np.cumsum([0] + [h * w for h, w in shapes[:-1]]).astype(np.int64)
    value = rng.standard_normal((B, S, M, D)).astype(np.float32)
    loc = rng.uniform(-0.15, 1.15, size=(B, Q, M, L, P, 2)).astype(np.float32)
    edge = rng.random(loc.shape) < 0.03        # exactly on the borders of the range test (ms_deform_attn.cu:249)
    loc[edge] = rng.choice(np.array([0.0, 1.0], dtype=np.float32), size=int(edge.sum()))
    w = rng.random((B, Q, M, L, P)).astype(np.float32)
    w /= w.sum(axis=(-1, -2), keepdims=True)
    return {"value": value, "spatial_shapes": np.asarray(shapes, dtype=np.int64), "level_start_index": starts,
            "sampling_loc": loc, "attn_weight": w}


@pytest.mark.parametrize("dt", ["f32", "f16", "bf16"])
@pytest.mark.parametrize("seed", range(24))
def test_random_shapes_against_oracle(seed, dt, cuda_device):
    arrs = _random_problem(seed)
    for flags in (0, cb.FLAG_FORCE_GENERIC):
        before = cb.launch_count()
        out, d = run_op(arrs, TORCH_DT[dt], cuda_device, flags)
        assert cb.launch_count() > before
        ref = ref32_of(d)
        got = out.float().cpu().numpy()
        B, Q = arrs["sampling_loc"].shape[:2]
        assert got.shape == (B, Q, arrs["value"].shape[2] * arrs["value"].shape[3])
        if dt == "f32":
            assert rel_l2(got, ref) <= FP32_REL_L2, cb.last_variant()
        else:
            assert max_rel(got, ref) <= (HALF_MAX_REL if dt == "f16" else BF16_MAX_REL), cb.last_variant()
