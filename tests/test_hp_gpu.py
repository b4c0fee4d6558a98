"""GPU tier: the head-pair kernel (msda_fwd_hp: coarse pyramid levels cached in shared memory, fine levels from
global memory) against the all-global vector kernel and the C oracle.

Both kernels run the same arithmetic in the same order for a (query, head) pair, so for every shape, dtype, math
mode and shared-memory budget the outputs must be bit-identical; parity with the reference then follows from the
vector kernel's own tests, and is re-checked here against the C oracle (oracle/msda_oracle.c, which restates
ms_deform_attn.cu:31-77 / :218-260) on shapes built to hit every branch of the new kernel.
"""
import numpy as np
import pytest
import torch

import codetr_b200 as cb
import oracle
from codetr_b200 import workloads as W
from parity import BF16_MAX_REL, HALF_MAX_REL, max_rel

pytestmark = pytest.mark.gpu

KEYS = ("value", "spatial_shapes", "level_start_index", "sampling_loc", "attn_weight")
DT = {"f16": torch.float16, "bf16": torch.bfloat16, "f32": torch.float32}


def _inputs(shapes, Q, B, seed, out_of_range=0.0, kind="encoder"):
    wl = W.Workload(name="hp_test", shapes=tuple(shapes), num_queries=Q, batch=B, kind=kind, seed=seed)
    return W.make_inputs(wl, out_of_range_frac=out_of_range)


def _dev(inp, dt, device):
    d = {}
    for k in KEYS:
        t = torch.from_numpy(np.ascontiguousarray(getattr(inp, k)))
        d[k] = t.to(device) if t.dtype == torch.int64 else t.to(device=device, dtype=dt)
    return d


def _run(d, flags=0):
    out = cb.multi_scale_deformable_attention(*(d[k] for k in KEYS), 64, flags=flags)
    torch.cuda.synchronize()
    return out, cb.last_variant()


# (pyramid, queries (0 = encoder: one per key), batch): query counts that are not multiples of 4, a batch, a pyramid
# whose coarse levels fit the shared-memory budget and one where only the last level does
SHAPES = [
    (W.pyramid_shapes(384, 256), 0, 1),
    (W.pyramid_shapes(384, 256), 0, 3),
    (W.pyramid_shapes(256, 384), 2501, 2),
    (((40, 60), (20, 30), (10, 15)), 3333, 1),          # 3 levels
    (((64, 64), (32, 32)), 4099, 1),                      # 2 levels (the pipeline's minimum)
]


@pytest.mark.parametrize("dt", ["f16", "bf16", "f32"])
@pytest.mark.parametrize("smem_kb", [0, 8, 36, 148, 200])
@pytest.mark.parametrize("shape_id", range(len(SHAPES)))
def test_head_pair_kernel_is_bit_identical_to_vector_kernel(shape_id, smem_kb, dt, cuda_device, monkeypatch):
    shapes, Q, B = SHAPES[shape_id]
    kind = "encoder" if Q == 0 else "decoder"
    inp = _inputs(shapes, Q, B, seed=11 + shape_id, out_of_range=0.05, kind=kind)
    d = _dev(inp, DT[dt], cuda_device)
    monkeypatch.setenv("MSDA_B200_SPLIT", "1")                      # the fixtures are small: keep them off the small-problem kernel
    monkeypatch.setenv("MSDA_B200_HP_MIN_QUADS_PER_WARP", "0")
    monkeypatch.setenv("MSDA_B200_HP_EXACT", "1")                   # also the exact-arithmetic instantiations (default: vector kernel)
    monkeypatch.setenv("MSDA_B200_HP_SMEM", str(smem_kb * 1024))
    for fl in (0, cb.FLAG_MATH_EXACT, cb.FLAG_MATH_FHFMA):
        want, v0 = _run(d, fl | cb.FLAG_NO_SMEM_LEVELS)
        assert v0.startswith("vec<"), v0
        before = cb.launch_count()
        got, v1 = _run(d, fl)
        assert cb.launch_count() == before + 1
        assert v1.startswith("hp<") and f"smem{smem_kb}K" in v1, v1
        if dt == "bf16" and fl == 0:
            # bf16 without a math flag: FHFMA with every weight as two bf16 terms (hi + lo, 2^-17 relative) -- not the
            # vector kernel's fp32-weight arithmetic bit for bit, but inside one output rounding of it
            assert v1.endswith("/fhfma-split"), v1
            assert max_rel(got.float().cpu().numpy(), want.float().cpu().numpy()) <= BF16_MAX_REL
            continue
        assert torch.equal(got, want), f"{v1} differs from {v0}: max abs {float((got.float() - want.float()).abs().max())}"
    # and against the C oracle (fp32 reference on the rounded inputs)
    ref = oracle.forward_c(d["value"].float().cpu().numpy(), inp.spatial_shapes, inp.level_start_index,
                           d["sampling_loc"].float().cpu().numpy(), d["attn_weight"].float().cpu().numpy())
    got, _ = _run(d, 0)
    if dt == "f32":  # fp32 (8 lanes per 128-byte corner row, 2 queries per warp): the fp32 gate, relative L2 <= 1e-5
        g64 = got.cpu().numpy().astype(np.float64)
        assert float(np.linalg.norm(g64 - ref) / np.linalg.norm(ref)) <= 1e-5
    else:
        assert max_rel(got.float().cpu().numpy(), ref) <= (HALF_MAX_REL if dt == "f16" else BF16_MAX_REL)


@pytest.mark.parametrize("dt", ["f16", "f32"])
@pytest.mark.parametrize("warps", [1, 7, 16, 25])
def test_head_pair_kernel_any_warp_count(warps, dt, cuda_device, monkeypatch):
    """The host picks the warps per CTA that balances the rounds; every count must give the same bits (including
    counts that leave whole warps without a unit)."""
    inp = _inputs(W.pyramid_shapes(256, 256), 0, 2, seed=5, out_of_range=0.1)
    d = _dev(inp, DT[dt], cuda_device)
    monkeypatch.setenv("MSDA_B200_SPLIT", "1")
    monkeypatch.setenv("MSDA_B200_HP_MIN_QUADS_PER_WARP", "0")
    monkeypatch.setenv("MSDA_B200_HP_EXACT", "1")                   # also the exact-arithmetic instantiations (default: vector kernel)
    monkeypatch.setenv("MSDA_B200_HP_SMEM", str(148 * 1024))
    want, _ = _run(d, cb.FLAG_NO_SMEM_LEVELS)
    monkeypatch.setenv("MSDA_B200_HP_WARPS", str(warps))
    got, v = _run(d, 0)
    assert v.startswith("hp<") and f"/{warps}warps/" in v, v
    assert torch.equal(got, want)


def test_head_pair_kernel_output_fully_written_and_graph_safe(cuda_device, monkeypatch):
    """Caller-owned NaN-filled output, external stream, CUDA-graph capture and replay: the kernel neither
    allocates nor synchronises, and writes every element."""
    monkeypatch.setenv("MSDA_B200_SPLIT", "1")
    monkeypatch.setenv("MSDA_B200_HP_MIN_QUADS_PER_WARP", "0")
    monkeypatch.setenv("MSDA_B200_HP_EXACT", "1")                   # also the exact-arithmetic instantiations (default: vector kernel)
    monkeypatch.setenv("MSDA_B200_HP_SMEM", str(148 * 1024))
    inp = _inputs(W.pyramid_shapes(320, 224), 0, 2, seed=9, out_of_range=0.05)
    d = _dev(inp, torch.float16, cuda_device)
    want, _ = _run(d, cb.FLAG_NO_SMEM_LEVELS)
    out = torch.full_like(want, float("nan"))
    side = torch.cuda.Stream(device=cuda_device)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        cb.forward_into(*(d[k] for k in KEYS), out)
        assert cb.last_variant().startswith("hp<")
        g = torch.cuda.CUDAGraph()
        out2 = torch.full_like(want, float("nan"))
        with torch.cuda.graph(g, stream=side):
            cb.forward_into(*(d[k] for k in KEYS), out2)
    torch.cuda.synchronize()
    assert torch.equal(out, want)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out2, want)


@pytest.mark.parametrize("dt", ["f16", "f32"])
def test_head_pair_kernel_non_finite_locations(dt, cuda_device, monkeypatch):
    """NaN / infinite sampling locations fail the reference's range test (ms_deform_attn.cu:249): the sample
    contributes nothing and nothing is read for it."""
    monkeypatch.setenv("MSDA_B200_SPLIT", "1")
    monkeypatch.setenv("MSDA_B200_HP_MIN_QUADS_PER_WARP", "0")
    monkeypatch.setenv("MSDA_B200_HP_EXACT", "1")                   # also the exact-arithmetic instantiations (default: vector kernel)
    monkeypatch.setenv("MSDA_B200_HP_SMEM", str(148 * 1024))
    inp = _inputs(W.pyramid_shapes(256, 256), 0, 1, seed=21)
    loc = inp.sampling_loc.copy()
    rng = np.random.default_rng(3)
    bad = rng.random(loc.shape[:-1]) < 0.05
    loc[bad] = rng.choice(np.array([np.nan, np.inf, -np.inf, 1e30, -1e30], dtype=loc.dtype), size=(int(bad.sum()), 1))
    d = _dev(inp, DT[dt], cuda_device)
    d["sampling_loc"] = torch.from_numpy(loc).to(device=cuda_device, dtype=DT[dt])
    want, _ = _run(d, cb.FLAG_NO_SMEM_LEVELS)
    got, v = _run(d, 0)
    assert v.startswith("hp<")
    assert not torch.isnan(got).any()
    assert torch.equal(got, want)


def test_head_pair_kernel_dynamic_scheduling_is_bit_identical(cuda_device, monkeypatch):
    """MSDA_B200_HP_DYN=1 (opt-in; measured slower): warps draw (quad, head pair) units from a self-resetting device
    counter.  Which warp computes a pair does not change its arithmetic; repeated launches re-arm the counter."""
    monkeypatch.setenv("MSDA_B200_SPLIT", "1")
    monkeypatch.setenv("MSDA_B200_HP_MIN_QUADS_PER_WARP", "0")
    inp = _inputs(W.pyramid_shapes(320, 256), 0, 3, seed=17, out_of_range=0.05)
    d = _dev(inp, torch.float16, cuda_device)
    want, _ = _run(d, cb.FLAG_NO_SMEM_LEVELS)
    monkeypatch.setenv("MSDA_B200_HP_SMEM", "0")
    monkeypatch.setenv("MSDA_B200_HP_DYN", "1")
    for _ in range(20):
        got, v = _run(d, 0)
        assert v.startswith("hp<") and v.endswith("/dyn"), v
        assert torch.equal(got, want)


@pytest.mark.parametrize("dt", ["f16", "bf16"])
def test_packed_value_pyramid_gather_is_bit_identical(dt, cuda_device, monkeypatch):
    """msda_b200_pack_value + msda_b200_forward_packed (pixel-pair packed pyramid, one 32-byte load per lane for the two
    corners of an image row -- measured no faster, kept as a documented experiment): the same bits as the op on the
    plain tensor, including the column -1 / last column / first row / last row cases of a 5 %-out-of-range input."""
    monkeypatch.setenv("MSDA_B200_SPLIT", "1")
    for shapes, Q, B in ((W.pyramid_shapes(320, 256), 0, 2), (((40, 60), (20, 30), (10, 15)), 2501, 1)):
        inp = _inputs(shapes, Q, B, seed=23, out_of_range=0.05, kind="encoder" if Q == 0 else "decoder")
        d = _dev(inp, DT[dt], cuda_device)
        for fl in (0, cb.FLAG_MATH_EXACT, cb.FLAG_MATH_FHFMA):
            want, _ = _run(d, fl | cb.FLAG_NO_SMEM_LEVELS)
            packed = cb.pack_value(d["value"], d["spatial_shapes"], d["level_start_index"])
            assert packed.numel() == d["value"].numel() * 4          # 128 bytes per (key, head): twice the plain tensor
            before = cb.launch_count()
            got = cb.forward_packed(packed, DT[dt], d["value"].shape[1], d["spatial_shapes"], d["level_start_index"],
                                    d["sampling_loc"], d["attn_weight"], flags=fl)
            torch.cuda.synchronize()
            assert cb.launch_count() == before + 1 and cb.last_variant().startswith("hp_packed<")
            assert torch.equal(got, want)
    # shapes the packed gather is not written for are refused, not mis-computed
    with pytest.raises(RuntimeError):
        cb.forward_packed(packed, DT[dt], d["value"].shape[1], d["spatial_shapes"][:1].contiguous(), d["level_start_index"][:1].contiguous(),
                          d["sampling_loc"][:, :, :, :1].contiguous(), d["attn_weight"][:, :, :, :1].contiguous())


def test_bf16_split_weight_fhfma_is_as_accurate_as_fp32_weights(cuda_device, monkeypatch):
    """bf16's default on the head-pair kernel's shapes: each corner weight enters FHFMA as two bf16 terms (hi + lo).
    Against the C oracle its worst error and its relative L2 must match the fp32-weight path's to within the split's
    own 2^-17, and stay far from the single-bf16-weight FHFMA's (opt-in, 2^-9 per product)."""
    monkeypatch.setenv("MSDA_B200_SPLIT", "1")
    monkeypatch.setenv("MSDA_B200_HP_MIN_QUADS_PER_WARP", "0")
    monkeypatch.setenv("MSDA_B200_HP_EXACT", "1")
    inp = _inputs(W.pyramid_shapes(512, 384), 0, 1, seed=31, out_of_range=0.05)
    d = _dev(inp, torch.bfloat16, cuda_device)
    ref = oracle.forward_c(d["value"].float().cpu().numpy(), inp.spatial_shapes, inp.level_start_index,
                           d["sampling_loc"].float().cpu().numpy(), d["attn_weight"].float().cpu().numpy())

    def errs(flags):
        out, v = _run(d, flags)
        g = out.float().cpu().numpy().astype(np.float64)
        return max_rel(g, ref), float(np.linalg.norm(g - ref) / np.linalg.norm(ref)), v

    e_exact, l2_exact, v_exact = errs(cb.FLAG_MATH_EXACT)
    e_split, l2_split, v_split = errs(0)
    e_one, l2_one, v_one = errs(cb.FLAG_MATH_FHFMA)
    assert v_exact.endswith("/exact") and v_split.endswith("/fhfma-split") and v_one.endswith("/fhfma"), (v_exact, v_split, v_one)
    assert e_split <= BF16_MAX_REL and e_split <= e_exact + 2.0 ** -14
    assert abs(l2_split - l2_exact) <= 1e-5
    assert l2_one > l2_split * 1.15      # the single-term FHFMA is measurably worse: that is why it is not the default
    # MSDA_B200_BF16_SPLIT=0 restores the fp32-weight default
    monkeypatch.setenv("MSDA_B200_BF16_SPLIT", "0")
    _, _, v = errs(0)
    assert v.endswith("/exact"), v


@pytest.mark.parametrize("dt", ["f16", "bf16", "f32"])
@pytest.mark.parametrize("kind,q,smem_kb", [("encoder", 0, 0), ("encoder", 0, 148), ("decoder", 2501, 0), ("decoder", 2501, 36)])
def test_head_pair_kernel_fused_producers(kind, q, smem_kb, dt, cuda_device, monkeypatch):
    """msda_b200_forward_fused on the head-pair kernel: softmax over the pair's L*4 logits and the location arithmetic
    (2-d reference points for the encoder, 4-d boxes for the decoder) in the geometry lanes.  Same helpers in the same
    order as the vector kernel's fused mode -> bit-identical to it (bf16 default: split-weight FHFMA, compared within one
    output rounding), and within the usual gate of the unfused pipeline built from the module's PyTorch ops."""
    monkeypatch.setenv("MSDA_B200_SPLIT", "1")
    monkeypatch.setenv("MSDA_B200_HP_MIN_QUADS_PER_WARP", "0")
    monkeypatch.setenv("MSDA_B200_HP_SMEM", str(smem_kb * 1024))
    monkeypatch.setenv("MSDA_B200_HP_FUSED", "1")                   # opt-in: measured slower than the vector kernel's fused mode
    wl = W.Workload(name="hp_fused", shapes=tuple(W.pyramid_shapes(256, 384)), num_queries=q, batch=2, kind=kind, seed=17)
    inp = W.make_inputs(wl)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda_device)
    cast = lambda a: dev(a).to(DT[dt])
    ref_pts, off, lg = cast(inp.reference_points), cast(inp.sampling_offsets), cast(inp.attn_logits)
    value, shapes, starts = cast(inp.value), dev(inp.spatial_shapes), dev(inp.level_start_index)
    assert ref_pts.shape[-1] == (2 if kind == "encoder" else 4)
    for fl in (0, cb.FLAG_MATH_EXACT) if dt != "f32" else (0,):
        monkeypatch.setenv("MSDA_B200_HP_EXACT", "1")
        want = cb.forward_fused(value, shapes, starts, ref_pts, off, lg, flags=fl | cb.FLAG_NO_SMEM_LEVELS)
        torch.cuda.synchronize()
        v0 = cb.last_variant()
        assert v0.startswith("vec<") and v0.endswith("/fused-producers"), v0
        before = cb.launch_count()
        got = cb.forward_fused(value, shapes, starts, ref_pts, off, lg, flags=fl)
        torch.cuda.synchronize()
        v1 = cb.last_variant()
        assert cb.launch_count() == before + 1
        assert v1.startswith("hp<") and v1.endswith("/fused-producers") and f"smem{smem_kb}K" in v1, v1
        if dt == "bf16" and fl == 0:
            assert "/fhfma-split/" in v1, v1
            assert max_rel(got.float().cpu().numpy(), want.float().cpu().numpy()) <= BF16_MAX_REL
        else:
            assert torch.equal(got, want), f"{v1} differs from {v0}: max abs {float((got.float() - want.float()).abs().max())}"
    # against the unfused pipeline: the module's PyTorch ops in the tensor dtype, then the C oracle
    from test_msda_gpu import _module_producers
    loc, w = _module_producers(shapes, ref_pts, off, lg, wl.num_points)
    ref = oracle.forward_c(value.float().cpu().numpy(), inp.spatial_shapes, inp.level_start_index,
                           loc.float().cpu().numpy(), w.float().cpu().numpy())
    g = got.float().cpu().numpy().astype(np.float64)
    if dt == "f32":
        assert float(np.linalg.norm(g - ref) / np.linalg.norm(ref)) <= 1e-5
    else:
        assert max_rel(g, ref) <= (HALF_MAX_REL if dt == "f16" else 2 * BF16_MAX_REL)
