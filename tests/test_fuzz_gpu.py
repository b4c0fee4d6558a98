"""Seeded shape fuzzing of the forward, fused-producer and backward paths against the C oracle (GPU tier).

The fixed cases elsewhere follow the reference's tests and the BASELINE configurations; this file walks 48 random
(B, M, D, L, P, level sizes, Q, dtype, flags) combinations -- head counts that do not divide a warp, channel counts
without a vector kernel, single-pixel levels, one query, P != 4, locations up to 30 % outside the image -- so that every
dispatch decision of the launcher (vector / split / small / generic kernel, tiled or linear order, L2 prefetch) is
exercised on inputs nobody tuned it for.  Gates: the north_star's (fp32 rel-L2 1e-5, fp16 2e-3, bf16 2^-8 max-normalised).
"""
import numpy as np
import pytest
import torch

import codetr_b200 as cb
import oracle
from parity import BF16_MAX_REL, FP32_REL_L2, HALF_MAX_REL, max_rel, rel_l2

pytestmark = pytest.mark.gpu

DTYPES = [(torch.float32, "rel_l2", FP32_REL_L2), (torch.float16, "max_rel", HALF_MAX_REL), (torch.bfloat16, "max_rel", BF16_MAX_REL)]
FLAGS = [0, 0, 0, cb.FLAG_LINEAR_ORDER, cb.FLAG_HEAD_MAJOR, cb.FLAG_MATH_EXACT, cb.FLAG_FORCE_GENERIC, cb.FLAG_PDL]


def _random_case(seed):
    rng = np.random.default_rng(1000 + seed)
    B = int(rng.integers(1, 4))
    M = int(rng.choice([1, 2, 3, 4, 8, 8, 8]))
    D = int(rng.choice([8, 16, 24, 32, 32, 32, 64, 128]))
    L = int(rng.integers(1, 6))
    P = int(rng.choice([1, 2, 3, 4, 4, 4, 8]))
    shapes = np.stack([rng.integers(1, 24, size=L), rng.integers(1, 24, size=L)], axis=1).astype(np.int64)
    sizes = shapes[:, 0] * shapes[:, 1]
    starts = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
    S = int(sizes.sum())
    encoder_like = bool(rng.integers(0, 2))
    Q = S if encoder_like else int(rng.integers(1, 300))
    value = rng.standard_normal((B, S, M, D)).astype(np.float32)
    loc = rng.uniform(-0.3, 1.3, size=(B, Q, M, L, P, 2)).astype(np.float32)
    w = rng.uniform(0.0, 1.0, size=(B, Q, M, L, P)).astype(np.float32)
    w /= w.sum(axis=(-1, -2), keepdims=True)
    return dict(value=value, spatial_shapes=shapes, level_start_index=starts, sampling_loc=loc, attn_weight=w), (B, S, M, D, L, P, Q)


@pytest.mark.parametrize("seed", range(48))
def test_random_shapes_against_oracle(seed, cuda_device):
    arrs, dims = _random_case(seed)
    dtype, metric, gate = DTYPES[seed % 3]
    flags = FLAGS[(seed // 3) % len(FLAGS)]
    if flags == cb.FLAG_MATH_EXACT and dtype == torch.float32:
        flags = 0
    d = {k: torch.from_numpy(v).to(cuda_device) if v.dtype == np.int64 else torch.from_numpy(v).to(device=cuda_device, dtype=dtype)
         for k, v in arrs.items()}
    out = torch.full((dims[0], dims[6], dims[2] * dims[3]), float("nan"), device=cuda_device, dtype=dtype)
    cb.forward_into(d["value"], d["spatial_shapes"], d["level_start_index"], d["sampling_loc"], d["attn_weight"], out, flags=flags)
    torch.cuda.synchronize()
    assert not torch.isnan(out).any(), f"{dims} {cb.last_variant()}: output not fully written"
    ref = oracle.forward_c(d["value"].float().cpu().numpy(), arrs["spatial_shapes"], arrs["level_start_index"],
                           d["sampling_loc"].float().cpu().numpy(), d["attn_weight"].float().cpu().numpy())
    got = out.float().cpu().numpy()
    err = rel_l2(got, ref) if metric == "rel_l2" else max_rel(got, ref)
    assert err <= gate, f"dims (B,S,M,D,L,P,Q)={dims} {dtype} flags={flags} {cb.last_variant()}: {metric}={err:.3e} > {gate:g}"


@pytest.mark.parametrize("seed", range(12))
def test_random_shapes_fused_producers_against_oracle(seed, cuda_device):
    """The same walk for the producer-fused entry point: softmax + location arithmetic in-kernel vs the C oracle's
    producers followed by its forward, fp32."""
    rng = np.random.default_rng(5000 + seed)
    B, M, D = int(rng.integers(1, 3)), int(rng.choice([2, 4, 8])), int(rng.choice([16, 32, 64]))
    L, P = int(rng.integers(1, 6)), int(rng.choice([2, 4, 4]))
    shapes = np.stack([rng.integers(1, 20, size=L), rng.integers(1, 20, size=L)], axis=1).astype(np.int64)
    sizes = shapes[:, 0] * shapes[:, 1]
    starts = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
    S, Q = int(sizes.sum()), int(rng.integers(1, 200))
    ref_dim = int(rng.choice([2, 4]))
    value = rng.standard_normal((B, S, M, D)).astype(np.float32)
    refp = rng.uniform(0.05, 0.95, size=(B, Q, L, ref_dim)).astype(np.float32)
    if ref_dim == 4:
        refp[..., 2:] = rng.uniform(0.05, 0.5, size=(B, Q, L, 2))
    off = (rng.standard_normal((B, Q, M, L, P, 2)) * 2.0).astype(np.float32)
    logits = rng.standard_normal((B, Q, M, L * P)).astype(np.float32)
    loc, w = oracle.producers_c(shapes, refp, off, logits)
    want = oracle.forward_c(value, shapes, starts, loc, w)
    t = lambda a: torch.from_numpy(a).to(cuda_device)
    got = cb.forward_fused(t(value), t(shapes), t(starts), t(refp), t(off), t(logits))
    torch.cuda.synchronize()
    err = rel_l2(got.cpu().numpy(), want)
    assert err <= 2e-5, f"(B,S,M,D,L,P,Q,ref_dim)={(B, S, M, D, L, P, Q, ref_dim)} {cb.last_variant()}: rel_l2={err:.3e}"


@pytest.mark.parametrize("seed", range(16))
def test_random_shapes_backward_against_oracle(seed, cuda_device):
    """The same walk for the backward entry point (fp32 and fp64): the three gradients against the C oracle's.  Uniform
    random locations do not land on integer pixel coordinates, where the bilinear gradient has a kink."""
    arrs, dims = _random_case(200 + seed)
    B, S, M, D, L, P, Q = dims
    dtype, np_dt, gate = (torch.float64, np.float64, 1e-12) if seed % 2 else (torch.float32, np.float32, 2e-5)
    rng = np.random.default_rng(seed)
    go = rng.standard_normal((B, Q, M * D)).astype(np_dt)
    cast = {k: (v if v.dtype == np.int64 else v.astype(np_dt)) for k, v in arrs.items()}
    d = {k: torch.from_numpy(v).to(cuda_device) for k, v in cast.items()}
    gv = torch.zeros_like(d["value"])
    gl = torch.full_like(d["sampling_loc"], float("nan"))
    gw = torch.full_like(d["attn_weight"], float("nan"))
    cb.backward_into(d["value"], d["spatial_shapes"], d["level_start_index"], d["sampling_loc"], d["attn_weight"],
                     torch.from_numpy(go).to(cuda_device), gv, gl, gw)
    torch.cuda.synchronize()
    o_gv, o_gl, o_gw = oracle.backward_c(cast["value"], cast["spatial_shapes"], cast["level_start_index"], cast["sampling_loc"],
                                         cast["attn_weight"], go)
    for name, got, want in (("grad_value", gv, o_gv), ("grad_sampling_loc", gl, o_gl), ("grad_attn_weight", gw, o_gw)):
        g = got.cpu().numpy()
        assert not np.isnan(g).any(), f"{name} not fully written, dims={dims} {cb.last_variant()}"
        err = rel_l2(g, want)
        assert err <= gate, f"{name}: dims (B,S,M,D,L,P,Q)={dims} {dtype} {cb.last_variant()}: rel_l2={err:.3e}"
