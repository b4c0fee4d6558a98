"""The op in its calling context: the torch-only mirror of the reference's MultiScaleDeformableAttention
module (reference tests/test_multi_scale_deformable_attention.py:122-226), fused vs unfused producers."""
import pytest
import torch

import codetr_b200 as cb
from codetr_b200 import workloads as W



def _inputs(device, dtype, num_query, embed=256, levels=5, bs=2, ref_dim=2, img=(64, 96)):
    shapes_py = W.pyramid_shapes(*img)[:levels]
    shapes = torch.tensor(shapes_py, dtype=torch.int64, device=device)
    lsi = torch.tensor(W.level_starts(shapes_py), dtype=torch.int64, device=device)
    S = W.num_keys(shapes_py)
    g = torch.Generator(device="cpu").manual_seed(1)
    nq = S if num_query is None else num_query
    query = torch.randn(nq, bs, embed, generator=g).to(device=device, dtype=dtype)
    value = torch.randn(S, bs, embed, generator=g).to(device=device, dtype=dtype)
    ref = torch.rand(bs, nq, levels, ref_dim, generator=g).to(device=device, dtype=dtype)
    mask = torch.zeros(bs, S, dtype=torch.bool, device=device)
    mask[:, -7:] = True
    return query, value, ref, shapes, lsi, mask


def test_constructor_contract():
    with pytest.raises(ValueError):  # reference tests:182-187
        cb.MultiScaleDeformableAttention(embed_dims=256, num_heads=7)
    m = cb.MultiScaleDeformableAttention(embed_dims=256, num_heads=8, num_levels=5, value_proj_ratio=0.5)  # tests:209-226
    assert m.value_proj.out_features == 128 and m.output_proj.in_features == 128
    assert set(dict(m.named_parameters())) == {
        "sampling_offsets.weight", "sampling_offsets.bias", "attention_weights.weight", "attention_weights.bias",
        "value_proj.weight", "value_proj.bias", "output_proj.weight", "output_proj.bias"}
    # the reference's bias init: head directions on the unit L-inf ring, point p at distance p+1
    b = m.sampling_offsets.bias.view(8, 5, 4, 2)
    assert torch.allclose(b[0, 0, :, 0], torch.tensor([1.0, 2.0, 3.0, 4.0])) and torch.allclose(b[0, 0, :, 1], torch.zeros(4), atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("ref_dim,num_query", [(2, None), (4, 37)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_module_forward_fused_matches_unfused(dtype, ref_dim, num_query, cuda_device):
    torch.manual_seed(0)
    mod = cb.MultiScaleDeformableAttention(embed_dims=256, num_heads=8, num_levels=5, num_points=4, dropout=0.0).to(cuda_device, dtype)
    with torch.no_grad():  # make the learned parts non-trivial
        mod.attention_weights.weight.normal_(0, 0.05)
        mod.sampling_offsets.weight.normal_(0, 0.02)
    mod.eval()
    query, value, ref, shapes, lsi, mask = _inputs(cuda_device, dtype, num_query, ref_dim=ref_dim)
    with torch.no_grad():
        a = mod(query, value=value, key_padding_mask=mask, reference_points=ref, spatial_shapes=shapes, level_start_index=lsi)
        assert cb.last_variant().startswith(("hp<", "vec<", "small<", "generic<"))
        mod.fused_producers = True
        b = mod(query, value=value, key_padding_mask=mask, reference_points=ref, spatial_shapes=shapes, level_start_index=lsi)
        assert "fused" in cb.last_variant()
    assert a.shape == query.shape and a.dtype == dtype
    err = float((a.float() - b.float()).abs().max() / a.float().abs().max())
    assert err < (1e-5 if dtype == torch.float32 else 4e-3), err


@pytest.mark.gpu
def test_module_trains_through_registered_autograd(cuda_device):
    mod = cb.MultiScaleDeformableAttention(embed_dims=64, num_heads=8, num_levels=3, num_points=4, dropout=0.0, batch_first=True).to(cuda_device)
    query, value, ref, shapes, lsi, _ = _inputs(cuda_device, torch.float32, 50, embed=64, levels=3)
    out = mod(query.transpose(0, 1).contiguous(), value=value.transpose(0, 1).contiguous(), reference_points=ref,
              spatial_shapes=shapes, level_start_index=lsi)
    out.square().mean().backward()
    for name, p in mod.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), name
    assert mod.value_proj.weight.grad.abs().sum() > 0 and mod.sampling_offsets.weight.grad.abs().sum() > 0


def test_module_rejects_cpu_tensors():
    mod = cb.MultiScaleDeformableAttention(embed_dims=64, num_heads=8, num_levels=3)
    q = torch.randn(10, 1, 64)
    with pytest.raises(RuntimeError, match="no CPU path"):
        mod(q, value=torch.randn(21, 1, 64), reference_points=torch.rand(1, 10, 3, 2),
            spatial_shapes=torch.tensor([[4, 4], [2, 2], [1, 1]]), level_start_index=torch.tensor([0, 16, 20]))
