"""The tensor-core producer of the op's ``value`` input (``msda_b200_value_proj``: Linear + masked_fill + head
split, reference multi_scale_deformable_attention.py:173-176) against plain PyTorch in fp32 on the same 16-bit
inputs.  Gate: one rounding of the 16-bit output (2^-11 relative for fp16, 2^-8 for bf16) plus one more unit
for accumulation-order effects at a rounding boundary; masked rows are exactly zero."""
import pytest
import torch
import torch.nn.functional as F

import codetr_b200 as cb

pytestmark = pytest.mark.gpu

ULP = {torch.float16: 2.0 ** -11, torch.bfloat16: 2.0 ** -8}


def _case(device, dtype, bs, keys, fin, fout, bias=True, mask=True, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(bs, keys, fin, generator=g).to(device=device, dtype=dtype)
    w = (torch.randn(fout, fin, generator=g) / fin ** 0.5).to(device=device, dtype=dtype)
    b = torch.randn(fout, generator=g).to(device=device, dtype=dtype) if bias else None
    m = None
    if mask:
        m = torch.rand(bs, keys, generator=g).to(device) < 0.2
        m[:, -1] = True
    return x, w, b, m


def _reference(x, w, b, m):
    ref = F.linear(x.float(), w.float(), None if b is None else b.float())
    return ref if m is None else ref.masked_fill(m[..., None], 0.0)


def _check(out, ref, dtype, m):
    assert out.dtype == dtype and out.shape == ref.shape
    err = (out.float() - ref).abs()
    bound = 2.0 * ULP[dtype] * ref.abs() + 1e-5  # + fp32 accumulation noise of a K=256 dot product
    worst = float((err - bound).max())
    assert worst <= 0.0, f"exceeds two output roundings by {worst:.3e}"
    if m is not None:
        assert torch.count_nonzero(out[m]) == 0


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("bs,keys", [(1, 1), (1, 127), (1, 128), (2, 129), (3, 1000), (1, 18414)])
def test_value_proj_matches_linear_masked_fill(bs, keys, dtype, cuda_device):
    x, w, b, m = _case(cuda_device, dtype, bs, keys, 256, 256)
    out = cb.value_proj(x, w, b, m)
    torch.cuda.synchronize()
    assert cb.last_variant().startswith("value_proj<") and "tcgen05" in cb.last_variant()
    _check(out, _reference(x, w, b, m), dtype, m)
    # and within the same bound of what the module would have computed with cuBLAS in 16 bits
    lib = F.linear(x, w, b).masked_fill(m[..., None], 0.0)
    assert float((out.float() - lib.float()).abs().max()) <= float(4.0 * ULP[dtype] * lib.float().abs().max() + 1e-6)


@pytest.mark.parametrize("fin,fout", [(64, 64), (128, 256), (256, 128), (192, 64), (64, 192)])
@pytest.mark.parametrize("bias,mask", [(True, False), (False, True), (False, False)])
def test_value_proj_shapes_and_optional_inputs(fin, fout, bias, mask, cuda_device):
    for dtype in (torch.float16, torch.bfloat16):
        x, w, b, m = _case(cuda_device, dtype, 2, 333, fin, fout, bias=bias, mask=mask, seed=fin + fout)
        out = cb.value_proj(x, w, b, m)
        torch.cuda.synchronize()
        _check(out, _reference(x, w, b, m), dtype, m)


@pytest.mark.parametrize("bs,keys", [(1, 100), (2, 129), (1, 40000)])
def test_value_proj_single_tile_variant_is_bit_identical(bs, keys, cuda_device, monkeypatch):
    """The default is the persistent, warp-specialised kernel; the one-tile-per-CTA kernel (TMA-stored output) stays
    selectable for A/B runs.  Same MMAs in the same order: bit-identical.  40,000 rows = 313 tiles: every persistent
    CTA walks both accumulator buffers and wraps the 5-slot ring several times."""
    for dtype in (torch.float16, torch.bfloat16):
        x, w, b, m = _case(cuda_device, dtype, bs, keys, 256, 256, seed=keys)
        monkeypatch.delenv("MSDA_B200_VPROJ_SINGLE_TILE", raising=False)
        a = cb.value_proj(x, w, b, m)
        assert cb.last_variant().endswith("/persistent")
        monkeypatch.setenv("MSDA_B200_VPROJ_SINGLE_TILE", "1")
        c = cb.value_proj(x, w, b, m)
        assert cb.last_variant().endswith("/single-tile")
        torch.cuda.synchronize()
        assert torch.equal(a, c)
        _check(a, _reference(x, w, b, m), dtype, m)


@pytest.mark.parametrize("cluster", [1, 2, 4])
@pytest.mark.parametrize("bs,keys", [(1, 100), (1, 7706), (2, 20000)])
def test_value_proj_cluster_sizes_are_bit_identical(bs, keys, cluster, cuda_device, monkeypatch):
    """Opt-in: the weight matrix reaches the CTAs of a thread-block cluster by TMA multicast (MSDA_B200_VPROJ_CLUSTER).  100 rows =
    one tile: the other CTAs of the cluster have no tile and only relay their weight slice; 7,706 rows = 61 tiles, an
    odd count; 40,000 rows = several tiles per CTA."""
    for dtype in (torch.float16, torch.bfloat16):
        x, w, b, m = _case(cuda_device, dtype, bs, keys, 256, 256, seed=keys + cluster)
        monkeypatch.setenv("MSDA_B200_VPROJ_CLUSTER", "1")
        want = cb.value_proj(x, w, b, m)
        monkeypatch.setenv("MSDA_B200_VPROJ_CLUSTER", str(cluster))
        got = cb.value_proj(x, w, b, m)
        assert f"/cluster{cluster}/" in cb.last_variant()
        torch.cuda.synchronize()
        assert torch.equal(got, want)
        _check(got, _reference(x, w, b, m), dtype, m)
    x, w, b, m = _case(cuda_device, torch.float16, 1, 300, 128, 64, seed=3)  # narrow output: slices of 32 / 16 rows
    assert torch.equal(cb.value_proj(x, w, b, m), F.linear(x, w, b).masked_fill(m[..., None], 0.0))


@pytest.mark.parametrize("fin,fout", [(256, 64), (64, 64), (128, 192), (256, 128)])
def test_value_proj_narrow_outputs_many_tiles_per_cta(fin, fout, cuda_device):
    """60,000 rows = 469 tiles on 148 persistent CTAs: every CTA reuses both accumulator buffers.  With 64 output columns
    half of the epilogue warps have no chunk and must still hand the accumulator back; 192 columns is the odd-chunk case."""
    x, w, b, m = _case(cuda_device, torch.float16, 1, 60000, fin, fout, seed=fin * 3 + fout)
    out = cb.value_proj(x, w, b, m)
    torch.cuda.synchronize()
    _check(out, _reference(x, w, b, m), torch.float16, m)
    res = torch.randn(1, 60000, fout, device=cuda_device).half()
    got = cb.output_proj(x, w, b, res)
    torch.cuda.synchronize()
    lib = F.linear(x, w, b) + res
    assert float((got.float() - lib.float()).abs().max()) <= float(4.0 * ULP[torch.float16] * lib.float().abs().max() + 1e-6)


@pytest.mark.parametrize("bs,keys,fout", [(1, 100, 256), (1, 900, 256), (2, 2000, 256), (1, 4700, 256), (1, 1000, 128)])
def test_column_split_for_few_row_tiles_is_bit_identical(bs, keys, fout, cuda_device, monkeypatch):
    """With few row tiles the launcher splits every tile's output columns over 2 or 4 CTAs (each loads only its slice of
    the weights).  Same MMAs per output element: bit-identical to the unsplit launch, for both entry points."""
    for dtype in (torch.float16, torch.bfloat16):
        x, w, b, m = _case(cuda_device, dtype, bs, keys, 256, fout, seed=keys)
        res = torch.randn(bs, keys, fout, device=cuda_device).to(dtype)
        monkeypatch.setenv("MSDA_B200_VPROJ_NSPLIT", "1")
        want_v, want_o = cb.value_proj(x, w, b, m), cb.output_proj(x, w, b, res)
        assert "/nsplit1/" in cb.last_variant()
        seen = set()
        for forced in ("2", "4", None):
            if forced is None:
                monkeypatch.delenv("MSDA_B200_VPROJ_NSPLIT")
            else:
                monkeypatch.setenv("MSDA_B200_VPROJ_NSPLIT", forced)
            got_v = cb.value_proj(x, w, b, m)
            seen.add(cb.last_variant().split("/nsplit")[1][0])
            got_o = cb.output_proj(x, w, b, res)
            torch.cuda.synchronize()
            assert torch.equal(got_v, want_v) and torch.equal(got_o, want_o)
        tiles = -(-bs * keys // 128)
        if tiles * 2 <= 148 and fout % 128 == 0:
            assert seen & {"2", "4"}, seen  # the automatic choice splits when it can
        _check(want_v, _reference(x, w, b, m), dtype, m)


def test_value_proj_output_is_the_ops_value_layout(cuda_device):
    x, w, b, m = _case(cuda_device, torch.float16, 2, 200, 256, 256)
    v = cb.value_proj(x, w, b, m, num_heads=8)
    assert tuple(v.shape) == (2, 200, 8, 32) and v.is_contiguous()
    flat = cb.value_proj(x, w, b, m)
    assert torch.equal(v.view(2, 200, 256), flat)  # deterministic, and the head split is a pure view


def test_value_proj_rejects_what_it_has_no_kernel_for(cuda_device):
    assert not cb.value_proj_supported(256, 256, torch.float32)
    assert not cb.value_proj_supported(320, 256, torch.float16) and not cb.value_proj_supported(256, 96, torch.float16)
    x, w, b, m = _case(cuda_device, torch.float16, 1, 10, 256, 256)
    with pytest.raises(RuntimeError):
        cb.value_proj(x.float(), w.float(), b.float(), m)
    with pytest.raises(RuntimeError):
        cb.value_proj(x.transpose(0, 1), w, b, m)
    with pytest.raises(RuntimeError):
        cb.value_proj(x, w, b, m.to(torch.uint8))
    empty = cb.value_proj(x[:, :0].contiguous(), w, b, None)
    assert tuple(empty.shape) == (1, 0, 256)


def test_value_proj_under_cuda_graph_and_side_stream(cuda_device):
    x, w, b, m = _case(cuda_device, torch.bfloat16, 2, 500, 256, 256)
    want = cb.value_proj(x, w, b, m)
    torch.cuda.synchronize()
    side = torch.cuda.Stream(device=cuda_device)
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            got = cb.value_proj(x, w, b, m)
    got.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(got, want)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("bs,queries,fin,fout", [(1, 1, 256, 256), (2, 129, 256, 256), (1, 18414, 256, 256), (3, 500, 128, 256), (1, 900, 256, 64)])
def test_output_proj_matches_linear_plus_residual(bs, queries, fin, fout, dtype, cuda_device):
    """``msda_b200_output_proj``: Linear rounded to 16 bits, then + residual, rounded again -- the same two roundings
    as ``output_proj(x) + identity`` in PyTorch, so the two agree to one unit of the last place of the sum."""
    x, w, b, _ = _case(cuda_device, dtype, bs, queries, fin, fout, seed=queries + fout)
    res = torch.randn(bs, queries, fout, device=cuda_device).to(dtype)
    out = cb.output_proj(x, w, b, res)
    torch.cuda.synchronize()
    assert cb.last_variant().startswith("output_proj<") and out.shape == res.shape and out.dtype == dtype
    lib = F.linear(x, w, b) + res
    ref = F.linear(x.float(), w.float(), b.float()).to(dtype).float() + res.float()
    err = (out.float() - ref).abs()
    assert float((err - (2.0 * ULP[dtype] * (ref.abs() + F.linear(x.float(), w.float(), b.float()).abs()) + 1e-5)).max()) <= 0.0
    assert float((out.float() - lib.float()).abs().max()) <= float(4.0 * ULP[dtype] * lib.float().abs().max() + 1e-6)
    # in place on the residual (the module's identity buffer may be reused)
    res2 = res.clone()
    lib_ = cb._native.load()
    rc = lib_.msda_b200_output_proj(x.data_ptr(), w.data_ptr(), b.data_ptr(), res2.data_ptr(), res2.data_ptr(), bs * queries, fin, fout,
                                    cb.ops._DTYPES[dtype], 0, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    assert torch.equal(res2, out)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_module_with_fused_value_proj(dtype, cuda_device):
    """``fused_value_proj=True`` swaps the module's Linear + masked_fill for the kernel; the layer output must
    agree with the unfused layer to 16-bit rounding."""
    from test_module import _inputs

    torch.manual_seed(0)
    mod = cb.MultiScaleDeformableAttention(embed_dims=256, num_heads=8, num_levels=5, num_points=4, dropout=0.0).to(cuda_device, dtype)
    with torch.no_grad():
        mod.attention_weights.weight.normal_(0, 0.05)
        mod.sampling_offsets.weight.normal_(0, 0.02)
        mod.value_proj.bias.normal_(0, 0.1)
    mod.eval()
    query, value, ref, shapes, lsi, mask = _inputs(cuda_device, dtype, None, bs=1)  # (n, 1, E): both memory orders coincide
    kw = dict(value=value, key_padding_mask=mask, reference_points=ref, spatial_shapes=shapes, level_start_index=lsi)
    with torch.no_grad():
        a = mod(query, **kw)
        launches = cb.launch_count()
        mod.fused_value_proj = True
        b = mod(query, **kw)
        assert cb.launch_count() == launches + 2  # value_proj + the sampling kernel
        mod.fused_output_proj = True
        c = mod(query, **kw)
        assert cb.launch_count() == launches + 5 and cb.last_variant().startswith("output_proj<")
        mod.fused_producers = True
        d = mod(query, **kw)
    assert c.shape == a.shape and d.shape == a.shape
    for other in (b, c, d):
        err = float((a.float() - other.float()).abs().max() / a.float().abs().max())
        assert err < (4e-3 if dtype == torch.float16 else 2e-2), err
