"""Host-side mirror of the reference's operator interface for the MSDA forward path.

What the reference exposes and what stands behind it here:

* ``torch.ops.codetr.multi_scale_deformable_attention(value, spatial_shapes,
  level_start_index, sampling_loc, attn_weight, im2col_step) -> Tensor``
  -- schema string of /root/reference/codetr/csrc/deformable_attention_torch.cpp:17-19,
  CUDA implementation registered at :28-31, fake (meta) kernel of
  /root/reference/codetr/ops.py:19-87.  Registered below from Python with the same
  schema; the CUDA implementation calls the C ABI ``msda_b200_forward``
  (include/msda_b200.h) through ctypes on torch's current stream.
* ``DeformableAttentionPlugin::enqueue`` (TensorRT, raw device pointers, external
  stream, device-resident int64 shapes,
  /root/reference/codetr/csrc/deformable_attention_plugin.cpp:285-355)
  -- :func:`plugin_enqueue` drives ``msda_b200_plugin_enqueue`` exactly that way.

There is no CPU implementation and no fallback: CPU tensors hit the dispatcher's
"no kernel for backend CPU" error exactly as they do with the reference's library
(only the CUDA key is registered there too), and a missing ``libmsda_b200.so``
raises at import of this module.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import torch
from torch import Tensor

from . import _native

_lib = _native.load()

BACKWARD_SCHEMA = (
    "multi_scale_deformable_attention_backward(Tensor value, Tensor spatial_shapes, Tensor level_start_index, "
    "Tensor sampling_loc, Tensor attn_weight, Tensor grad_output, Tensor(a!) grad_value, "
    "Tensor(b!) grad_sampling_loc, Tensor(c!) grad_attn_weight, int im2col_step) -> ()"
)
OP_SCHEMA = (
    "multi_scale_deformable_attention(Tensor value, Tensor spatial_shapes, "
    "Tensor level_start_index, Tensor sampling_loc, Tensor attn_weight, "
    "int im2col_step) -> Tensor"
)

_DTYPES = {
    torch.float32: _native.DTYPE_F32,
    torch.float16: _native.DTYPE_F16,
    torch.bfloat16: _native.DTYPE_BF16,
    torch.float64: _native.DTYPE_F64,
}
# nvinfer1::DataType integer values used by the plugin path
TRT_FLOAT, TRT_HALF, TRT_BF16 = 0, 1, 7
_TRT_DTYPES = {torch.float32: TRT_FLOAT, torch.float16: TRT_HALF, torch.bfloat16: TRT_BF16}

_default_flags = 0
# The packed-pyramid path (msda_b200_forward_ws) is opt-in: measured on B200 it lowers the L1 wavefront count
# (75 % -> 61 % of peak) but not the run time -- the gather is latency-bound at 32 warps/SM -- and its pre-pass
# costs ~8-11 us at the headline shape.
_use_workspace = False


def set_use_workspace(enabled: bool) -> bool:
    """Give the registered torch op a scratch buffer per call (packed-pyramid path).  Applies to the Python
    registration and to :func:`multi_scale_deformable_attention`; the native registration
    (``codetr_b200_torch.so``, ``op_registration == "native"``) always runs the library defaults -- set
    ``MSDA_B200_PYTHON_OP=1`` before importing the package to route the op through Python instead."""
    global _use_workspace
    old, _use_workspace = _use_workspace, bool(enabled)
    return old


def set_default_flags(flags: int) -> int:
    """Launch flags (``_native.FLAG_*``) applied by the registered torch op.  Like :func:`set_use_workspace` this
    reaches the Python registration only; the native registration passes ``MSDA_FLAG_DEFAULT`` (the library's
    ``MSDA_B200_*`` environment knobs still apply to both)."""
    global _default_flags
    old, _default_flags = _default_flags, int(flags)
    return old


def _check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError(f"msda_b200 call failed ({rc}): {_native.error_string(rc)}")


def _require(cond: bool, msg: str) -> None:
    # the reference raises c10::Error (a RuntimeError in Python) from AT_ASSERTM
    if not cond:
        raise RuntimeError(msg)


def _validate(value, spatial_shapes, level_start_index, sampling_loc, attn_weight) -> None:
    # contiguity and device requirements of ms_deform_attn.cu:902-912
    names = ("value", "spatial_shapes", "level_start_index", "sampling_loc", "attn_weight")
    for name, t in zip(names, (value, spatial_shapes, level_start_index, sampling_loc, attn_weight)):
        _require(t.is_contiguous(), f"{name} tensor has to be contiguous")
        _require(t.is_cuda, f"{name} must be a CUDA tensor")
        _require(t.device == value.device, f"{name} must be on the same device as value")
    # rank / dtype / extent agreement of the reference's fake kernel, ops.py:59-84
    _require(value.dim() == 4, "value must be [bs, num_keys, num_heads, dim_per_head]")
    _require(spatial_shapes.dim() == 2 and spatial_shapes.shape[1] == 2, "spatial_shapes must be [num_levels, 2]")
    _require(level_start_index.dim() == 1, "level_start_index must be [num_levels]")
    _require(sampling_loc.dim() == 6, "sampling_loc must be [bs, num_queries, num_heads, num_levels, num_points, 2]")
    _require(attn_weight.dim() == 5, "attn_weight must be [bs, num_queries, num_heads, num_levels, num_points]")
    _require(value.dtype in _DTYPES, f"unsupported dtype {value.dtype}")
    _require(value.dtype == sampling_loc.dtype == attn_weight.dtype, "value, sampling_loc, attn_weight dtypes differ")
    _require(spatial_shapes.dtype == torch.int64, "spatial_shapes must be int64")
    _require(level_start_index.dtype == torch.int64, "level_start_index must be int64")
    bs, _, heads, _ = value.shape
    levels = spatial_shapes.shape[0]
    _require(level_start_index.shape[0] == levels, "level_start_index / spatial_shapes level count differ")
    _require(
        sampling_loc.shape[0] == bs and sampling_loc.shape[2] == heads and sampling_loc.shape[3] == levels
        and sampling_loc.shape[5] == 2,
        "sampling_loc shape does not match value / spatial_shapes",
    )
    _require(tuple(attn_weight.shape) == tuple(sampling_loc.shape[:5]), "attn_weight shape does not match sampling_loc")


def workspace_bytes(value: Tensor, sampling_loc: Tensor) -> int:
    """Scratch bytes with which ``msda_b200_forward_ws`` can take the packed-pyramid path for these
    shapes (0 = the path does not apply and no workspace is needed)."""
    bs, keys, heads, chans = value.shape
    return int(_lib.msda_b200_workspace_bytes(bs, keys, heads, chans, sampling_loc.shape[3], sampling_loc.shape[1],
                                              sampling_loc.shape[4], _DTYPES[value.dtype]))


def _stream_ptr(device: torch.device, stream: Optional[int]) -> int:
    return int(stream) if stream is not None else int(torch.cuda.current_stream(device).cuda_stream)


def forward_into(
    value: Tensor,
    spatial_shapes: Tensor,
    level_start_index: Tensor,
    sampling_loc: Tensor,
    attn_weight: Tensor,
    output: Tensor,
    im2col_step: int = 64,
    flags: Optional[int] = None,
    stream: Optional[int] = None,
    workspace: Optional[Tensor] = None,
) -> Tensor:
    """``codetr::ms_deform_attn_forward_reference`` (ms_deform_attn.cu:899-956): caller-owned output.
    ``workspace``: optional uint8 CUDA scratch tensor of ``workspace_bytes(...)`` bytes (packed path)."""
    _validate(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)
    bs, keys, heads, chans = value.shape
    queries, levels, points = sampling_loc.shape[1], sampling_loc.shape[3], sampling_loc.shape[4]
    _require(output.is_contiguous() and output.is_cuda, "output tensor has to be a contiguous CUDA tensor")
    _require(tuple(output.shape) == (bs, queries, heads * chans), "output must be [bs, num_queries, num_heads*channels]")
    _require(output.dtype == value.dtype and output.device == value.device, "output dtype/device must match value")
    dev = value.device
    guard = torch.cuda.device(dev) if torch.cuda.current_device() != dev.index else None
    if guard is not None:
        guard.__enter__()
    try:
        rc = _lib.msda_b200_forward_ws(
            value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), sampling_loc.data_ptr(),
            attn_weight.data_ptr(), output.data_ptr(), 0 if workspace is None else workspace.data_ptr(),
            0 if workspace is None else workspace.numel() * workspace.element_size(), bs, keys, heads, chans, levels,
            queries, points, int(im2col_step), _DTYPES[value.dtype], _default_flags if flags is None else int(flags),
            _stream_ptr(dev, stream),
        )
    finally:
        if guard is not None:
            guard.__exit__(None, None, None)
    _check(rc)
    return output


def multi_scale_deformable_attention(
    value: Tensor,
    spatial_shapes: Tensor,
    level_start_index: Tensor,
    sampling_loc: Tensor,
    attn_weight: Tensor,
    im2col_step: int = 64,
    flags: Optional[int] = None,
) -> Tensor:
    """``codetr::ms_deform_attn_forward`` (ms_deform_attn.cu:958-973): allocates ``[bs, num_queries, M*D]``."""
    _validate(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)
    bs, _, heads, chans = value.shape
    out = torch.empty((bs, sampling_loc.shape[1], heads * chans), dtype=value.dtype, device=value.device)
    ws = None
    if _use_workspace:
        need = workspace_bytes(value, sampling_loc)
        if need:
            ws = torch.empty(need, dtype=torch.uint8, device=value.device)  # torch's caching allocator: no cudaMalloc in steady state
    return forward_into(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, out, im2col_step, flags,
                        workspace=ws)


def backward_into(
    value: Tensor,
    spatial_shapes: Tensor,
    level_start_index: Tensor,
    sampling_loc: Tensor,
    attn_weight: Tensor,
    grad_output: Tensor,
    grad_value: Tensor,
    grad_sampling_loc: Tensor,
    grad_attn_weight: Tensor,
    im2col_step: int = 64,
    flags: Optional[int] = None,
) -> None:
    """``codetr::ms_deform_attn_backward`` (ms_deform_attn.cu:975-1028): accumulates into ``grad_value``
    (zero it first, as codetr/ops.py:94-96 does), overwrites the other two gradients."""
    _validate(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)
    bs, keys, heads, chans = value.shape
    queries, levels, points = sampling_loc.shape[1], sampling_loc.shape[3], sampling_loc.shape[4]
    for name, t, like in (("grad_output", grad_output, None), ("grad_value", grad_value, value),
                          ("grad_sampling_loc", grad_sampling_loc, sampling_loc), ("grad_attn_weight", grad_attn_weight, attn_weight)):
        _require(t.is_contiguous(), f"{name} tensor has to be contiguous")
        _require(t.is_cuda and t.device == value.device, f"{name} must be a CUDA tensor on value's device")
        _require(t.dtype == value.dtype, f"{name} dtype must match value")
        if like is not None:
            _require(tuple(t.shape) == tuple(like.shape), f"{name} shape mismatch")
    _require(tuple(grad_output.shape) == (bs, queries, heads * chans), "grad_output must be [bs, num_queries, num_heads*channels]")
    with torch.cuda.device(value.device):
        rc = _lib.msda_b200_backward(
            value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), sampling_loc.data_ptr(),
            attn_weight.data_ptr(), grad_output.data_ptr(), grad_value.data_ptr(), grad_sampling_loc.data_ptr(),
            grad_attn_weight.data_ptr(), bs, keys, heads, chans, levels, queries, points, int(im2col_step),
            _DTYPES[value.dtype], _default_flags if flags is None else int(flags), _stream_ptr(value.device, None),
        )
    _check(rc)


class PreparedForward:
    """A validated, pointer-bound call (the analogue of a TensorRT execution context with its tensor
    addresses set): arguments are checked once, each ``__call__`` is a single C-ABI call on the
    given (default: torch's current) stream.  The tensors are kept alive by the object."""

    def __init__(self, value: Tensor, spatial_shapes: Tensor, level_start_index: Tensor, sampling_loc: Tensor,
                 attn_weight: Tensor, output: Optional[Tensor] = None, im2col_step: int = 64, flags: Optional[int] = None,
                 workspace: Optional[Tensor] = None):
        _validate(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)
        bs, keys, heads, chans = value.shape
        queries, levels, points = sampling_loc.shape[1], sampling_loc.shape[3], sampling_loc.shape[4]
        if output is None:
            output = torch.empty((bs, queries, heads * chans), dtype=value.dtype, device=value.device)
        _require(output.is_contiguous() and tuple(output.shape) == (bs, queries, heads * chans)
                 and output.dtype == value.dtype and output.device == value.device, "bad output tensor")
        self.tensors = (value, spatial_shapes, level_start_index, sampling_loc, attn_weight, output, workspace)
        self.output = output
        self.device = value.device
        self._args = (value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), sampling_loc.data_ptr(),
                      attn_weight.data_ptr(), output.data_ptr(), 0 if workspace is None else workspace.data_ptr(),
                      0 if workspace is None else workspace.numel() * workspace.element_size(),
                      bs, keys, heads, chans, levels, queries, points,
                      int(im2col_step), _DTYPES[value.dtype], _default_flags if flags is None else int(flags))

    def __call__(self, stream: Optional[int] = None) -> Tensor:
        rc = _lib.msda_b200_forward_ws(*self._args, _stream_ptr(self.device, stream))
        if rc != 0:
            _check(rc)
        return self.output


def forward_fused(
    value: Tensor,
    spatial_shapes: Tensor,
    level_start_index: Tensor,
    reference_points: Tensor,
    sampling_offsets: Tensor,
    attn_logits: Tensor,
    flags: Optional[int] = None,
) -> Tensor:
    """Opt-in producer-fused mode: softmax over L*P and the sampling-location arithmetic of
    /root/reference/codetr/multi_scale_deformable_attention.py:180-200 happen inside the kernel.

    ``reference_points [bs, Q, L, 2|4]``, ``sampling_offsets [bs, Q, M, L, P, 2]``,
    ``attn_logits [bs, Q, M, L*P]`` (or ``[bs, Q, M, L, P]``)."""
    for name, t in (("value", value), ("reference_points", reference_points), ("sampling_offsets", sampling_offsets),
                    ("attn_logits", attn_logits), ("spatial_shapes", spatial_shapes),
                    ("level_start_index", level_start_index)):
        _require(t.is_contiguous(), f"{name} tensor has to be contiguous")
        _require(t.is_cuda and t.device == value.device, f"{name} must be a CUDA tensor on value's device")
    _require(value.dim() == 4 and sampling_offsets.dim() == 6 and reference_points.dim() == 4, "bad ranks")
    _require(value.dtype in _DTYPES, f"unsupported dtype {value.dtype}")
    _require(value.dtype == reference_points.dtype == sampling_offsets.dtype == attn_logits.dtype, "dtypes differ")
    _require(spatial_shapes.dtype == torch.int64 and level_start_index.dtype == torch.int64, "shapes must be int64")
    bs, keys, heads, chans = value.shape
    _, queries, _, levels, points, _ = sampling_offsets.shape
    ref_dim = reference_points.shape[-1]
    _require(ref_dim in (2, 4), f"Last dim of reference_points must be 2 or 4, but get {ref_dim} instead.")
    _require(tuple(reference_points.shape) == (bs, queries, levels, ref_dim), "reference_points shape mismatch")
    _require(tuple(sampling_offsets.shape) == (bs, queries, heads, levels, points, 2), "sampling_offsets shape mismatch")
    _require(attn_logits.numel() == bs * queries * heads * levels * points, "attn_logits shape mismatch")
    out = torch.empty((bs, queries, heads * chans), dtype=value.dtype, device=value.device)
    with torch.cuda.device(value.device):
        rc = _lib.msda_b200_forward_fused(
            value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), reference_points.data_ptr(),
            sampling_offsets.data_ptr(), attn_logits.data_ptr(), out.data_ptr(), bs, keys, heads, chans, levels,
            queries, points, ref_dim, _DTYPES[value.dtype], _default_flags if flags is None else int(flags),
            _stream_ptr(value.device, None),
        )
    _check(rc)
    return out


def value_proj_supported(in_features: int, out_features: int, dtype: torch.dtype) -> bool:
    """Whether :func:`value_proj` has a tensor-core kernel for this Linear (16-bit dtype, both widths multiples
    of 64 and at most 256)."""
    return dtype in _DTYPES and bool(_lib.msda_b200_value_proj_supported(int(in_features), int(out_features), _DTYPES[dtype]))


def value_proj(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None, key_padding_mask: Optional[Tensor] = None,
               num_heads: Optional[int] = None) -> Tensor:
    """Producer of the op's ``value`` input in one tcgen05 kernel: ``masked_fill(linear(x, weight, bias),
    key_padding_mask[..., None], 0)`` (/root/reference/codetr/multi_scale_deformable_attention.py:173-176).

    ``x [bs, S, in]``, ``weight [out, in]``, ``bias [out]`` or None, ``key_padding_mask [bs, S]`` bool or None.
    Returns ``[bs, S, out]``, or ``[bs, S, num_heads, out // num_heads]`` when ``num_heads`` is given (a view:
    the kernel's output already is the op's ``[bs, S, M, D]`` layout).  Raises for shapes / dtypes without a
    kernel (see :func:`value_proj_supported`); there is no fallback inside this function."""
    _require(x.is_cuda and weight.is_cuda and x.device == weight.device, "value_proj needs CUDA tensors on one device")
    _require(x.dim() == 3 and weight.dim() == 2 and weight.shape[1] == x.shape[-1], "bad x / weight shapes")
    _require(x.is_contiguous() and weight.is_contiguous(), "x and weight have to be contiguous")
    _require(x.dtype == weight.dtype, "x and weight dtypes differ")
    bs, keys, fin = x.shape
    fout = weight.shape[0]
    _require(value_proj_supported(fin, fout, x.dtype), f"no value_proj kernel for {fin} -> {fout} in {x.dtype}")
    if bias is not None:
        _require(bias.is_cuda and bias.is_contiguous() and bias.dtype == x.dtype and tuple(bias.shape) == (fout,), "bad bias")
    if key_padding_mask is not None:
        _require(key_padding_mask.is_cuda and key_padding_mask.dtype == torch.bool and key_padding_mask.is_contiguous()
                 and tuple(key_padding_mask.shape) == (bs, keys), "key_padding_mask must be a contiguous bool [bs, S] CUDA tensor")
    out = torch.empty((bs, keys, fout), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.msda_b200_value_proj(x.data_ptr(), weight.data_ptr(), 0 if bias is None else bias.data_ptr(),
                                       0 if key_padding_mask is None else key_padding_mask.data_ptr(), out.data_ptr(),
                                       bs * keys, fin, fout, _DTYPES[x.dtype], 0, _stream_ptr(x.device, None))
    _check(rc)
    return out if num_heads is None else out.view(bs, keys, num_heads, fout // num_heads)


def output_proj(attended: Tensor, weight: Tensor, bias: Optional[Tensor], residual: Tensor) -> Tensor:
    """Consumer of the op's output in one tcgen05 kernel: ``linear(attended, weight, bias) + residual``
    (/root/reference/codetr/multi_scale_deformable_attention.py:212-218 with dropout in inference mode).

    ``attended [bs, Q, in]`` (the op's output), ``weight [out, in]``, ``bias [out]`` or None, ``residual [bs, Q, out]``.
    Same shape / dtype limits as :func:`value_proj`; raises outside them."""
    _require(attended.is_cuda and weight.is_cuda and residual.is_cuda and attended.device == weight.device == residual.device,
             "output_proj needs CUDA tensors on one device")
    _require(attended.dim() == 3 and weight.dim() == 2 and weight.shape[1] == attended.shape[-1], "bad attended / weight shapes")
    _require(attended.is_contiguous() and weight.is_contiguous() and residual.is_contiguous(), "tensors have to be contiguous")
    _require(attended.dtype == weight.dtype == residual.dtype, "dtypes differ")
    bs, queries, fin = attended.shape
    fout = weight.shape[0]
    _require(tuple(residual.shape) == (bs, queries, fout), "residual shape mismatch")
    _require(value_proj_supported(fin, fout, attended.dtype), f"no projection kernel for {fin} -> {fout} in {attended.dtype}")
    if bias is not None:
        _require(bias.is_cuda and bias.is_contiguous() and bias.dtype == attended.dtype and tuple(bias.shape) == (fout,), "bad bias")
    out = torch.empty_like(residual)
    with torch.cuda.device(attended.device):
        rc = _lib.msda_b200_output_proj(attended.data_ptr(), weight.data_ptr(), 0 if bias is None else bias.data_ptr(),
                                        residual.data_ptr(), out.data_ptr(), bs * queries, fin, fout, _DTYPES[attended.dtype], 0,
                                        _stream_ptr(attended.device, None))
    _check(rc)
    return out


def pack_value(value: Tensor, spatial_shapes: Tensor, level_start_index: Tensor) -> Tensor:
    """Plain ``[bs, S, M, 32]`` 16-bit value tensor -> pixel-pair packed pyramid (``msda_b200_pack_value``): a uint8
    tensor of ``bs * S * M * 128`` bytes for :func:`forward_packed`.  The projection kernel can write this layout
    directly (``value_proj(..., packed=True)``), which is the point; this pre-pass exists for callers that already
    hold a plain tensor and for tests."""
    _require(value.is_cuda and value.is_contiguous() and value.dim() == 4, "value has to be a contiguous CUDA tensor [bs, S, M, D]")
    _require(value.dtype in (torch.float16, torch.bfloat16) and value.shape[3] == 32, "packed layout: fp16 / bf16, D = 32")
    bs, keys, heads, chans = value.shape
    need = int(_lib.msda_b200_packed_value_bytes(bs, keys, heads, chans, _DTYPES[value.dtype]))
    _require(need > 0, "no packed form for this shape")
    packed = torch.empty(need, dtype=torch.uint8, device=value.device)
    with torch.cuda.device(value.device):
        _check(_lib.msda_b200_pack_value(value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), packed.data_ptr(),
                                         bs, keys, heads, chans, spatial_shapes.shape[0], _DTYPES[value.dtype],
                                         _stream_ptr(value.device, None)))
    return packed


def forward_packed(packed_value: Tensor, dtype: torch.dtype, num_keys: int, spatial_shapes: Tensor, level_start_index: Tensor,
                   sampling_loc: Tensor, attn_weight: Tensor, flags: Optional[int] = None, out: Optional[Tensor] = None) -> Tensor:
    """The operator on a packed value pyramid (``msda_b200_forward_packed``): same ``sampling_loc`` / ``attn_weight`` /
    output contract and the same bits as the op on the plain tensor.  ``packed_value``: uint8 buffer from
    :func:`pack_value` or ``value_proj(..., packed=True)``; ``dtype`` the element type it holds; 8 heads x 32 channels."""
    _require(packed_value.is_cuda and packed_value.is_contiguous() and packed_value.dtype == torch.uint8, "packed_value: contiguous CUDA uint8 buffer")
    for name, t in (("sampling_loc", sampling_loc), ("attn_weight", attn_weight), ("spatial_shapes", spatial_shapes),
                    ("level_start_index", level_start_index)):
        _require(t.is_contiguous() and t.is_cuda and t.device == packed_value.device, f"{name} must be a contiguous CUDA tensor on the value's device")
    _require(sampling_loc.dim() == 6 and attn_weight.dim() == 5 and sampling_loc.dtype == attn_weight.dtype == dtype, "bad sampling_loc / attn_weight")
    bs, queries, heads, levels, points, _ = sampling_loc.shape
    _require(tuple(attn_weight.shape) == (bs, queries, heads, levels, points), "attn_weight shape mismatch")
    _require(packed_value.numel() == bs * int(num_keys) * heads * 128, "packed_value size does not match bs * num_keys * heads * 128")
    if out is None:
        out = torch.empty((bs, queries, heads * 32), dtype=dtype, device=packed_value.device)
    with torch.cuda.device(packed_value.device):
        rc = _lib.msda_b200_forward_packed(packed_value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                                           sampling_loc.data_ptr(), attn_weight.data_ptr(), out.data_ptr(), bs, int(num_keys), heads, 32,
                                           levels, queries, points, _DTYPES[dtype], _default_flags if flags is None else int(flags),
                                           _stream_ptr(packed_value.device, None))
    _check(rc)
    return out


def plugin_enqueue(
    value_dims: Sequence[int],
    loc_dims: Sequence[int],
    trt_dtype: int,
    input_ptrs: Sequence[int],
    output_ptr: int,
    stream: int,
    im2col_step: int = 64,
    workspace_ptr: int = 0,
    workspace_bytes: int = 0,
) -> int:
    """Drive the library the way TensorRT drives ``DeformableAttentionPlugin::enqueue``
    (deformable_attention_plugin.cpp:285-355): dims from the tensor descriptors, five raw
    device pointers, one raw output pointer, an external ``cudaStream_t``.  Returns the
    plugin's status code (0 = success) instead of raising, like ``enqueue``."""
    vd = (ctypes.c_int64 * 4)(*[int(v) for v in value_dims])
    ld = (ctypes.c_int64 * 6)(*[int(v) for v in loc_dims])
    ins = (ctypes.c_void_p * 5)(*[int(p) for p in input_ptrs])
    outs = (ctypes.c_void_p * 1)(int(output_ptr))
    return int(_lib.msda_b200_plugin_enqueue(vd, ld, int(trt_dtype), ins, outs, int(workspace_ptr) or None, int(workspace_bytes),
                                             int(im2col_step), int(stream)))


class HostForward:
    """End-to-end call for HOST buffers (``msda_b200_forward_host``): pinned inputs are copied
    host->device into a reusable device workspace, the kernel runs, the result is copied back to a
    pinned output -- all on one stream.  This is what ``bench.py`` times as ``e2e``."""

    def __init__(self, device: torch.device):
        self.device = torch.device(device)
        self._ws: Optional[Tensor] = None

    def __call__(self, value: Tensor, spatial_shapes: Tensor, level_start_index: Tensor, sampling_loc: Tensor,
                 attn_weight: Tensor, output: Optional[Tensor] = None, im2col_step: int = 64,
                 flags: Optional[int] = None, synchronize: bool = True) -> Tensor:
        for t in (value, spatial_shapes, level_start_index, sampling_loc, attn_weight):
            _require(not t.is_cuda and t.is_contiguous(), "HostForward takes contiguous host tensors")
        bs, keys, heads, chans = value.shape
        queries, levels, points = sampling_loc.shape[1], sampling_loc.shape[3], sampling_loc.shape[4]
        dt = _DTYPES[value.dtype]
        need = int(_lib.msda_b200_host_workspace_bytes(bs, keys, heads, chans, levels, queries, points, dt))
        with torch.cuda.device(self.device):
            if self._ws is None or self._ws.numel() < need:
                self._ws = torch.empty(max(need, 256), dtype=torch.uint8, device=self.device)
            if output is None:
                output = torch.empty((bs, queries, heads * chans), dtype=value.dtype).pin_memory()
            stream = torch.cuda.current_stream(self.device)
            rc = _lib.msda_b200_forward_host(
                value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), sampling_loc.data_ptr(),
                attn_weight.data_ptr(), output.data_ptr(), self._ws.data_ptr(), self._ws.numel(), bs, keys, heads,
                chans, levels, queries, points, int(im2col_step), dt,
                _default_flags if flags is None else int(flags), int(stream.cuda_stream),
            )
            _check(rc)
            if synchronize:
                stream.synchronize()
        return output

    @staticmethod
    def bytes_moved(value: Tensor, spatial_shapes: Tensor, level_start_index: Tensor, sampling_loc: Tensor,
                    attn_weight: Tensor) -> tuple:
        h2d = sum(t.numel() * t.element_size() for t in (value, spatial_shapes, level_start_index, sampling_loc, attn_weight))
        d2h = value.shape[0] * sampling_loc.shape[1] * value.shape[2] * value.shape[3] * value.element_size()
        return int(h2d), int(d2h)


class HostPipeline:
    """Software-pipelined end-to-end calls for HOST buffers: ``depth`` slots, each with its own stream,
    device workspace and pinned output, so the host->device copy of call i+1, the kernel of call i and the
    device->host copy of call i-1 overlap (PCIe is full duplex and the copy engines run beside the SMs).
    Every call still moves all of its inputs and its result across PCIe through
    ``msda_b200_forward_host``.

        pipe = HostPipeline(device)
        t = pipe.submit(value, shapes, starts, loc, weight)     # returns immediately
        out = pipe.result(t)                                     # pinned host tensor, synchronised
    """

    def __init__(self, device: torch.device, depth: int = 3):
        self.device = torch.device(device)
        self.depth = int(depth)
        with torch.cuda.device(self.device):
            self._streams = [torch.cuda.Stream(device=self.device) for _ in range(self.depth)]
        self._events = [None] * self.depth
        self._ws = [None] * self.depth
        self._out = [None] * self.depth
        self._inputs = [None] * self.depth  # host tensors of the slot's call in flight (kept alive for the async copies)
        self._next = 0

    def submit(self, value: Tensor, spatial_shapes: Tensor, level_start_index: Tensor, sampling_loc: Tensor,
               attn_weight: Tensor, im2col_step: int = 64, flags: Optional[int] = None) -> int:
        for t in (value, spatial_shapes, level_start_index, sampling_loc, attn_weight):
            _require(not t.is_cuda and t.is_contiguous(), "HostPipeline takes contiguous host tensors")
        slot = self._next
        self._next = (self._next + 1) % self.depth
        if self._events[slot] is not None:
            self._events[slot].synchronize()  # the slot's previous call (and its output copy) has finished
            self._inputs[slot] = None
        bs, keys, heads, chans = value.shape
        queries, levels, points = sampling_loc.shape[1], sampling_loc.shape[3], sampling_loc.shape[4]
        dt = _DTYPES[value.dtype]
        need = int(_lib.msda_b200_host_workspace_bytes(bs, keys, heads, chans, levels, queries, points, dt))
        with torch.cuda.device(self.device):
            if self._ws[slot] is None or self._ws[slot].numel() < need:
                self._ws[slot] = torch.empty(max(need, 256), dtype=torch.uint8, device=self.device)
            shape = (bs, queries, heads * chans)
            if self._out[slot] is None or tuple(self._out[slot].shape) != shape or self._out[slot].dtype != value.dtype:
                self._out[slot] = torch.empty(shape, dtype=value.dtype).pin_memory()
            stream = self._streams[slot]
            rc = _lib.msda_b200_forward_host(
                value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), sampling_loc.data_ptr(),
                attn_weight.data_ptr(), self._out[slot].data_ptr(), self._ws[slot].data_ptr(), self._ws[slot].numel(),
                bs, keys, heads, chans, levels, queries, points, int(im2col_step), dt,
                _default_flags if flags is None else int(flags), int(stream.cuda_stream),
            )
            _check(rc)
            ev = torch.cuda.Event()
            ev.record(stream)
            self._events[slot] = ev
            # the copies are asynchronous: hold the caller's host tensors until the slot's event has completed, so a
            # temporary (e.g. ``x.pin_memory()``) cannot be recycled by the host allocator while the DMA reads it.
            # The CONTENTS are still the caller's: do not write to an input before ``result(ticket)`` returns.
            self._inputs[slot] = (value, spatial_shapes, level_start_index, sampling_loc, attn_weight)
        return slot

    def result(self, ticket: int) -> Tensor:
        """The call's output: the slot's pinned host buffer, valid until ``depth`` further ``submit`` calls reuse
        the slot -- copy it if it has to live longer."""
        self._events[ticket].synchronize()
        self._inputs[ticket] = None
        return self._out[ticket]

    def drain(self) -> None:
        for ev in self._events:
            if ev is not None:
                ev.synchronize()


def read_bandwidth_probe(device: torch.device, working_set_bytes: int, repeats: int, trials: int = 3) -> float:
    """GB/s of 16-byte streaming reads over a working set (``msda_b200_read_probe``): a set well below
    the 126 MB L2 measures L2->SM bandwidth, one far above it HBM.  Used for the roofline denominators
    MEASURED_PEAKS.json does not carry."""
    with torch.cuda.device(device):
        buf = torch.empty(int(working_set_bytes), dtype=torch.uint8, device=device).random_(0, 255)
        sink = torch.zeros(4, dtype=torch.int32, device=device)
        stream = torch.cuda.current_stream(device).cuda_stream
        for _ in range(2):
            _check(_lib.msda_b200_read_probe(buf.data_ptr(), buf.numel(), repeats, sink.data_ptr(), stream))
        torch.cuda.synchronize(device)
        best = 0.0
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(trials):
            start.record()
            _check(_lib.msda_b200_read_probe(buf.data_ptr(), buf.numel(), repeats, sink.data_ptr(), stream))
            end.record()
            torch.cuda.synchronize(device)
            best = max(best, buf.numel() * repeats / (start.elapsed_time(end) * 1e-3) / 1e9)
    return best


# ---------------------------------------------------------------------------
# torch.library registration: same namespace, name and schema as the reference
# ---------------------------------------------------------------------------
def _fake(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    """Meta kernel: the rank / dtype / extent agreement the reference's fake kernel asserts (ops.py:59-84) and
    its output properties (:86-87), written as one table of conditions."""
    ranks = ((value, 4), (spatial_shapes, 2), (level_start_index, 1), (sampling_loc, 6), (attn_weight, 5))
    for t, r in ranks:
        torch._check(t.dim() == r)
    bs, _, heads, per_head = value.shape
    levels = spatial_shapes.shape[0]
    conditions = (
        value.dtype == attn_weight.dtype, value.dtype == sampling_loc.dtype,
        spatial_shapes.dtype == torch.int64, level_start_index.dtype == torch.int64,
        spatial_shapes.shape[1] == 2, level_start_index.shape[0] == levels,
        sampling_loc.shape[0] == bs, sampling_loc.shape[2] == heads, sampling_loc.shape[3] == levels,
        sampling_loc.shape[5] == 2,
    )
    for ok in conditions:
        torch._check(ok)
    for axis in range(5):
        torch._check(attn_weight.shape[axis] == sampling_loc.shape[axis])
    return value.new_empty((bs, sampling_loc.shape[1], heads * per_head))


def _fake_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output, grad_value,
                   grad_sampling_loc, grad_attn_weight, im2col_step):
    # the backward op only mutates its three gradient arguments
    torch._check(grad_value.shape == value.shape)
    torch._check(grad_sampling_loc.shape == sampling_loc.shape)
    torch._check(grad_attn_weight.shape == attn_weight.shape)
    return None


def _op_cuda(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    # the dispatcher's CUDA-key entry: one validation pass, one allocation, one C-ABI call
    _validate(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)
    bs, keys, heads, chans = value.shape
    queries, levels, points = sampling_loc.shape[1], sampling_loc.shape[3], sampling_loc.shape[4]
    dev = value.device
    out = torch.empty((bs, queries, heads * chans), dtype=value.dtype, device=dev)
    ws_ptr = ws_bytes = 0
    if _use_workspace:
        ws_bytes = workspace_bytes(value, sampling_loc)
        if ws_bytes:
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            ws_ptr = ws.data_ptr()
    guard = torch.cuda.device(dev) if torch.cuda.current_device() != dev.index else None
    if guard is not None:
        guard.__enter__()
    try:
        rc = _lib.msda_b200_forward_ws(
            value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), sampling_loc.data_ptr(),
            attn_weight.data_ptr(), out.data_ptr(), ws_ptr, ws_bytes, bs, keys, heads, chans, levels, queries, points,
            int(im2col_step), _DTYPES[value.dtype], _default_flags, torch.cuda.current_stream(dev).cuda_stream)
    finally:
        if guard is not None:
            guard.__exit__(None, None, None)
    if rc != 0:
        _check(rc)
    return out


def _op_backward_cuda(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output, grad_value,
                      grad_sampling_loc, grad_attn_weight, im2col_step):
    backward_into(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output, grad_value,
                  grad_sampling_loc, grad_attn_weight, im2col_step)


def _autograd_backward(ctx, grad):
    # same glue as the reference's codetr/ops.py:90-114: zero-filled gradient buffers, one backward op call
    value, spatial_shapes, level_start_index, sampling_loc, attn_weight = ctx.saved_tensors
    grad_value = torch.zeros_like(value)
    grad_loc = torch.empty_like(sampling_loc)
    grad_w = torch.empty_like(attn_weight)
    torch.ops.codetr.multi_scale_deformable_attention_backward(
        value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad.contiguous(), grad_value, grad_loc,
        grad_w, ctx.im2col_step)
    return grad_value, None, None, grad_loc, grad_w, None


def _autograd_setup(ctx, inputs, output):
    value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step = inputs
    ctx.im2col_step = im2col_step
    ctx.save_for_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight)


_torch_lib = None
op_registration = "none"   # "native" (codetr_b200_torch.so), "python", or "external" (someone else defined codetr::)


def register_torch_op() -> None:
    """Make ``torch.ops.codetr.multi_scale_deformable_attention`` (+ ``_backward``) available.

    Preference order: (1) a library that already defines the namespace (the reference's own
    ``codetr_cpp_extension.so`` or the drop-in build) is left alone; (2) the package's optional native
    registration ``csrc/_torch/codetr_b200_torch.so`` (7 us of host time per call) unless
    ``MSDA_B200_PYTHON_OP=1``; (3) registration from Python (26 us per call; needs nothing but the C-ABI
    library).  The fake kernels and the autograd glue are registered from Python in cases (2) and (3)."""
    global _torch_lib, op_registration
    if _torch_lib is not None:
        return
    import os

    already = hasattr(torch.ops, "codetr") and hasattr(torch.ops.codetr, "multi_scale_deformable_attention")
    if already:
        _torch_lib, op_registration = False, "external"
        return
    native = _native.TORCH_BINDING_PATH
    if os.path.isfile(native) and os.environ.get("MSDA_B200_PYTHON_OP", "0") != "1":
        try:
            torch.ops.load_library(native)
            lib = torch.library.Library("codetr", "FRAGMENT")
            op_registration = "native"
        except Exception:  # stale / incompatible build: fall through to the Python registration
            lib = None
    else:
        lib = None
    if lib is None:
        lib = torch.library.Library("codetr", "DEF")
        lib.define(OP_SCHEMA)
        lib.define(BACKWARD_SCHEMA)
        lib.impl("multi_scale_deformable_attention", _op_cuda, "CUDA")
        lib.impl("multi_scale_deformable_attention_backward", _op_backward_cuda, "CUDA")
        op_registration = "python"
    torch.library.register_fake("codetr::multi_scale_deformable_attention", _fake, lib=lib)
    torch.library.register_fake("codetr::multi_scale_deformable_attention_backward", _fake_backward, lib=lib)
    torch.library.register_autograd("codetr::multi_scale_deformable_attention", _autograd_backward,
                                    setup_context=_autograd_setup, lib=lib)
    _torch_lib = lib


register_torch_op()
