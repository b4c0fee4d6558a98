"""Caller of the op: a plain-``torch.nn`` mirror of the reference's ``MultiScaleDeformableAttention``
module (/root/reference/codetr/multi_scale_deformable_attention.py:15-218, an mmengine ``BaseModule``).

The reference module itself stays usable unchanged on top of ``torch.ops.codetr.multi_scale_deformable_attention``
(same op name and schema).  This mirror exists for two reasons: (1) it has no mmengine / mmcv dependency, so
the op can be exercised in its calling context here (same constructor arguments, same parameter names --
``sampling_offsets``, ``attention_weights``, ``value_proj``, ``output_proj`` -- so a reference ``state_dict``
loads, same forward signature and shape conventions); (2) ``fused_producers=True`` routes the softmax and the
sampling-location arithmetic (:180-200) into the kernel through ``msda_b200_forward_fused`` instead of running
them as separate PyTorch ops.

Differences, both deliberate: CPU tensors raise (the reference falls back to
``multi_scale_deformable_attention_pytorch`` at :207-210; this package has no CPU path), and ``norm_cfg`` /
``init_cfg`` (mmengine plumbing, unused by the forward) are accepted and ignored.
"""
from __future__ import annotations

import math
import warnings
from typing import Optional

import torch
from torch import nn

from . import ops


class MultiScaleDeformableAttention(nn.Module):
    def __init__(
        self,
        embed_dims: int = 256,
        num_heads: int = 8,
        num_levels: int = 4,
        num_points: int = 4,
        im2col_step: int = 64,
        dropout: float = 0.1,
        batch_first: bool = False,
        norm_cfg: Optional[dict] = None,
        init_cfg: Optional[dict] = None,
        value_proj_ratio: float = 1.0,
        fused_producers: bool = False,
    ):
        super().__init__()
        if embed_dims % num_heads != 0:  # reference :56-57
            raise ValueError(f"embed_dims must be divisible by num_heads, but got {embed_dims} and {num_heads}")
        dim_per_head = embed_dims // num_heads
        if dim_per_head & (dim_per_head - 1) or dim_per_head == 0:  # reference :69-75
            warnings.warn("the dimension of each attention head should be a power of 2 for the vector kernels")
        self.norm_cfg = norm_cfg
        self.dropout = nn.Dropout(dropout)
        self.batch_first = batch_first
        self.im2col_step = im2col_step
        self.embed_dims = embed_dims
        self.num_levels = num_levels
        self.num_heads = num_heads
        self.num_points = num_points
        self.fused_producers = fused_producers
        self.sampling_offsets = nn.Linear(embed_dims, num_heads * num_levels * num_points * 2)
        self.attention_weights = nn.Linear(embed_dims, num_heads * num_levels * num_points)
        value_proj_size = int(embed_dims * value_proj_ratio)
        self.value_proj = nn.Linear(embed_dims, value_proj_size)
        self.output_proj = nn.Linear(value_proj_size, embed_dims)
        self.init_weights()

    def init_weights(self) -> None:
        """Offsets start on a ring of directions, one per head, point p at distance p+1; attention logits
        start at zero; projections Xavier-uniform (reference :86-115)."""
        nn.init.zeros_(self.sampling_offsets.weight)
        thetas = torch.arange(self.num_heads, dtype=torch.float32) * (2.0 * math.pi / self.num_heads)
        grid = torch.stack([thetas.cos(), thetas.sin()], -1)
        grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(self.num_heads, 1, 1, 2)
        grid = grid.repeat(1, self.num_levels, self.num_points, 1)
        for i in range(self.num_points):
            grid[:, :, i, :] *= i + 1
        with torch.no_grad():
            self.sampling_offsets.bias.copy_(grid.view(-1))
        nn.init.zeros_(self.attention_weights.weight)
        nn.init.zeros_(self.attention_weights.bias)
        for lin in (self.value_proj, self.output_proj):
            nn.init.xavier_uniform_(lin.weight)
            nn.init.zeros_(lin.bias)

    def forward(
        self,
        query: torch.Tensor,
        key: Optional[torch.Tensor] = None,
        value: Optional[torch.Tensor] = None,
        identity: Optional[torch.Tensor] = None,
        query_pos: Optional[torch.Tensor] = None,
        key_padding_mask: Optional[torch.Tensor] = None,
        reference_points: Optional[torch.Tensor] = None,
        spatial_shapes: Optional[torch.Tensor] = None,
        level_start_index: Optional[torch.Tensor] = None,
        **kwargs,
    ) -> torch.Tensor:
        """Same contract as the reference forward (:117-218): ``query`` ``(num_query, bs, embed_dims)`` unless
        ``batch_first``; ``reference_points`` ``(bs, num_query, num_levels, 2|4)``; returns
        ``dropout(output_proj(msda)) + identity`` in the layout of ``query``."""
        if value is None:
            value = query
        if identity is None:
            identity = query
        if query_pos is not None:
            query = query + query_pos
        if not self.batch_first:
            query = query.permute(1, 0, 2)
            value = value.permute(1, 0, 2)
        bs, num_query, _ = query.shape
        _, num_value, _ = value.shape

        value = self.value_proj(value)
        if key_padding_mask is not None:
            value = value.masked_fill(key_padding_mask[..., None], 0.0)
        value = value.view(bs, num_value, self.num_heads, -1)
        offsets = self.sampling_offsets(query).view(bs, num_query, self.num_heads, self.num_levels, self.num_points, 2)
        logits = self.attention_weights(query).view(bs, num_query, self.num_heads, self.num_levels * self.num_points)
        ref_dim = reference_points.shape[-1]
        if ref_dim not in (2, 4):
            raise ValueError(f"Last dim of reference_points must be 2 or 4, but get {ref_dim} instead.")
        if not value.is_cuda:
            raise RuntimeError("codetr_b200 has no CPU path: MultiScaleDeformableAttention needs CUDA tensors")

        if self.fused_producers and not torch.is_grad_enabled():
            # softmax over L*P and the location arithmetic run inside the kernel
            output = ops.forward_fused(value.contiguous(), spatial_shapes, level_start_index,
                                       reference_points.to(value.dtype).contiguous(), offsets.contiguous(),
                                       logits.contiguous())
        else:
            weights = logits.softmax(-1).view(bs, num_query, self.num_heads, self.num_levels, self.num_points)
            if ref_dim == 2:
                normalizer = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1)
                locations = reference_points[:, :, None, :, None, :] + offsets / normalizer[None, None, None, :, None, :]
            else:
                locations = (reference_points[:, :, None, :, None, :2]
                             + offsets / self.num_points * reference_points[:, :, None, :, None, 2:] * 0.5)
            output = torch.ops.codetr.multi_scale_deformable_attention(
                value, spatial_shapes, level_start_index, locations.contiguous(), weights, self.im2col_step)

        output = self.output_proj(output)
        if not self.batch_first:
            output = output.permute(1, 0, 2)
        return self.dropout(output) + identity
