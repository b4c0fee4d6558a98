"""Caller of the op: a plain-``torch.nn`` mirror of the reference's ``MultiScaleDeformableAttention``
module (/root/reference/codetr/multi_scale_deformable_attention.py:15-218, an mmengine ``BaseModule``).

The reference module itself stays usable unchanged on top of ``torch.ops.codetr.multi_scale_deformable_attention``
(same op name and schema).  This mirror exists for two reasons: (1) it has no mmengine / mmcv dependency, so
the op can be exercised in its calling context here (same constructor arguments, same parameter names --
``sampling_offsets``, ``attention_weights``, ``value_proj``, ``output_proj`` -- so a reference ``state_dict``
loads, same forward signature and shape conventions); (2) ``fused_producers=True`` routes the softmax and the
sampling-location arithmetic (:180-200) into the kernel through ``msda_b200_forward_fused`` instead of running
them as separate PyTorch ops, and ``fused_value_proj=True`` replaces ``value_proj`` +
``masked_fill`` (:173-176) by the tensor-core kernel ``msda_b200_value_proj`` (section 8(f).4); ``fused_output_proj=True``
does the same for ``output_proj`` + residual (:212-218, inference mode) with ``msda_b200_output_proj``.

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
    """Constructor arguments, parameter names and the forward contract are the reference's (see the module
    docstring); ``fused_producers``, ``fused_value_proj`` and ``fused_output_proj`` are the only additions."""

    def __init__(self, embed_dims: int = 256, num_heads: int = 8, num_levels: int = 4, num_points: int = 4,
                 im2col_step: int = 64, dropout: float = 0.1, batch_first: bool = False, norm_cfg: Optional[dict] = None,
                 init_cfg: Optional[dict] = None, value_proj_ratio: float = 1.0, fused_producers: bool = False,
                 fused_value_proj: bool = False, fused_output_proj: bool = False):
        super().__init__()
        per_head, rem = divmod(embed_dims, num_heads)
        if rem:  # reference :56-57
            raise ValueError(f"embed_dims must be divisible by num_heads, but got {embed_dims} and {num_heads}")
        if per_head <= 0 or per_head & (per_head - 1):  # reference :69-75: a warning, not an error
            warnings.warn(f"{per_head} channels per head is not a power of two: the op falls back to its generic kernel")
        for name, val in dict(embed_dims=embed_dims, num_heads=num_heads, num_levels=num_levels, num_points=num_points,
                              im2col_step=im2col_step, batch_first=batch_first, norm_cfg=norm_cfg,
                              fused_producers=fused_producers, fused_value_proj=fused_value_proj,
                              fused_output_proj=fused_output_proj).items():
            setattr(self, name, val)
        samples = num_heads * num_levels * num_points
        inner = int(embed_dims * value_proj_ratio)
        # the reference's four Linear layers, under the reference's names (so its state_dict loads)
        self.sampling_offsets = nn.Linear(embed_dims, 2 * samples)
        self.attention_weights = nn.Linear(embed_dims, samples)
        self.value_proj = nn.Linear(embed_dims, inner)
        self.output_proj = nn.Linear(inner, embed_dims)
        self.dropout = nn.Dropout(dropout)
        self.init_weights()

    @torch.no_grad()
    def init_weights(self) -> None:
        """Reference :86-115.  Offset bias: head h points along direction 2*pi*h/num_heads (normalised to unit
        L-inf length), point p sits p+1 steps out, identical for every level; offset weights and attention
        logits start at zero; the two projections are Xavier-uniform with zero bias."""
        H, L, P = self.num_heads, self.num_levels, self.num_points
        angle = torch.arange(H, dtype=torch.float32) * (2.0 * math.pi / H)
        direction = torch.stack((angle.cos(), angle.sin()), dim=-1)
        direction = direction / direction.abs().amax(dim=-1, keepdim=True)
        steps = torch.arange(1, P + 1, dtype=torch.float32).view(1, 1, P, 1)
        bias = direction.view(H, 1, 1, 2) * steps                      # [H, 1, P, 2]
        self.sampling_offsets.bias.copy_(bias.expand(H, L, P, 2).reshape(-1))
        self.sampling_offsets.weight.zero_()
        self.attention_weights.weight.zero_()
        self.attention_weights.bias.zero_()
        for proj in (self.value_proj, self.output_proj):
            nn.init.xavier_uniform_(proj.weight)
            proj.bias.zero_()

    # -- pieces of the forward -----------------------------------------------------------------------
    def _keys(self, value: torch.Tensor, key_padding_mask: Optional[torch.Tensor]) -> torch.Tensor:
        """value_proj, zero the padded keys, split heads (reference :173-176) -> [bs, S, M, D]."""
        lin = self.value_proj
        if (self.fused_value_proj and not torch.is_grad_enabled()
                and ops.value_proj_supported(lin.in_features, lin.out_features, value.dtype)):
            # Linear + masked_fill + head split in one tcgen05 kernel
            mask = None if key_padding_mask is None else key_padding_mask.to(torch.bool).contiguous()
            return ops.value_proj(value.contiguous(), lin.weight, lin.bias, mask, num_heads=self.num_heads)
        v = lin(value)
        if key_padding_mask is not None:
            v = v.masked_fill(key_padding_mask.unsqueeze(-1), 0.0)
        return v.unflatten(-1, (self.num_heads, -1))

    def _unfused_producers(self, offsets, logits, reference_points, spatial_shapes):
        """softmax over L*P and the two location formulas (reference :180-200), as separate PyTorch ops."""
        bs, nq = offsets.shape[:2]
        weights = logits.softmax(-1).view(bs, nq, self.num_heads, self.num_levels, self.num_points)
        ref = reference_points[:, :, None, :, None, :]
        if reference_points.shape[-1] == 2:
            wh = spatial_shapes.flip(-1)  # (H, W) -> (W, H)
            locations = ref + offsets / wh[None, None, None, :, None, :]
        else:
            locations = ref[..., :2] + offsets / self.num_points * ref[..., 2:] * 0.5
        return locations.contiguous(), weights

    def forward(self, query: torch.Tensor, key: Optional[torch.Tensor] = None, value: Optional[torch.Tensor] = None,
                identity: Optional[torch.Tensor] = None, query_pos: Optional[torch.Tensor] = None,
                key_padding_mask: Optional[torch.Tensor] = None, reference_points: Optional[torch.Tensor] = None,
                spatial_shapes: Optional[torch.Tensor] = None, level_start_index: Optional[torch.Tensor] = None,
                **kwargs) -> torch.Tensor:
        """Reference :117-218.  ``query`` is ``(num_query, bs, embed_dims)`` unless ``batch_first``;
        ``reference_points`` ``(bs, num_query, num_levels, 2|4)``; returns
        ``dropout(output_proj(msda(...))) + identity`` in the layout of ``query``."""
        residual = query if identity is None else identity
        keys_in = query if value is None else value
        q = query if query_pos is None else query + query_pos
        if not self.batch_first:  # to (bs, n, embed_dims)
            q, keys_in = q.transpose(0, 1), keys_in.transpose(0, 1)
        if reference_points.shape[-1] not in (2, 4):
            raise ValueError(f"Last dim of reference_points must be 2 or 4, but get {reference_points.shape[-1]} instead.")
        if not q.is_cuda:
            raise RuntimeError("codetr_b200 has no CPU path: MultiScaleDeformableAttention needs CUDA tensors")

        bs, nq = q.shape[:2]
        keys = self._keys(keys_in, key_padding_mask)
        offsets = self.sampling_offsets(q).view(bs, nq, self.num_heads, self.num_levels, self.num_points, 2)
        logits = self.attention_weights(q).view(bs, nq, self.num_heads, self.num_levels * self.num_points)

        if self.fused_producers and not torch.is_grad_enabled():
            attended = ops.forward_fused(keys.contiguous(), spatial_shapes, level_start_index,
                                         reference_points.to(keys.dtype).contiguous(), offsets.contiguous(), logits.contiguous())
        else:
            locations, weights = self._unfused_producers(offsets, logits, reference_points, spatial_shapes)
            attended = torch.ops.codetr.multi_scale_deformable_attention(
                keys.contiguous(), spatial_shapes, level_start_index, locations, weights, self.im2col_step)

        lin = self.output_proj
        if (self.fused_output_proj and not torch.is_grad_enabled() and (not self.training or self.dropout.p == 0.0)
                and ops.value_proj_supported(lin.in_features, lin.out_features, attended.dtype)):
            # output_proj + residual in one tcgen05 kernel; the residual must have the (bs, n, E) memory order
            res = residual if self.batch_first else residual.transpose(0, 1)
            if res.is_contiguous() and res.dtype == attended.dtype:
                out = ops.output_proj(attended, lin.weight, lin.bias, res)
                return out if self.batch_first else out.transpose(0, 1)
        out = lin(attended)
        if not self.batch_first:
            out = out.transpose(0, 1)
        return self.dropout(out) + residual
