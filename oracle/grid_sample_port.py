"""PyTorch ``grid_sample`` restatement of the reference's CPU path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  This is the function
``bench.py`` times as the CPU baseline ("kind": "port"): it issues the same
sequence of ATen operations as the reference's
``multi_scale_deformable_attention_pytorch``
(/root/reference/codetr/ops.py:129-186) so the timing is representative of
what a user of the reference pays on host cores:

1. split the flattened pyramid per level                      (ops.py:155)
2. map locations from [0,1] to grid_sample's [-1,1]            (ops.py:156)
3. per level: view the level as ``[B*M, D, H, W]`` and sample it with
   ``F.grid_sample(bilinear, zeros, align_corners=False)``    (ops.py:158-175)
4. stack levels, multiply by the per-point weights, reduce    (ops.py:176-186)

The arithmetic that matters lives in ATen's ``grid_sampler_2d`` (third-party
relative to the reference: torch, reference pinned 2.6.0, this image 2.11.0);
parity is pinned by ``tests/golden`` (generated from the reference's function
itself) rather than by reading ATen.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def msda_forward_grid_sample(
    value: torch.Tensor,
    spatial_shapes,
    sampling_loc: torch.Tensor,
    attn_weight: torch.Tensor,
) -> torch.Tensor:
    """``value [B,S,M,D]``, ``spatial_shapes [L,2]`` (H,W), ``sampling_loc
    [B,Q,M,L,P,2]`` (x,y in [0,1]), ``attn_weight [B,Q,M,L,P]`` ->
    ``[B,Q,M*D]``."""
    n_img, _, n_head, n_chan = value.shape
    n_query, n_level, n_point = sampling_loc.shape[1], sampling_loc.shape[3], sampling_loc.shape[4]
    hw = [(int(h), int(w)) for h, w in torch.as_tensor(spatial_shapes).tolist()]

    per_level = torch.split(value, [h * w for h, w in hw], dim=1)
    grids = sampling_loc * 2 - 1
    sampled = []
    for lvl, (h, w) in enumerate(hw):
        # [B, H*W, M, D] -> [B, H*W, M*D] -> [B, M*D, H*W] -> [B*M, D, H, W]
        fmap = per_level[lvl].flatten(2).transpose(1, 2).reshape(n_img * n_head, n_chan, h, w)
        # [B, Q, M, P, 2] -> [B, M, Q, P, 2] -> [B*M, Q, P, 2]
        grid = grids[:, :, :, lvl].transpose(1, 2).flatten(0, 1)
        sampled.append(F.grid_sample(fmap, grid, mode="bilinear", padding_mode="zeros", align_corners=False))
    # [B, Q, M, L, P] -> [B*M, 1, Q, L*P]
    wts = attn_weight.transpose(1, 2).reshape(n_img * n_head, 1, n_query, n_level * n_point)
    # [B*M, D, Q, L, P] -> [B*M, D, Q, L*P] -> weighted sum over the samples
    acc = (torch.stack(sampled, dim=-2).flatten(-2) * wts).sum(-1)
    return acc.view(n_img, n_head * n_chan, n_query).transpose(1, 2).contiguous()
