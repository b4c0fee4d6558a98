"""Calling-context timing (manual, GPU box):  python tests/perf_module.py [--out gpurun_out/module.json]

Times one forward of the ``MultiScaleDeformableAttention`` module (the caller of the op, SURVEY section 8(f).1:
value_proj + masked_fill, the two producer Linears, softmax + location arithmetic, the op, output_proj +
residual) at the BASELINE shapes, with CUDA events over back-to-back forwards:

  unfused      producers as separate PyTorch ops + this repo's op (what the unchanged reference module does)
  fused        ``fused_producers=True``: softmax + locations inside the kernel (``msda_b200_forward_fused``)
  fused_all    + ``fused_value_proj`` / ``fused_output_proj``: value_proj + masked_fill and output_proj + residual in the
               tcgen05 projection kernel (``msda_b200_value_proj`` / ``msda_b200_output_proj``)
  *_graphed    the same forwards captured in a CUDA graph: device time without the eager-mode host overhead
  reference    the same module with the reference's own CUDA kernel (oracle/_ref, rebuilt for sm_100a) as the op
  op only      this repo's op alone on the tensors the module feeds it

Lives under tests/ because it executes oracle/_ref; nothing here is on the product path.
"""
import argparse
import json
import os
import sys
from unittest import mock

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

import codetr_b200 as cb
from codetr_b200 import workloads as W
from oracle import build_ref


def time_fn(fn, iters, warmup=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = None
    for _ in range(3):
        s.record()
        for _ in range(iters):
            fn()
        e.record()
        torch.cuda.synchronize()
        us = 1e3 * s.elapsed_time(e) / iters
        best = us if best is None else min(best, us)
    return best


def time_graphed(fn, copies=4, replays=10):
    """Device time per forward with the host out of the picture (what an engine runtime sees)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
        side.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(copies):
                fn()
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = None
    for _ in range(3):
        s.record()
        for _ in range(replays):
            g.replay()
        e.record()
        torch.cuda.synchronize()
        us = 1e3 * s.elapsed_time(e) / (replays * copies)
        best = us if best is None else min(best, us)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "module.json"))
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    have_ref = bool(build_ref.load_if_built())
    rows = []
    for name, batch, dtn in (("swinl_enc_1152x768", 1, "float16"), ("swinl_enc_1152x768", 1, "bfloat16"),
                             ("swinl_enc_1152x768", 1, "float32"), ("r50_enc_608", 1, "float16"),
                             ("swinl_dec_1152x768", 1, "float16"), ("swinl_enc_1920x1280", 2, "float16")):
        wl = W.CONFIGS[name]
        dt = getattr(torch, dtn)
        inp = W.make_inputs(wl, batch=batch, seed=wl.seed)
        torch.manual_seed(0)
        embed = wl.num_heads * wl.channels
        mods = {}
        for fused in (False, True):
            m = cb.MultiScaleDeformableAttention(embed_dims=embed, num_heads=wl.num_heads, num_levels=wl.L,
                                                 num_points=wl.num_points, batch_first=True, dropout=0.0,
                                                 fused_producers=fused)
            if mods:
                m.load_state_dict(mods[False].state_dict())
            else:  # non-trivial producers: the reference initialises both to zero weight
                torch.nn.init.normal_(m.sampling_offsets.weight, std=0.02)
                torch.nn.init.normal_(m.attention_weights.weight, std=0.5)
            mods[fused] = m.to(device=dev, dtype=dt).eval()
        query = torch.randn(batch, wl.Q, embed, device=dev, dtype=dt)
        feats = query if wl.Q == wl.S else torch.randn(batch, wl.S, embed, device=dev, dtype=dt)
        ref_pts = torch.from_numpy(inp.reference_points).to(device=dev, dtype=dt)
        shapes = torch.from_numpy(inp.spatial_shapes).to(dev)
        starts = torch.from_numpy(inp.level_start_index).to(dev)
        mask = torch.zeros(batch, wl.S, dtype=torch.bool, device=dev)
        mask[:, -wl.S // 10:] = True
        kw = dict(value=feats, key_padding_mask=mask, reference_points=ref_pts, spatial_shapes=shapes, level_start_index=starts)
        iters = 100 if wl.Q * batch < 40000 else 40
        with torch.inference_mode():
            out_unfused = mods[False](query, **kw)
            out_fused = mods[True](query, **kw)
            scale = out_unfused.float().abs().max()
            row = {"workload": name, "batch": batch, "dtype": dtn,
                   "fused_vs_unfused_max_rel": float((out_fused.float() - out_unfused.float()).abs().max() / scale),
                   "unfused_us": time_fn(lambda: mods[False](query, **kw), iters),
                   "fused_us": time_fn(lambda: mods[True](query, **kw), iters)}
            if dt != torch.float32:  # + the tensor-core value producer (Linear + masked_fill in one kernel)
                mods[True].fused_value_proj = mods[True].fused_output_proj = True
                out_all = mods[True](query, **kw)
                row["fused_all_vs_unfused_max_rel"] = float((out_all.float() - out_unfused.float()).abs().max() / scale)
                row["fused_all_us"] = time_fn(lambda: mods[True](query, **kw), iters)
                row["fused_all_graphed_us"] = time_graphed(lambda: mods[True](query, **kw))
                mods[True].fused_value_proj = mods[True].fused_output_proj = False
            row["unfused_graphed_us"] = time_graphed(lambda: mods[False](query, **kw))
            row["fused_graphed_us"] = time_graphed(lambda: mods[True](query, **kw))
            # the op alone, on what the module feeds it
            m = mods[False]
            keys = m._keys(feats, mask).contiguous()
            offs = m.sampling_offsets(query).view(batch, wl.Q, wl.num_heads, wl.L, wl.num_points, 2)
            logits = m.attention_weights(query).view(batch, wl.Q, wl.num_heads, wl.L * wl.num_points)
            loc, wts = m._unfused_producers(offs, logits, ref_pts, shapes)
            wts = wts.contiguous()
            call = cb.PreparedForward(keys, shapes, starts, loc, wts)
            row["op_only_us"] = time_fn(call, iters)
            row["producers_only_us"] = time_fn(lambda: m._unfused_producers(offs, logits, ref_pts, shapes), iters)
            if have_ref and dtn in ("float16", "float32"):
                ref_op = lambda v, s, l, lo, w, step: torch.ops.codetr_ref.msda_forward(v, s, l, lo, w.contiguous(), step)
                with mock.patch.object(torch.ops.codetr, "multi_scale_deformable_attention", new=ref_op):
                    out_ref = mods[False](query, **kw)
                    row["reference_kernel_us"] = time_fn(lambda: mods[False](query, **kw), max(10, iters // 2))
                row["ours_vs_reference_kernel_max_rel"] = float((out_unfused.float() - out_ref.float()).abs().max() / scale)
        rows.append(row)
        print("  ".join(f"{k}={v:.3g}" if isinstance(v, float) else f"{k}={v}" for k, v in row.items()), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump({"device": torch.cuda.get_device_name(dev), "rows": rows}, f, indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
