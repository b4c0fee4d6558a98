"""Timing of the value producer (manual, GPU box):  python tests/perf_value_proj.py [--out gpurun_out/value_proj.json]

``msda_b200_value_proj`` (one tcgen05 kernel: Linear + bias + masked_fill, output in the op's layout) against what
the reference module runs for the same lines (multi_scale_deformable_attention.py:173-176): ``nn.Linear`` (cuBLAS)
followed by ``masked_fill``.  CUDA events over back-to-back calls, inputs rotated over > L2 worth of buffers.
Algorithmic HBM bytes per call = rows * (K + N) * 2 + N * K * 2 + rows (x read, value written, weights, mask).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.nn.functional as F

import codetr_b200 as cb
from codetr_b200 import workloads as W

L2 = 126 * 1024 * 1024


def time_calls(fns, iters, warmup=20):
    for i in range(warmup):
        fns[i % len(fns)]()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = None
    for _ in range(3):
        s.record()
        for i in range(iters):
            fns[i % len(fns)]()
        e.record()
        torch.cuda.synchronize()
        us = 1e3 * s.elapsed_time(e) / iters
        best = us if best is None else min(best, us)
    return best


def time_graphed(fns, replays=20):
    """Device time per call with the host out of the picture: all calls of `fns` captured into one CUDA graph."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for f in fns:
            f()
        side.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for f in fns:
                f()
    for _ in range(3):
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
        us = 1e3 * s.elapsed_time(e) / (replays * len(fns))
        best = us if best is None else min(best, us)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "value_proj.json"))
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    rows_out = []
    for name, batch, dtn in (("swinl_enc_1152x768", 1, "float16"), ("swinl_enc_1152x768", 1, "bfloat16"),
                             ("swinl_enc_1152x768", 4, "float16"), ("r50_enc_608", 1, "float16"),
                             ("swinl_enc_1920x1280", 2, "float16"), ("swinl_enc_1152x768_s4", 1, "float16"),
                             ("swinl_dec_1152x768:queries", 1, "float16")):  # 900 rows: the decoder's output_proj
        wl = W.CONFIGS[name.split(":")[0]]
        dt = getattr(torch, dtn)
        K = N = wl.num_heads * wl.channels
        keys = wl.Q if name.endswith(":queries") else wl.S
        rows = batch * keys
        per_set = rows * (K + N) * 2
        n_sets = min(32, max(2, -(-int(1.5 * L2) // per_set)))
        torch.manual_seed(0)
        w = (torch.randn(N, K, device=dev) / K ** 0.5).to(dt)
        b = torch.randn(N, device=dev).to(dt)
        xs = [torch.randn(batch, keys, K, device=dev).to(dt) for _ in range(n_sets)]
        mask = torch.zeros(batch, keys, dtype=torch.bool, device=dev)
        mask[:, -keys // 10:] = True
        ours = [(lambda x=x: cb.value_proj(x, w, b, mask, num_heads=wl.num_heads)) for x in xs]
        lib = [(lambda x=x: F.linear(x, w, b).masked_fill(mask[..., None], 0.0).unflatten(-1, (wl.num_heads, -1))) for x in xs]
        gemm_only = [(lambda x=x: F.linear(x, w, b)) for x in xs]
        iters = 200 if rows < 40000 else 60
        with torch.inference_mode():
            got, want = ours[0](), lib[0]()
            err = float((got.float() - want.float()).abs().max() / want.float().abs().max())
            t_ours, t_lib, t_gemm = time_calls(ours, iters), time_calls(lib, iters), time_calls(gemm_only, iters)
            g_ours, g_lib, g_gemm = time_graphed(ours), time_graphed(lib), time_graphed(gemm_only)
            g_cluster = {}
            for c in (1, 2, 4):
                os.environ["MSDA_B200_VPROJ_CLUSTER"] = str(c)
                g_cluster[f"graphed_value_proj_cluster{c}_us"] = time_graphed(ours)
            os.environ.pop("MSDA_B200_VPROJ_CLUSTER")
            for ns in (1, 2, 4):
                os.environ["MSDA_B200_VPROJ_NSPLIT"] = str(ns)
                g_cluster[f"graphed_value_proj_nsplit{ns}_us"] = time_graphed(ours)
            os.environ.pop("MSDA_B200_VPROJ_NSPLIT")
            os.environ["MSDA_B200_VPROJ_SINGLE_TILE"] = "1"
            g_single = time_graphed(ours)
            os.environ.pop("MSDA_B200_VPROJ_SINGLE_TILE")
        # the consumer side: output_proj + residual (same kernel, residual epilogue) vs Linear followed by an add
        res = [torch.randn(batch, keys, N, device=dev).to(dt) for _ in range(n_sets)]
        ours_o = [(lambda x=x, r=r: cb.output_proj(x, w, b, r)) for x, r in zip(xs, res)]
        lib_o = [(lambda x=x, r=r: F.linear(x, w, b) + r) for x, r in zip(xs, res)]
        with torch.inference_mode():
            g_out, g_out_lib = time_graphed(ours_o), time_graphed(lib_o)
        hbm = rows * (K + N) * 2 + N * K * 2 + rows
        row = {"workload": name, "batch": batch, "dtype": dtn, "rows": rows, "K": K, "N": N, "n_sets": n_sets,
               "value_proj_us": t_ours, "linear_masked_fill_us": t_lib, "linear_only_us": t_gemm,
               "graphed_value_proj_us": g_ours, **g_cluster, "graphed_single_tile_variant_us": g_single, "graphed_linear_masked_fill_us": g_lib, "graphed_linear_only_us": g_gemm,
               "graphed_output_proj_us": g_out, "graphed_linear_add_us": g_out_lib,
               "output_proj_hbm_GBps": (rows * (K + 2 * N) * 2 + N * K * 2) / g_out / 1e3,
               "hbm_GBps": hbm / g_ours / 1e3, "tflops": 2.0 * rows * K * N / g_ours / 1e6,
               "max_rel_vs_cublas_path": err}
        rows_out.append(row)
        print("  ".join(f"{k}={v:.4g}" if isinstance(v, float) else f"{k}={v}" for k, v in row.items()), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump({"device": torch.cuda.get_device_name(dev), "measured_peaks": peaks, "rows": rows_out}, f, indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
