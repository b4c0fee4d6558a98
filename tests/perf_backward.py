"""Backward timing (manual, GPU box): this repo's backward vs the reference's own CUDA backward rebuilt
for sm_100a (oracle/_ref), same tensors, CUDA events.  python tests/perf_backward.py [--out ...]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

import codetr_b200 as cb
from codetr_b200 import workloads as W
from oracle import build_ref

KEYS = ("value", "spatial_shapes", "level_start_index", "sampling_loc", "attn_weight")


def timeit(fn, iters=30, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return 1e3 * s.elapsed_time(e) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "backward.json"))
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    have_ref = build_ref.load_if_built()
    rows = []
    for name, dtn in (("swinl_enc_1152x768", "float16"), ("swinl_enc_1152x768", "float32"), ("swinl_dec_1152x768", "float16"),
                      ("r50_enc_608", "float16")):
        wl = W.CONFIGS[name]
        dt = getattr(torch, dtn)
        inp = W.make_inputs(wl, batch=1)
        d = {k: torch.from_numpy(getattr(inp, k)) for k in KEYS}
        d = {k: (v.to(dev) if v.dtype == torch.int64 else v.to(device=dev, dtype=dt)) for k, v in d.items()}
        go = torch.randn(1, wl.Q, 256, device=dev, dtype=dt)
        gv, gl, gw = torch.zeros_like(d["value"]), torch.empty_like(d["sampling_loc"]), torch.empty_like(d["attn_weight"])

        def ours():
            gv.zero_()
            cb.backward_into(*(d[k] for k in KEYS), go, gv, gl, gw)

        us = timeit(ours)
        row = {"workload": name, "dtype": dtn, "ours_us": us, "variant": cb.last_variant()}
        if have_ref:
            rgv, rgl, rgw = torch.zeros_like(gv), torch.zeros_like(gl), torch.zeros_like(gw)

            def ref():
                rgv.zero_(); rgl.zero_(); rgw.zero_()
                torch.ops.codetr_ref.msda_backward(*(d[k] for k in KEYS), go, rgv, rgl, rgw, 64)

            row["reference_cuda_us"] = timeit(ref)
            ours(); ref(); torch.cuda.synchronize()
            row["grad_loc_max_rel_vs_ref"] = float((gl.float() - rgl.float()).abs().max() / rgl.float().abs().max())
            row["grad_value_max_rel_vs_ref"] = float((gv.float() - rgv.float()).abs().max() / rgv.float().abs().max())
        rows.append(row)
        print(row, flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
