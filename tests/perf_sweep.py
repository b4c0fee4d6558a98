"""Tuning / comparison harness (manual, GPU box):  python tests/perf_sweep.py [--out gpurun_out/sweep.json]

Times, with CUDA events over back-to-back launches on rotating (> L2) input sets:
  * this repo's kernel under several launch configurations (env knobs of msda_sm100.cu),
  * the reference's own CUDA kernel rebuilt for sm_100a (oracle/_ref, when present) on the same tensors,
  * the L2 / HBM read probes that give the roofline denominators MEASURED_PEAKS.json lacks.
Lives under tests/ because it executes oracle/_ref; nothing here is on the product path.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import codetr_b200 as cb
from codetr_b200 import workloads as W
from oracle import build_ref

KEYS = ("value", "spatial_shapes", "level_start_index", "sampling_loc", "attn_weight")
L2 = 126 * 1024 * 1024


def device_sets(wl, batch, dt, dev, loc_mode=None, n_min=2):
    esize = torch.empty((), dtype=dt).element_size()
    hbm = W.algorithmic_hbm_bytes(wl, batch, esize)
    n = min(64, max(n_min, -(-int(1.5 * L2) // hbm)))
    host = []
    for i in range(2):
        inp = W.make_inputs(wl, batch=batch, seed=wl.seed + i, loc_mode=loc_mode)
        host.append({k: torch.from_numpy(getattr(inp, k)) for k in KEYS})
    sets = []
    for i in range(n):
        h = host[i % 2]
        sets.append({k: (h[k].to(dev) if h[k].dtype == torch.int64 else h[k].to(device=dev, dtype=dt)) for k in KEYS})
    return sets, hbm


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


def set_env(cfg):
    for k in ("MSDA_B200_TILE_W", "MSDA_B200_TILE_H", "MSDA_B200_HEAD_MAJOR", "MSDA_B200_SPLIT", "MSDA_B200_CTAS_PER_SM",
              "MSDA_B200_SMALL", "MSDA_B200_SPLIT_MAX_CTAS", "MSDA_B200_CHUNKED", "MSDA_B200_DYN", "MSDA_B200_L2_PREFETCH"):
        os.environ.pop(k, None)
    for k, v in cfg.items():
        if k.startswith("MSDA_"):
            os.environ[k] = str(v)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.json"))
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default=None, help="'headline': headline + decoder rows with a short config list")
    ap.add_argument("--no-probes", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    results = {"device": torch.cuda.get_device_name(dev), "rows": [], "probes": []}
    have_ref = build_ref.load_if_built()
    results["reference_cuda_loaded"] = bool(have_ref)

    # ---- read probes: L2 (working set << L2) and HBM (working set >> L2) ----
    sink = torch.zeros(4, dtype=torch.int32, device=dev)
    for mb, reps in (() if args.no_probes else ((8, 64), (32, 32), (64, 16), (96, 12), (512, 2), (2048, 1))):
        buf = torch.empty(mb * 1024 * 1024, dtype=torch.uint8, device=dev).random_(0, 255)
        lib = cb._native.load()
        stream = torch.cuda.current_stream().cuda_stream
        for _ in range(2):
            lib.msda_b200_read_probe(buf.data_ptr(), buf.numel(), reps, sink.data_ptr(), stream)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = None
        for _ in range(3):
            s.record()
            lib.msda_b200_read_probe(buf.data_ptr(), buf.numel(), reps, sink.data_ptr(), stream)
            e.record()
            torch.cuda.synchronize()
            gbs = buf.numel() * reps / (s.elapsed_time(e) * 1e-3) / 1e9
            best = gbs if best is None else max(best, gbs)
        results["probes"].append({"working_set_MB": mb, "repeats": reps, "read_GBps": best})
        print(f"probe {mb:5d} MB x{reps:3d}: {best:8.1f} GB/s", flush=True)
        del buf

    workloads = [("swinl_enc_1152x768", 1, "float16", None), ("swinl_enc_1152x768", 1, "float16", "uniform"),
                 ("swinl_enc_1152x768", 1, "float16", "adversarial"),
                 ("swinl_enc_1152x768", 4, "float16", None), ("swinl_enc_1152x768", 1, "bfloat16", None),
                 ("swinl_enc_1152x768", 1, "float32", None),
                 ("r50_enc_608", 1, "float16", None), ("swinl_dec_1152x768", 1, "float16", None),
                 ("swinl_dec_1152x768", 8, "float16", None), ("swinl_dec_1900q", 1, "float16", None),
                 ("swinl_enc_1920x1280", 2, "float16", None), ("swinl_enc_1152x768_s4", 1, "float16", None),
                 ("ref_test_mid_fp32", 1, "float32", None)]
    if args.quick:
        workloads = workloads[:2]
    if args.only == "chunk":
        workloads = [("swinl_enc_1152x768", 1, "float16", None), ("swinl_enc_1152x768", 4, "float16", None),
                     ("swinl_enc_1152x768", 1, "float32", None), ("swinl_enc_1920x1280", 2, "float16", None),
                     ("r50_enc_608", 1, "float16", None)]
    if args.only == "ctas":
        workloads = [("swinl_enc_1152x768", 1, "float16", None), ("swinl_enc_1152x768", 4, "float16", None),
                     ("swinl_enc_1152x768", 1, "float32", None), ("swinl_enc_1920x1280", 2, "float16", None),
                     ("r50_enc_608", 1, "float16", None), ("swinl_dec_1152x768", 8, "float16", None)]
    if args.only == "dyn":
        workloads = [("swinl_enc_1152x768", 1, "float16", None), ("swinl_enc_1152x768", 1, "float16", "adversarial"),
                     ("swinl_enc_1152x768", 4, "float16", None), ("swinl_enc_1152x768", 1, "bfloat16", None),
                     ("swinl_enc_1152x768", 1, "float32", None), ("swinl_enc_1920x1280", 2, "float16", None),
                     ("r50_enc_608", 1, "float16", None), ("r50_enc_608", 2, "float16", None),
                     ("swinl_dec_1152x768", 8, "float16", None), ("swinl_dec_1900q", 1, "float16", None),
                     ("swinl_enc_1152x768_s4", 1, "float16", None)]
    if args.only == "prefetch":
        workloads = [("swinl_enc_1152x768", 1, "float16", None), ("swinl_enc_1152x768", 1, "float16", "uniform"),
                     ("swinl_enc_1152x768", 4, "float16", None), ("swinl_enc_1152x768", 1, "bfloat16", None),
                     ("swinl_enc_1152x768", 1, "float32", None), ("r50_enc_608", 1, "float16", None),
                     ("swinl_dec_1152x768", 8, "float16", None), ("swinl_dec_1900q", 1, "float16", None),
                     ("swinl_enc_1920x1280", 2, "float16", None), ("ref_test_mid_fp32", 1, "float32", None)]
    if args.only == "exact":
        workloads = [("swinl_enc_1152x768", 1, "bfloat16", None), ("swinl_enc_1152x768", 4, "bfloat16", None),
                     ("swinl_enc_1152x768", 1, "float32", None), ("swinl_enc_1152x768", 1, "float16", None),
                     ("r50_enc_608", 1, "bfloat16", None), ("swinl_dec_1152x768", 1, "bfloat16", None),
                     ("swinl_dec_1152x768", 8, "bfloat16", None), ("swinl_enc_1920x1280", 2, "bfloat16", None),
                     ("ref_test_mid_fp32", 1, "float32", None)]
    if args.only == "decoder":
        workloads = [("swinl_dec_1152x768", 1, "float16", None), ("swinl_dec_1900q", 1, "float16", None),
                     ("swinl_dec_1152x768", 8, "float16", None), ("ref_test_mid_fp32", 1, "float32", None),
                     ("r50_enc_608", 1, "float16", None)]
    if args.only == "headline":
        workloads = [("swinl_enc_1152x768", 1, "float16", None), ("swinl_enc_1152x768", 1, "float32", None),
                     ("swinl_dec_1152x768", 1, "float16", None), ("swinl_enc_1920x1280", 2, "float16", None)]
    base_cfgs = [
        {"name": "default", "flags": 0},
        {"name": "fhfma", "flags": cb.FLAG_MATH_FHFMA},
        {"name": "linear", "flags": cb.FLAG_LINEAR_ORDER},
        {"name": "linear+fhfma", "flags": cb.FLAG_LINEAR_ORDER | cb.FLAG_MATH_FHFMA},
        {"name": "query-major", "flags": 0, "MSDA_B200_HEAD_MAJOR": 0},
        {"name": "query-major+linear", "flags": cb.FLAG_LINEAR_ORDER, "MSDA_B200_HEAD_MAJOR": 0},
        {"name": "generic", "flags": cb.FLAG_FORCE_GENERIC},
        {"name": "packed", "flags": 0, "ws": True},
        {"name": "packed+exact", "flags": cb.FLAG_MATH_EXACT, "ws": True},
        {"name": "nopacked+exact", "flags": cb.FLAG_MATH_EXACT},
        {"name": "staged", "flags": cb.FLAG_STAGE_TMA},
    ]
    tile_cfgs = [{"name": f"tile{w}x{h}+fhfma", "flags": cb.FLAG_MATH_FHFMA, "MSDA_B200_TILE_W": w, "MSDA_B200_TILE_H": h}
                 for (w, h) in ((8, 1), (8, 2), (8, 8), (16, 2), (16, 4), (4, 4), (32, 2), (32, 8))]
    ctas_cfgs = [{"name": f"ctas_per_sm{c}", "flags": 0, "MSDA_B200_CTAS_PER_SM": c} for c in (4, 8, 12, 16)]
    split_cfgs = [{"name": f"split{s}", "flags": 0, "MSDA_B200_SPLIT": s} for s in (1, 2, 4)]
    split_cfgs += [{"name": f"split{s}+fhfma", "flags": cb.FLAG_MATH_FHFMA, "MSDA_B200_SPLIT": s} for s in (1, 4)]

    for name, batch, dtn, loc_mode in workloads:
        wl = W.CONFIGS[name]
        dt = getattr(torch, dtn)
        sets, hbm = device_sets(wl, batch, dt, dev, loc_mode)
        esize = torch.empty((), dtype=dt).element_size()
        gather = W.algorithmic_gather_bytes(wl, batch, esize)
        iters = 200 if wl.Q * batch < 40000 else 60
        cfgs = list(base_cfgs)
        if name == "swinl_enc_1152x768" and loc_mode is None and dtn == "float16":
            cfgs += tile_cfgs
        if wl.kind == "decoder" or name == "ref_test_mid_fp32":
            cfgs += split_cfgs
        if dtn == "float32":
            cfgs = [c for c in cfgs if "fhfma" not in c["name"]]
        if args.only == "chunk":
            cfgs = [{"name": "default", "flags": 0}, {"name": "strided", "flags": 0, "MSDA_B200_CHUNKED": 0}]
            cfgs += [{"name": f"chunked tile{w}x{h}", "flags": 0, "MSDA_B200_TILE_W": w, "MSDA_B200_TILE_H": h}
                     for (w, h) in ((8, 1), (8, 2), (8, 8), (8, 16), (16, 4), (16, 8), (32, 8))]
            cfgs += [{"name": "chunked head-major 8x8", "flags": cb.FLAG_HEAD_MAJOR, "MSDA_B200_TILE_W": 8, "MSDA_B200_TILE_H": 8}]
            have_ref = False
        if args.only == "ctas":
            cfgs = [{"name": "default", "flags": 0}] + ctas_cfgs
            have_ref = False
        if args.only == "decoder":
            cfgs = [{"name": "default", "flags": 0}, {"name": "no-small-kernel", "flags": 0, "MSDA_B200_SMALL": 0},
                    {"name": "small<=300ctas", "flags": 0, "MSDA_B200_SPLIT_MAX_CTAS": 300},
                    {"name": "small<=600ctas", "flags": 0, "MSDA_B200_SPLIT_MAX_CTAS": 600},
                    {"name": "small<=1200ctas", "flags": 0, "MSDA_B200_SPLIT_MAX_CTAS": 1200},
                    {"name": "split1", "flags": 0, "MSDA_B200_SPLIT": 1}]
            have_ref = False
        if args.only == "dyn":
            cfgs = [{"name": "default", "flags": 0}, {"name": "dyn", "flags": 0, "MSDA_B200_DYN": 1},
                    {"name": "dyn ctas_per_sm4", "flags": 0, "MSDA_B200_DYN": 1, "MSDA_B200_CTAS_PER_SM": 4},
                    {"name": "dyn ctas_per_sm8", "flags": 0, "MSDA_B200_DYN": 1, "MSDA_B200_CTAS_PER_SM": 8},
                    {"name": "dyn split1", "flags": 0, "MSDA_B200_DYN": 1, "MSDA_B200_SPLIT": 1},
                    {"name": "default again", "flags": 0}]
            have_ref = False
        if args.only == "prefetch":
            cfgs = [{"name": "default", "flags": 0}, {"name": "no L2 prefetch", "flags": 0, "MSDA_B200_L2_PREFETCH": 0},
                    {"name": "default again", "flags": 0}]
            have_ref = False
        if args.only == "exact":
            cfgs = [{"name": "default", "flags": 0}, {"name": "exact", "flags": cb.FLAG_MATH_EXACT}]
            have_ref = False
        if args.only == "headline":
            cfgs = [c for c in cfgs if c["name"] in ("default", "exact-placeholder", "tile8x8+fhfma", "tile16x4+fhfma",
                                                      "tile4x4+fhfma", "split1", "split4", "split1+fhfma", "packed", "packed+exact", "nopacked+exact", "staged")]
            have_ref = False
        ref32 = None
        if dtn != "float32":
            s0 = sets[0]
            ref32 = cb.multi_scale_deformable_attention(s0["value"].float(), s0["spatial_shapes"], s0["level_start_index"],
                                                        s0["sampling_loc"].float(), s0["attn_weight"].float(), flags=0)
        for cfg in cfgs:
            set_env(cfg)
            ws = None
            if cfg.get("ws"):
                need = cb.workspace_bytes(sets[0]["value"], sets[0]["sampling_loc"])
                if need == 0:
                    continue
                ws = torch.empty(need, dtype=torch.uint8, device=dev)
            calls = [cb.PreparedForward(*(s[k] for k in KEYS), flags=cfg["flags"], workspace=ws) for s in sets]
            us = time_calls(calls, iters)
            row = {"workload": name, "batch": batch, "dtype": dtn, "loc_mode": loc_mode or wl.kind, "config": cfg["name"],
                   "variant": cb.last_variant(), "us_per_call": us, "hbm_GBps": hbm / us / 1e3, "gather_GBps": gather / us / 1e3,
                   "images_per_s": batch / (us * 1e-6), "n_sets": len(sets)}
            if ref32 is not None:
                got = calls[0]().float()
                torch.cuda.synchronize()
                row["max_rel_vs_fp32"] = float((got - ref32).abs().max() / ref32.abs().max())
            results["rows"].append(row)
            print(f"{name:24s} b{batch} {dtn:8s} {row['loc_mode']:8s} {cfg['name']:20s} {us:9.2f} us  hbm {row['hbm_GBps']:7.1f} GB/s  "
                  f"gather {row['gather_GBps']:8.1f} GB/s  err {row.get('max_rel_vs_fp32', 0):.2e}  {row['variant']}", flush=True)
        set_env({})
        # opt-in fused-producer entry (softmax + location arithmetic in-kernel) on the same shapes
        if wl.kind in ("encoder", "decoder") and loc_mode is None:
            inp = W.make_inputs(wl, batch=batch, seed=wl.seed)
            cast = lambda a: torch.from_numpy(a).to(device=dev, dtype=dt)
            refp, off, lg = cast(inp.reference_points), cast(inp.sampling_offsets), cast(inp.attn_logits)
            s0 = sets[0]
            outf = torch.empty((batch, wl.Q, wl.num_heads * wl.channels), dtype=dt, device=dev)
            lib = cb._native.load()
            fargs = (s0["value"].data_ptr(), s0["spatial_shapes"].data_ptr(), s0["level_start_index"].data_ptr(), refp.data_ptr(),
                     off.data_ptr(), lg.data_ptr(), outf.data_ptr(), batch, wl.S, wl.num_heads, wl.channels, wl.L, wl.Q,
                     wl.num_points, refp.shape[-1], cb.ops._DTYPES[dt], 0)
            stream = torch.cuda.current_stream().cuda_stream
            us = time_calls([lambda: lib.msda_b200_forward_fused(*fargs, stream)], iters)
            row = {"workload": name, "batch": batch, "dtype": dtn, "loc_mode": wl.kind, "config": "fused-producers",
                   "variant": cb.last_variant(), "us_per_call": us, "images_per_s": batch / (us * 1e-6)}
            results["rows"].append(row)
            print(f"{name:24s} b{batch} {dtn:8s} {wl.kind:8s} {'fused-producers':20s} {us:9.2f} us  (L2-warm single input set)  {row['variant']}", flush=True)
        if have_ref and dtn in ("float16", "float32"):
            fns = [(lambda s=s: torch.ops.codetr_ref.msda_forward(*(s[k] for k in KEYS), 64)) for s in sets]
            us = time_calls(fns, max(20, iters // 2))
            row = {"workload": name, "batch": batch, "dtype": dtn, "loc_mode": loc_mode or wl.kind, "config": "reference_cuda_sm100a",
                   "variant": "codetr_ref::ms_deformable_im2col_gpu_kernel (+2 memsets, torch op overhead)", "us_per_call": us,
                   "hbm_GBps": hbm / us / 1e3, "gather_GBps": gather / us / 1e3, "images_per_s": batch / (us * 1e-6), "n_sets": len(sets)}
            results["rows"].append(row)
            print(f"{name:24s} b{batch} {dtn:8s} {row['loc_mode']:8s} {'reference_cuda':20s} {us:9.2f} us", flush=True)
            # parity of the two CUDA implementations on the same tensors (informational)
            a = cb.multi_scale_deformable_attention(*(sets[0][k] for k in KEYS)).float()
            b = torch.ops.codetr_ref.msda_forward(*(sets[0][k] for k in KEYS), 64).float()
            row["max_rel_vs_ours"] = float((a - b).abs().max() / b.abs().max())
        del sets
        torch.cuda.empty_cache()

    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(results, f, indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
