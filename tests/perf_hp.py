"""A/B harness for the head-pair kernel (manual, GPU box):  python tests/perf_hp.py [workload-set]

For each workload: the all-global vector kernel (MSDA_B200_HP=0) against the head-pair kernel with several
shared-memory budgets, CUDA events over back-to-back launches on rotating (> L2) input sets, plus a bit-for-bit
comparison of the two outputs (same arithmetic, same order: they must be identical).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch

import codetr_b200 as cb
from codetr_b200 import workloads as W
from perf_sweep import KEYS, device_sets, time_calls

WORKLOADS = {
    "headline": [("swinl_enc_1152x768", 1, "float16", None)],
    "f32": [("swinl_enc_1152x768", 1, "float32", None), ("swinl_enc_1920x1280", 2, "float32", None),
            ("r50_enc_608", 1, "float32", None), ("swinl_enc_1152x768", 1, "bfloat16", None)],
    "bf16": [("swinl_enc_1152x768", 1, "bfloat16", None), ("r50_enc_608", 1, "bfloat16", None),
             ("swinl_enc_1920x1280", 2, "bfloat16", None), ("swinl_enc_1152x768_s4", 1, "bfloat16", None)],
    "all": [("swinl_enc_1152x768", 1, "float16", None), ("swinl_enc_1152x768", 1, "float16", "uniform"),
            ("swinl_enc_1152x768", 4, "float16", None), ("swinl_enc_1152x768", 1, "bfloat16", None),
            ("swinl_enc_1152x768", 1, "float32", None), ("swinl_enc_1920x1280", 2, "float32", None),
            ("r50_enc_608", 1, "float16", None), ("swinl_enc_1920x1280", 2, "float16", None),
            ("swinl_enc_1152x768_s4", 1, "float16", None), ("swinl_dec_1900q", 1, "float16", None)],
}


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "headline"
    cfgs = [{"name": "vec (HP=0)", "MSDA_B200_HP": 0}]
    for kb in (os.environ.get("HP_SMEM_LIST", "148,200,100,64,0").split(",")):
        cfgs.append({"name": f"hp smem{kb}K", "MSDA_B200_HP": 1, "MSDA_B200_HP_SMEM": int(kb) * 1024, "MSDA_B200_HP_MIN_QUADS_PER_WARP": 0,
                     "MSDA_B200_HP_EXACT": int(os.environ.get("HP_EXACT", "0")), "MSDA_B200_BF16_SPLIT": 0})
    if which == "bf16":
        cfgs.append({"name": "hp split smem0K", "MSDA_B200_HP": 1, "MSDA_B200_HP_SMEM": 0, "MSDA_B200_HP_MIN_QUADS_PER_WARP": 0, "MSDA_B200_BF16_SPLIT": 1})
    dev = torch.device("cuda:0")
    rows = []
    for name, batch, dtn, loc_mode in WORKLOADS[which]:
        wl = W.CONFIGS[name]
        dt = getattr(torch, dtn)
        sets, hbm = device_sets(wl, batch, dt, dev, loc_mode)
        iters = 200 if wl.Q * batch < 40000 else 60
        base = None
        for cfg in cfgs:
            for k, v in cfg.items():
                if k.startswith("MSDA_"):
                    os.environ[k] = str(v)
            calls = [cb.PreparedForward(*(s[k] for k in KEYS), flags=0) for s in sets]
            us = time_calls(calls, iters)
            out = calls[0]().clone()
            torch.cuda.synchronize()
            if base is None:
                base = out
            same = bool(torch.equal(out, base))
            maxdiff = float((out.float() - base.float()).abs().max())
            row = {"workload": name, "batch": batch, "dtype": dtn, "loc_mode": loc_mode or wl.kind, "config": cfg["name"],
                   "variant": cb.last_variant(), "us_per_call": us, "bit_identical_to_vec": same, "max_abs_diff": maxdiff}
            rows.append(row)
            print(f"{name:24s} b{batch} {dtn:8s} {row['loc_mode']:8s} {cfg['name']:16s} {us:9.2f} us  same={same} diff={maxdiff:.2e}  {row['variant']}", flush=True)
        del sets
        torch.cuda.empty_cache()
    out_path = os.path.join(ROOT, "gpurun_out", f"perf_hp_{which}{os.environ.get('HP_TAG', '')}.json")
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
