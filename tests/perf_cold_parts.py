"""Which input's coldness costs what (manual, GPU box): the headline call timed with only some tensors rotated over
> L2 worth of copies (cold) while the others are reused (L2-warm).  python tests/perf_cold_parts.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

import codetr_b200 as cb
from codetr_b200 import workloads as W

KEYS = ("value", "spatial_shapes", "level_start_index", "sampling_loc", "attn_weight")


def main():
    dev = torch.device("cuda:0")
    name = sys.argv[1] if len(sys.argv) > 1 else "swinl_enc_1152x768"
    wl = W.CONFIGS[name]
    dt = torch.float16
    n = 16
    inp = W.make_inputs(wl, batch=1, seed=wl.seed)
    base = {k: torch.from_numpy(getattr(inp, k)) for k in KEYS}
    base = {k: (v.to(dev) if v.dtype == torch.int64 else v.to(device=dev, dtype=dt)) for k, v in base.items()}
    copies = {k: [base[k].clone() for _ in range(n)] for k in ("value", "sampling_loc", "attn_weight")}
    outs = [torch.empty((1, wl.Q, wl.num_heads * wl.channels), dtype=dt, device=dev) for _ in range(n)]
    for label, cold in (("all warm", ()), ("value cold", ("value",)), ("loc+weights cold", ("sampling_loc", "attn_weight")),
                        ("output cold", ("out",)), ("all cold", ("value", "sampling_loc", "attn_weight", "out"))):
        calls = []
        for i in range(n):
            pick = lambda k: copies[k][i] if k in cold else copies[k][0]
            calls.append(cb.PreparedForward(pick("value"), base["spatial_shapes"], base["level_start_index"], pick("sampling_loc"),
                                            pick("attn_weight"), output=outs[i] if "out" in cold else outs[0]))
        for i in range(2 * n):
            calls[i % n]()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = None
        for _ in range(3):
            s.record()
            for i in range(20 * n):
                calls[i % n]()
            e.record()
            torch.cuda.synchronize()
            us = 1e3 * s.elapsed_time(e) / (20 * n)
            best = us if best is None else min(best, us)
        print(f"{name} {label:18s} {best:7.2f} us  {cb.last_variant()}", flush=True)


if __name__ == "__main__":
    main()
