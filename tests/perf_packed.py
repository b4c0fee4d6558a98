"""Packed-pyramid gather A/B (manual, GPU box): python tests/perf_packed.py
hp kernel on the plain tensor vs the packed gather (pre-pass timed separately), rotating cold inputs, bit-identity."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import codetr_b200 as cb
from codetr_b200 import workloads as W
from perf_sweep import KEYS, device_sets, time_calls

dev = torch.device("cuda:0")
for name, batch, dtn in (("swinl_enc_1152x768", 1, "float16"), ("swinl_enc_1152x768", 4, "float16"), ("swinl_enc_1152x768", 1, "bfloat16"),
                         ("r50_enc_608", 1, "float16"), ("swinl_enc_1920x1280", 2, "float16"), ("swinl_dec_1900q", 1, "float16")):
    wl = W.CONFIGS[name]; dt = getattr(torch, dtn)
    sets, _ = device_sets(wl, batch, dt, dev)
    iters = 200 if wl.Q * batch < 40000 else 60
    plain = [cb.PreparedForward(*(s[k] for k in KEYS), flags=0) for s in sets]
    us_plain = time_calls(plain, iters); v_plain = cb.last_variant()
    want = plain[0]().clone()
    packs = [cb.pack_value(s["value"], s["spatial_shapes"], s["level_start_index"]) for s in sets]
    us_pack = time_calls([(lambda s=s: cb.pack_value(s["value"], s["spatial_shapes"], s["level_start_index"])) for s in sets], iters)
    outs = [torch.empty_like(want) for _ in sets]
    fns = [(lambda s=s, pk=pk, o=o: cb.forward_packed(pk, dt, wl.S, s["spatial_shapes"], s["level_start_index"], s["sampling_loc"], s["attn_weight"], out=o))
           for s, pk, o in zip(sets, packs, outs)]
    us_g = time_calls(fns, iters); v_g = cb.last_variant()
    got = fns[0](); torch.cuda.synchronize()
    print(f"{name:22s} b{batch} {dtn:8s} plain {us_plain:8.2f} us ({v_plain}) | packed gather {us_g:8.2f} us + pre-pass {us_pack:6.2f} us ({v_g}) same={bool(torch.equal(got, want))}", flush=True)
    del sets, packs
    torch.cuda.empty_cache()
