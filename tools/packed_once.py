"""A few launches of the packed-pyramid gather at the headline shape (for ncu): python tools/packed_once.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import codetr_b200 as cb
from codetr_b200 import workloads as W
from perf_sweep import KEYS, device_sets
dev = torch.device("cuda:0")
wl = W.CONFIGS["swinl_enc_1152x768"]
sets, _ = device_sets(wl, 1, torch.float16, dev)
packs = [cb.pack_value(s["value"], s["spatial_shapes"], s["level_start_index"]) for s in sets]
for i in range(8):
    s, pk = sets[i % len(sets)], packs[i % len(sets)]
    cb.forward_packed(pk, torch.float16, wl.S, s["spatial_shapes"], s["level_start_index"], s["sampling_loc"], s["attn_weight"])
torch.cuda.synchronize()
print(cb.last_variant())
