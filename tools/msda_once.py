"""Launch the sampling kernel a few times on one workload (for ncu):  python tools/msda_once.py [workload] [dtype] [batch] [n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import codetr_b200 as cb
from codetr_b200 import workloads as W
from perf_sweep import KEYS, device_sets
name = sys.argv[1] if len(sys.argv) > 1 else "swinl_enc_1152x768"
dt = getattr(torch, sys.argv[2] if len(sys.argv) > 2 else "float16")
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 1
n = int(sys.argv[4]) if len(sys.argv) > 4 else 8
dev = torch.device("cuda:0")
sets, _ = device_sets(W.CONFIGS[name], batch, dt, dev)
calls = [cb.PreparedForward(*(s[k] for k in KEYS), flags=int(os.environ.get("MSDA_FLAGS", "0"))) for s in sets]
for i in range(n):
    calls[i % len(calls)]()
torch.cuda.synchronize()
print(cb.last_variant())
