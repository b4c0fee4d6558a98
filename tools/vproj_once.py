"""A few eager calls of the projection kernel for profiler captures (GPU box):
ncu --set full -k regex:value_proj_persistent ... python tools/vproj_once.py [rows ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import codetr_b200 as cb

dev = torch.device("cuda:0")
w = (torch.randn(256, 256, device=dev) / 16).half()
b = torch.randn(256, device=dev).half()
for rows in [int(a) for a in sys.argv[1:]] or [18414, 102300]:
    x = torch.randn(1, rows, 256, device=dev).half()
    m = torch.zeros(1, rows, dtype=torch.bool, device=dev)
    m[:, -rows // 10:] = True
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for _ in range(3):
        flush.zero_()  # evict x from L2 between calls
        cb.value_proj(x, w, b, m)
    torch.cuda.synchronize()
    print(rows, cb.last_variant())
