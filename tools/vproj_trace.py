"""Phase timeline of the persistent value_proj kernel (debug build with -DMSDA_VPROJ_TRACE, see value_proj_sm100.cu).
usage (GPU box): MSDA_B200_LIB=build_variants/vproj_trace.so python tools/vproj_trace.py [rows]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import codetr_b200 as cb

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 18414
dev = torch.device("cuda:0")
x = torch.randn(1, rows, 256, device=dev).half()
w = (torch.randn(256, 256, device=dev) / 16).half()
b = torch.randn(256, device=dev).half()
m = torch.zeros(1, rows, dtype=torch.bool, device=dev)
lib = cb._native.load()
lib.msda_b200_debug_vproj_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
for it in range(4):
    xs = torch.randn_like(x)
    torch.cuda.synchronize()
    cb.value_proj(xs, w, b, m)
    torch.cuda.synchronize()
ctas = min(148, (rows + 127) // 128)
buf = np.zeros((ctas, 8), dtype=np.uint64)
assert lib.msda_b200_debug_vproj_trace(buf.ctypes.data, ctas) == 0
t = buf.astype(np.int64)
t0 = t[:, 0].min()
if os.environ.get("VPROJ_TRACE_EPI"):
    names = ["acc ready", "tmem ld done (chunk 0)", "slab free (prev store read)", "slab written + fence", "tma store issued", None, None,
             "tile 0 epilogue done"]
else:
    names = ["entry", "setup done", "weights landed", "first x chunk", "tile0 MMAs issued", "tile0 acc ready", "tile0 stored", "exit"]
print(f"rows={rows} ctas={ctas}  (ns after the first CTA's entry; median / min / max over CTAs)")
for i, n in enumerate(names):
    if n is None:
        continue
    d = t[:, i] - t0
    print(f"  {n:20s} {int(np.median(d)):7d} {int(d.min()):7d} {int(d.max()):7d}")
