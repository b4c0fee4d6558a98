"""Host-side cost per call of each binding (manual, GPU box).  Tiny tensors, so the GPU is never the limit:
the figure is CPU microseconds per call through (a) the prepared C-ABI call, (b) the plugin-enqueue-shaped
entry, (c) the package's torch op (native or Python registration, whichever is active; run again with MSDA_B200_PYTHON_OP=1 for the other), (d) the drop-in native extension (the reference's own binding file
linked against this repo's adapter; separate process), (e) the reference's own extension (oracle/_ref)."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

N = 20000


def tensors():
    dev = torch.device("cuda:0")
    shapes = torch.tensor([[8, 8], [4, 4]], dtype=torch.int64, device=dev)
    lsi = torch.tensor([0, 64], dtype=torch.int64, device=dev)
    value = torch.rand(1, 80, 8, 32, device=dev, dtype=torch.float16)
    loc = torch.rand(1, 16, 8, 2, 4, 2, device=dev, dtype=torch.float16)
    w = torch.rand(1, 16, 8, 2, 4, device=dev, dtype=torch.float16)
    return value, shapes, lsi, loc, w


def host_us(fn, n=N):
    for _ in range(200):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    dt = time.perf_counter() - t0
    torch.cuda.synchronize()
    return 1e6 * dt / n


if len(sys.argv) > 1 and sys.argv[1] == "dropin":
    torch.ops.load_library(sys.argv[2])
    t = tensors()
    print("dropin_native_op_us", round(host_us(lambda: torch.ops.codetr.multi_scale_deformable_attention(*t, 64)), 2))
    sys.exit(0)

import codetr_b200 as cb
from oracle import build_ref

t = tensors()
prep = cb.PreparedForward(*t)
stream = torch.cuda.current_stream().cuda_stream
out = torch.empty(1, 16, 256, device="cuda:0", dtype=torch.float16)
ptrs = [x.data_ptr() for x in t]
res = {
    "prepared_cabi_call_us": host_us(lambda: prep(stream)),
    "plugin_enqueue_entry_us": host_us(lambda: cb.plugin_enqueue(t[0].shape, t[3].shape, 1, ptrs, out.data_ptr(), stream)),
    f"torch_op_{cb.ops.op_registration}_registration_us": host_us(lambda: torch.ops.codetr.multi_scale_deformable_attention(*t, 64)),
    "python_functional_api_us": host_us(lambda: cb.multi_scale_deformable_attention(*t)),
}
if build_ref.load_if_built():
    res["reference_native_op_us"] = host_us(lambda: torch.ops.codetr_ref.msda_forward(*t, 64))
dropin = os.path.join(ROOT, "co-detr-tensorrt_b200", "csrc", "_dropin", "codetr_cpp_extension.so")
if os.path.isfile(dropin):
    o = subprocess.run([sys.executable, __file__, "dropin", dropin], capture_output=True, text=True)
    for line in o.stdout.splitlines():
        if line.startswith("dropin_native_op_us"):
            res["dropin_native_op_us"] = float(line.split()[1])
for k, v in res.items():
    print(f"{k:32s} {v:8.2f} us/call (host)")
