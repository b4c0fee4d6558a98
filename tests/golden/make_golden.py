"""Generate the golden fixtures in this directory from the REFERENCE's own function.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It AST-extracts ``multi_scale_deformable_attention_pytorch`` from
/root/reference/codetr/ops.py (lines 129-186; the module itself cannot be imported because it
imports tensorrt / torch_tensorrt at the top), executes it on the seeded inputs defined in
``cases()`` and writes one ``<case>.npz`` per case holding the inputs (when small), a sha256 of the
inputs, and the reference outputs in float32 and float64.  The GPU box has no /root/reference; the
tests read only the .npz files.
"""
from __future__ import annotations

import ast
import hashlib
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

REFERENCE_OPS = "/root/reference/codetr/ops.py"
GRAD_CASES = ("ref_seed3", "edge_borders", "codino_dec_tiny", "odd_dims")


def load_reference_function():
    src = open(REFERENCE_OPS).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == "multi_scale_deformable_attention_pytorch":
            mod = ast.Module(body=[node], type_ignores=[])
            ns = {"torch": torch, "F": F, "Tensor": Tensor}
            exec(compile(mod, REFERENCE_OPS, "exec"), ns)
            return ns["multi_scale_deformable_attention_pytorch"], (node.lineno, node.end_lineno)
    raise RuntimeError("reference function not found")


def main() -> None:
    from golden_cases import cases, inputs_digest  # tests/golden_cases.py

    ref_fn, lines = load_reference_function()
    print(f"reference function: {REFERENCE_OPS}:{lines[0]}-{lines[1]}, torch {torch.__version__}")
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    for case in cases():
        arrs = case.build()
        shapes = torch.from_numpy(arrs["spatial_shapes"])
        outs = {}
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            v = torch.from_numpy(arrs["value"]).to(dt)
            loc = torch.from_numpy(arrs["sampling_loc"]).to(dt)
            w = torch.from_numpy(arrs["attn_weight"]).to(dt)
            with torch.no_grad():
                outs[tag] = ref_fn(v, shapes, loc, w).numpy()
        # gradients: autograd through the reference's own (differentiable) function, fp64, for a seeded
        # grad_output; the reference's backward kernel is tested against exactly this (tests:367-414)
        grads = {}
        if case.name in GRAD_CASES:
            g = np.random.default_rng(4242).standard_normal(outs["f64"].shape)
            v = torch.from_numpy(arrs["value"]).double().requires_grad_(True)
            loc = torch.from_numpy(arrs["sampling_loc"]).double().requires_grad_(True)
            w = torch.from_numpy(arrs["attn_weight"]).double().requires_grad_(True)
            ref_fn(v, shapes, loc, w).backward(torch.from_numpy(g))
            grads = {"grad_out": g, "grad_value": v.grad.numpy(), "grad_loc": loc.grad.numpy(), "grad_weight": w.grad.numpy()}
        payload = {
            **grads,
            "out_f32": outs["f32"],
            "out_f64": outs["f64"],
            "digest": np.frombuffer(inputs_digest(arrs).encode(), dtype=np.uint8),
            "torch_version": np.frombuffer(torch.__version__.encode(), dtype=np.uint8),
        }
        if case.store_inputs:
            payload.update({k: arrs[k] for k in ("value", "spatial_shapes", "level_start_index", "sampling_loc", "attn_weight")})
        path = os.path.join(HERE, case.name + ".npz")
        np.savez_compressed(path, **payload)
        print(f"{case.name:28s} out {outs['f32'].shape}  |out|max {np.abs(outs['f64']).max():.4g}  "
              f"f32-vs-f64 {np.abs(outs['f32'] - outs['f64']).max():.3g}  {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
