"""Compile the REFERENCE's CUDA op for sm_100a into oracle/_ref/ (build container only).

    python oracle/build_ref.py

Sources are compiled where they lie under /root/reference (nothing is copied into this repo):
  /root/reference/codetr/csrc/ms_deform_attn.cu     the reference kernel + ATen launcher, unmodified
  oracle/ref_binding.cpp                            ~40 lines exposing it as torch.ops.codetr_ref.msda_forward / msda_backward
with the reference's own extension flags (-O3 --use_fast_math, /root/reference/setup.py:71) and
-gencode arch=compute_100a,code=sm_100a instead of its sm_89 default (setup.py:5-22).  The reference's
build system (setup.py / CMake) is not run.  Output: oracle/_ref/msda_ref_cuda.so -- git-ignored, but it
travels to the GPU box with the snapshot, where it is loaded with torch.ops.load_library().
"""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
REF_CU = "/root/reference/codetr/csrc/ms_deform_attn.cu"
LIB_NAME = "msda_ref_cuda"
LIB_PATH = os.path.join(OUT_DIR, LIB_NAME + ".so")


def build(verbose: bool = False) -> str:
    if not os.path.isfile(REF_CU):
        raise FileNotFoundError(f"{REF_CU} not present (only the build container has the reference)")
    if os.path.isfile(LIB_PATH) and os.path.getmtime(LIB_PATH) > max(os.path.getmtime(REF_CU), os.path.getmtime(os.path.join(HERE, "ref_binding.cpp"))):
        return LIB_PATH
    os.makedirs(OUT_DIR, exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    os.environ["CC"] = "/usr/bin/gcc"
    os.environ["CXX"] = "/usr/bin/g++"
    from torch.utils.cpp_extension import load

    load(
        name=LIB_NAME,
        sources=[REF_CU, os.path.join(HERE, "ref_binding.cpp")],
        extra_cflags=["-O3"],
        extra_cuda_cflags=["-O3", "--use_fast_math", "-lineinfo"],
        build_directory=OUT_DIR,
        is_python_module=False,
        verbose=verbose,
    )
    return LIB_PATH


def load_if_built() -> bool:
    """Load oracle/_ref/msda_ref_cuda.so if it exists; registers torch.ops.codetr_ref.*"""
    import torch

    if hasattr(torch.ops, "codetr_ref") and hasattr(torch.ops.codetr_ref, "msda_forward"):
        return True
    if not os.path.isfile(LIB_PATH):
        return False
    torch.ops.load_library(LIB_PATH)
    return True


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
