"""Loader (and in-tree builder) of the C-ABI library ``libmsda_b200.so``.

The library is compiled from ``csrc/msda_sm100.cu`` and ``csrc/value_proj_sm100.cu`` for ``sm_100a`` only and is
the *only* compute path of this package: if it cannot be loaded the package
raises -- there is deliberately no PyTorch/CPU fallback (the reference's module
falls back to ``multi_scale_deformable_attention_pytorch`` for CPU tensors,
/root/reference/codetr/multi_scale_deformable_attention.py:207-210; here that
function exists only as test infrastructure under ``oracle/``).
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
from typing import List, Optional

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
_REPO_ROOT = os.path.dirname(_PKG_DIR)
CSRC_DIR = os.path.join(_PKG_DIR, "csrc")
INCLUDE_DIR = os.path.join(_REPO_ROOT, "include")
# MSDA_B200_LIB lets tuning experiments load an alternative build of the same sources
LIB_PATH = os.environ.get("MSDA_B200_LIB") or os.path.join(CSRC_DIR, "libmsda_b200.so")
SOURCES = [os.path.join(CSRC_DIR, "msda_sm100.cu"), os.path.join(CSRC_DIR, "value_proj_sm100.cu")]
HEADERS = [os.path.join(INCLUDE_DIR, "msda_b200.h"), os.path.join(CSRC_DIR, "msda_internal.hpp"), os.path.join(CSRC_DIR, "msda_fwd_hp.cuh")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-diag-suppress", "177",
]

# every symbol include/msda_b200.h declares
EXPORTED_SYMBOLS = [
    "msda_b200_forward",
    "msda_b200_forward_ws",
    "msda_b200_workspace_bytes",
    "msda_b200_plugin_workspace_bytes",
    "msda_b200_plugin_enqueue",
    "msda_b200_forward_fused",
    "msda_b200_backward",
    "msda_b200_host_workspace_bytes",
    "msda_b200_forward_host",
    "msda_b200_abi_version",
    "msda_b200_error_string",
    "msda_b200_launch_count",
    "msda_b200_last_variant",
    "msda_b200_algorithmic_hbm_bytes",
    "msda_b200_algorithmic_gather_bytes",
    "msda_b200_read_probe",
    "msda_b200_packed_value_bytes",
    "msda_b200_pack_value",
    "msda_b200_forward_packed",
    "msda_b200_value_proj_supported",
    "msda_b200_value_proj",
    "msda_b200_output_proj",
]

DTYPE_F32, DTYPE_F16, DTYPE_BF16, DTYPE_F64 = 0, 1, 2, 3
FLAG_FORCE_GENERIC = 1 << 0
FLAG_LINEAR_ORDER = 1 << 1
FLAG_MATH_FHFMA = 1 << 2
FLAG_MATH_EXACT = 1 << 3
FLAG_NO_STAGING = 1 << 4
FLAG_STAGE_TMA = 1 << 5
FLAG_NO_PACKED = 1 << 6
FLAG_HEAD_MAJOR = 1 << 7
FLAG_PDL = 1 << 8
FLAG_NO_SMEM_LEVELS = 1 << 9


class NativeLibraryError(RuntimeError):
    pass


def _nvcc() -> str:
    cand = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    if os.path.isfile(cand):
        return cand
    found = shutil.which("nvcc")
    if not found:
        raise NativeLibraryError("nvcc not found; cannot build libmsda_b200.so")
    return found


def is_stale() -> bool:
    if os.environ.get("MSDA_B200_LIB"):
        return False
    if not os.path.isfile(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(f) > built for f in SOURCES + HEADERS)


def build_native(force: bool = False, verbose: bool = False, extra_flags: Optional[List[str]] = None) -> str:
    """Compile the CUDA sources in-tree for sm_100a (cross-compiles without a GPU)."""
    if not force and not is_stale():
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, *(extra_flags or []), "-I", INCLUDE_DIR, "-o", LIB_PATH, *SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    env = dict(os.environ)
    # the image exports CC/CXX wrappers that nvcc does not need; use the system host compiler
    proc = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if proc.returncode != 0:
        raise NativeLibraryError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    return LIB_PATH


TORCH_BINDING_DIR = os.path.join(CSRC_DIR, "_torch")
TORCH_BINDING_PATH = os.path.join(TORCH_BINDING_DIR, "codetr_b200_torch.so")
TORCH_BINDING_SOURCES = [os.path.join(CSRC_DIR, "codetr_torch_binding.cpp"), os.path.join(CSRC_DIR, "codetr_aten_adapter.cpp")]


def build_torch_binding(force: bool = False, verbose: bool = False) -> str:
    """Compile the optional native operator registration (codetr_torch_binding.cpp + the ATen adapter) in-tree
    with torch.utils.cpp_extension; links libmsda_b200.so through an $ORIGIN-relative rpath."""
    build_native()
    if not force and os.path.isfile(TORCH_BINDING_PATH) and all(
            os.path.getmtime(TORCH_BINDING_PATH) > os.path.getmtime(f) for f in TORCH_BINDING_SOURCES + HEADERS):
        return TORCH_BINDING_PATH
    os.makedirs(TORCH_BINDING_DIR, exist_ok=True)
    env_backup = {k: os.environ.get(k) for k in ("CC", "CXX", "TORCH_CUDA_ARCH_LIST")}
    os.environ["CC"], os.environ["CXX"] = "/usr/bin/gcc", "/usr/bin/g++"
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    try:
        from torch.utils.cpp_extension import load

        load(name="codetr_b200_torch", sources=TORCH_BINDING_SOURCES, extra_cflags=["-O2"],
             extra_include_paths=[INCLUDE_DIR], extra_ldflags=[f"-L{CSRC_DIR}", "-lmsda_b200", "-Wl,-rpath,\\$$ORIGIN/.."],
             build_directory=TORCH_BINDING_DIR, is_python_module=False, with_cuda=True, verbose=verbose)
    finally:
        for k, v in env_backup.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return TORCH_BINDING_PATH


_lib: Optional[ctypes.CDLL] = None


def load() -> ctypes.CDLL:
    """dlopen the library and declare the prototypes of include/msda_b200.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or codetr_b200.build_native()).  There is no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    missing = [s for s in EXPORTED_SYMBOLS if not hasattr(lib, s)]
    if missing:
        raise NativeLibraryError(f"{LIB_PATH} lacks symbols {missing}; rebuild it")
    vp, i64, ci, cu = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_uint
    lib.msda_b200_forward.restype = ci
    lib.msda_b200_forward.argtypes = [vp, vp, vp, vp, vp, vp, i64, i64, i64, i64, i64, i64, i64, i64, ci, cu, vp]
    lib.msda_b200_forward_fused.restype = ci
    lib.msda_b200_forward_fused.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, i64, i64, i64, i64, i64, i64, i64, ci, cu, vp]
    lib.msda_b200_forward_ws.restype = ci
    lib.msda_b200_forward_ws.argtypes = [vp, vp, vp, vp, vp, vp, vp, ctypes.c_size_t,
                                         i64, i64, i64, i64, i64, i64, i64, i64, ci, cu, vp]
    lib.msda_b200_workspace_bytes.restype = ctypes.c_size_t
    lib.msda_b200_workspace_bytes.argtypes = [i64, i64, i64, i64, i64, i64, i64, ci]
    lib.msda_b200_plugin_workspace_bytes.restype = ctypes.c_size_t
    lib.msda_b200_plugin_workspace_bytes.argtypes = [vp, vp, ci]
    lib.msda_b200_backward.restype = ci
    lib.msda_b200_backward.argtypes = [vp] * 9 + [i64] * 8 + [ci, cu, vp]
    lib.msda_b200_plugin_enqueue.restype = ci
    lib.msda_b200_plugin_enqueue.argtypes = [vp, vp, ci, vp, vp, vp, ctypes.c_size_t, i64, vp]
    lib.msda_b200_host_workspace_bytes.restype = ctypes.c_size_t
    lib.msda_b200_host_workspace_bytes.argtypes = [i64, i64, i64, i64, i64, i64, i64, ci]
    lib.msda_b200_forward_host.restype = ci
    lib.msda_b200_forward_host.argtypes = [vp, vp, vp, vp, vp, vp, vp, ctypes.c_size_t,
                                           i64, i64, i64, i64, i64, i64, i64, i64, ci, cu, vp]
    lib.msda_b200_abi_version.restype = ci
    lib.msda_b200_error_string.restype = ctypes.c_char_p
    lib.msda_b200_error_string.argtypes = [ci]
    lib.msda_b200_launch_count.restype = ctypes.c_uint64
    lib.msda_b200_last_variant.restype = ctypes.c_char_p
    lib.msda_b200_algorithmic_hbm_bytes.restype = ctypes.c_uint64
    lib.msda_b200_algorithmic_hbm_bytes.argtypes = [i64, i64, i64, i64, i64, i64, i64, ci]
    lib.msda_b200_algorithmic_gather_bytes.restype = ctypes.c_uint64
    lib.msda_b200_algorithmic_gather_bytes.argtypes = [i64, i64, i64, i64, i64, i64, ci]
    lib.msda_b200_read_probe.restype = ci
    lib.msda_b200_read_probe.argtypes = [vp, ctypes.c_size_t, ci, vp, vp]
    lib.msda_b200_packed_value_bytes.restype = ctypes.c_size_t
    lib.msda_b200_packed_value_bytes.argtypes = [i64, i64, i64, i64, ci]
    lib.msda_b200_pack_value.restype = ci
    lib.msda_b200_pack_value.argtypes = [vp, vp, vp, vp, i64, i64, i64, i64, i64, ci, vp]
    lib.msda_b200_forward_packed.restype = ci
    lib.msda_b200_forward_packed.argtypes = [vp, vp, vp, vp, vp, vp, i64, i64, i64, i64, i64, i64, i64, ci, cu, vp]
    lib.msda_b200_value_proj_supported.restype = ci
    lib.msda_b200_value_proj_supported.argtypes = [i64, i64, ci]
    lib.msda_b200_value_proj.restype = ci
    lib.msda_b200_value_proj.argtypes = [vp, vp, vp, vp, vp, i64, i64, i64, ci, cu, vp]
    lib.msda_b200_output_proj.restype = ci
    lib.msda_b200_output_proj.argtypes = [vp, vp, vp, vp, vp, i64, i64, i64, ci, cu, vp]
    if lib.msda_b200_abi_version() != 1:
        raise NativeLibraryError("libmsda_b200.so ABI version mismatch; rebuild it")
    _lib = lib
    return lib


def error_string(code: int) -> str:
    return load().msda_b200_error_string(int(code)).decode()


def launch_count() -> int:
    return int(load().msda_b200_launch_count())


def last_variant() -> str:
    return load().msda_b200_last_variant().decode()
