"""Build the drop-in ``codetr_cpp_extension.so`` (build container only; needs /root/reference).

It proves the boundary: the reference's own, UNCHANGED torch binding
(/root/reference/codetr/csrc/deformable_attention_torch.cpp, compiled from where it lies) links against
this repo's ATen adapter (co-detr-tensorrt_b200/csrc/codetr_aten_adapter.cpp) + libmsda_b200.so instead of
the reference's ms_deform_attn.cu.  Output: co-detr-tensorrt_b200/csrc/_dropin/codetr_cpp_extension.so --
the file name codetr/__init__.py:8-11 loads.  tests/test_dropin_gpu.py exercises it in a subprocess.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "co-detr-tensorrt_b200", "csrc")
OUT = os.path.join(CSRC, "_dropin")
REF_BINDING = "/root/reference/codetr/csrc/deformable_attention_torch.cpp"


def build(verbose=False):
    if not os.path.isfile(REF_BINDING):
        raise FileNotFoundError(REF_BINDING)
    target = os.path.join(OUT, "codetr_cpp_extension.so")
    srcs = [REF_BINDING, os.path.join(CSRC, "codetr_aten_adapter.cpp")]
    deps = srcs + [os.path.join(CSRC, "libmsda_b200.so")]
    if os.path.isfile(target) and all(os.path.getmtime(target) > os.path.getmtime(d) for d in deps[:2]):
        return target
    os.makedirs(OUT, exist_ok=True)
    os.environ["CC"], os.environ["CXX"] = "/usr/bin/gcc", "/usr/bin/g++"
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    from torch.utils.cpp_extension import load

    load(name="codetr_cpp_extension", sources=srcs, extra_cflags=["-O2"], extra_include_paths=[os.path.join(ROOT, "include")],
         extra_ldflags=[f"-L{CSRC}", "-lmsda_b200", "-Wl,-rpath,\\$$ORIGIN/.."], build_directory=OUT, is_python_module=False,
         with_cuda=True, verbose=verbose)
    return target


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
