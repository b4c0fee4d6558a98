"""Build a tuning variant of libmsda_b200.so into build_variants/ (never the in-tree product library):
    python tools/build_variant.py NAME -DMSDA_HP_THREADS=640 -DMSDA_HP_DEPTH=2 ...
Load it with MSDA_B200_LIB=$PWD/build_variants/libmsda_NAME.so (tests/perf_hp.py, tools/gpu_round.sh hpsweep)."""
import importlib.util
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("_n", os.path.join(ROOT, "co-detr-tensorrt_b200", "_native.py"))
n = importlib.util.module_from_spec(spec)
spec.loader.exec_module(n)
name, flags = sys.argv[1], sys.argv[2:]
out = os.path.join(ROOT, "build_variants", f"libmsda_{name}.so")
os.makedirs(os.path.dirname(out), exist_ok=True)
cmd = [n._nvcc(), *n.NVCC_FLAGS, *flags, "-I", n.INCLUDE_DIR, "-o", out, *n.SOURCES]
subprocess.run(cmd, check=True)
print(out)
