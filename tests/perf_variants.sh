#!/bin/bash
# Runs a short sweep once per alternative build of libmsda_b200.so found in build_variants/
# (same sources, different -D tuning macros), each in its own process.  usage: perf_variants.sh [only-mode]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
MODE=${1:-ctas}
rm -f gpurun_out/variants.log
for lib in co-detr-tensorrt_b200/csrc/libmsda_b200.so build_variants/*.so; do
  name=$(basename $lib .so)
  echo "=== $name" | tee -a gpurun_out/variants.log
  MSDA_B200_LIB=$PWD/$lib timeout 600 python tests/perf_sweep.py --only $MODE --no-probes --out gpurun_out/sweep_$name.json 2>&1 | grep -v "^wrote" | cut -c1-170 | tee -a gpurun_out/variants.log
done
