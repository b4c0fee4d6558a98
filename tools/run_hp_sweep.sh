#!/bin/bash
# head-pair kernel: warps-per-CTA / shared-memory sweep at the headline shape, both register builds
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for w in 25 24 21 16; do
  echo "== default build (<=800 threads), warps=$w"; MSDA_B200_HP_WARPS=$w HP_TAG=_w$w HP_SMEM_LIST=${SMEMS:-148,0} python tests/perf_hp.py headline 2>&1 | grep "hp smem"
done
for w in 32 29 25; do
  echo "== 1024-thread build, warps=$w"; MSDA_B200_LIB=$PWD/build_variants/libmsda_hp1024.so MSDA_B200_HP_WARPS=$w HP_TAG=_1024_w$w HP_SMEM_LIST=${SMEMS:-148,0} python tests/perf_hp.py headline 2>&1 | grep "hp smem"
done
echo "== all workloads, default build"; HP_SMEM_LIST=148,0 python tests/perf_hp.py all 2>&1 | tail -24
