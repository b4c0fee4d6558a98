#!/bin/bash
# ncu --set full of the head-pair kernel (and the all-global vector kernel beside it) at the headline shape
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for tag in "$@"; do
  case $tag in
    hp)    env="MSDA_B200_HP=1" ;;
    hp0)   env="MSDA_B200_HP=1 MSDA_B200_HP_SMEM=0" ;;
    vec)   env="MSDA_B200_HP=0" ;;
  esac
  env $env timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_fwd -s 5 -c 1 -f -o gpurun_out/prof_$tag \
    python tools/msda_once.py > gpurun_out/ncu_$tag.log 2>&1; echo "ncu $tag exit $?"
  python tools/ncu_summary.py gpurun_out/prof_$tag.ncu-rep > gpurun_out/ncu_${tag}_summary.txt; cat gpurun_out/ncu_${tag}_summary.txt
done
