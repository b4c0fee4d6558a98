#!/bin/bash
# Gather-primitive ceilings on one B200 (tools/gather_probe.cu).  Usage: gpurun -- bash tools/run_gather_probe.sh
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
P=build_variants/gather_probe
OUT=gpurun_out/gather_probe.jsonl
: > $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader >> gpurun_out/gather_probe_env.txt
for mode in 0 1 2 7; do
  for cps in 2 4 6 8; do
    for u in 2 4 8; do
      timeout 60 $P $mode $cps $u 0 >> $OUT 2>&1
    done
  done
  # L1-hitting window (256 rows = 16 KB per lane group neighbourhood)
  timeout 60 $P $mode 4 4 256 >> $OUT 2>&1
  timeout 60 $P $mode 8 4 256 >> $OUT 2>&1
done
for mode in 3 4; do
  for u in 2 4 8; do timeout 60 $P $mode 1 $u >> $OUT 2>&1; done
done
for mode in 5 6; do
  for cps in 1 2 4; do
    for u in 2 4 8; do timeout 60 $P $mode $cps $u >> $OUT 2>&1; done
  done
done
cat $OUT | cut -c1-400
# wavefront accounting for the LDG / LDS shapes
M=l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sectors_op_read.sum,smsp__inst_executed.sum,sm__cycles_elapsed.max,l1tex__t_sector_hit_rate.pct,smsp__inst_executed_op_shared_ld.sum
for mode in 0 1 2 3 4 7; do
  cps=4; [ $mode -ge 3 ] && [ $mode -le 4 ] && cps=1
  timeout 120 ncu --metrics $M --clock-control none -s 1 -c 1 --csv $P $mode $cps 4 2>/dev/null | grep -v "^==" > gpurun_out/gather_probe_ncu_mode$mode.csv
done
python - <<'PY'
import csv, glob
for f in sorted(glob.glob('gpurun_out/gather_probe_ncu_mode*.csv')):
    rows = list(csv.DictReader(open(f)))
    print(f, {r['Metric Name']: r['Metric Value'] for r in rows if 'Metric Name' in r})
PY
