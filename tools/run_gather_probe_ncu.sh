#!/bin/bash
# ncu accounting of the gather probe's LDG / LDS shapes: wavefronts, LSU write-back, cycles -- the evidence behind
# "global loads return about 64 bytes per clock per SM whatever their shape" (DESIGN.md section 5).
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
P=build_variants/gather_probe
M=l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__lsu_writeback_active.sum,l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__m_xbar2l1tex_read_bytes.sum,l1tex__t_sector_hit_rate.pct,sm__cycles_elapsed.max,smsp__inst_executed.sum,gpu__time_duration.sum
OUT=gpurun_out/gather_probe_ncu.txt
: > $OUT
for cfg in "0 4 4" "1 4 4" "2 8 4" "7 8 8" "3 1 8" "4 1 8"; do
  set -- $cfg
  echo "## mode $1 ctas_per_sm $2 unroll $3" >> $OUT
  timeout 120 ncu --metrics $M --clock-control none -s 1 -c 1 $P $1 $2 $3 2>&1 | grep -E "^\s+(l1tex|sm__|smsp__|gpu__)|\{\"mode\"" | sed -E 's/\s+/ /g' >> $OUT
done
cat $OUT
