cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_ -s 12 -c 2 -f -o gpurun_out/prof_packed \
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --workspace > gpurun_out/ncu_packed.log 2>&1; echo "ncu packed exit $?"
MSDA_B200_HEAD_MAJOR=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_fwd -s 6 -c 1 -f -o gpurun_out/prof_qmajor \
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_qmajor.log 2>&1; echo "ncu qmajor exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_fwd -s 6 -c 1 -f -o gpurun_out/prof_dec \
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --workload swinl_dec_1152x768 > gpurun_out/ncu_dec.log 2>&1; echo "ncu dec exit $?"
