cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_fwd -s 6 -c 1 -f -o gpurun_out/prof_default \
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-batch-sweep > gpurun_out/ncu_default.log 2>&1; echo "ncu exit $?"
timeout 300 python bench.py --steps 2000 --warmup 20 --no-cpu-baseline --no-batch-sweep | cut -c1-300
