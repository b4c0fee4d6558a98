cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for extra in "" "--cuda-graph" "--l2-warm" "--l2-warm --cuda-graph"; do
  for wlb in "swinl_dec_1152x768 1" "swinl_dec_1152x768 8"; do
    set -- $wlb
    echo "== $1 b$2 $extra"
    python bench.py --workload $1 --batch $2 --steps 2000 --warmup 20 --no-cpu-baseline --no-e2e --no-batch-sweep $extra | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['us_per_call'],2),'us/call', round(d['value']),'img/s', d['config']['launch'][:12], '|', d['config']['l2_policy'][:30], '| clocks', d['clocks'])"
  done
done 2>&1 | tee gpurun_out/dec_bench.log
