cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for wl in "swinl_enc_1152x768 1" "swinl_dec_1152x768 1" "r50_enc_608 1" "swinl_enc_1152x768 4"; do
  set -- $wl
  for mode in "off 0 0" "default 1 0" "early-tables 1 256"; do
    set -- $wl $mode
    MSDA_B200_PDL=$4 python bench.py --workload $1 --batch $2 --flags $5 --steps 2000 --warmup 20 --no-cpu-baseline --no-e2e --no-batch-sweep | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1 b$2 pdl=$3', round(d['us_per_call'],2),'us/call', round(d['value']),'img/s')"
  done
done 2>&1 | tee gpurun_out/pdl_bench.log
MSDA_B200_PDL=1 python bench.py --workload swinl_dec_1152x768 --cuda-graph --steps 2000 --warmup 20 --no-cpu-baseline --no-e2e --no-batch-sweep | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('decoder cuda-graph pdl default', round(d['us_per_call'],2))" | tee -a gpurun_out/pdl_bench.log
