cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for api in cabi plugin torch_op; do
  for wl in "swinl_enc_1920x1280 2" "swinl_enc_1152x768 1" "swinl_dec_1152x768 1"; do
    set -- $wl
    python bench.py --workload $1 --batch $2 --api $api --steps 1000 --warmup 20 --no-cpu-baseline --no-e2e --no-batch-sweep | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$api', d['config']['workload'], 'b', d['config']['per_gpu_batch'], round(d['us_per_call'],2),'us/call', round(d['value']),'img/s | per-call', {k: round(v,2) if isinstance(v,float) else v for k,v in d['per_call_us'].items() if k!='note'})"
  done
done 2>&1 | tee gpurun_out/api_bench.log
