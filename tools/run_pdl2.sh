cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for wl in "swinl_dec_1900q 1" "swinl_dec_1152x768 8" "ref_test_mid_fp32 1" "swinl_dec_1152x768 1" "swinl_enc_1920x1280 2"; do
  set -- $wl
  for pdl in 1; do
    MSDA_B200_PDL=$pdl python bench.py --workload $1 --batch $2 --steps 2000 --warmup 20 --no-cpu-baseline --no-e2e --no-batch-sweep | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1 b$2 pdl=$pdl', round(d['us_per_call'],2),'us/call', d['roofline']['kernel'][:40], {k: round(v,2) for k,v in d['per_call_us'].items() if k in ('median','min')})"
  done
done 2>&1 | tee gpurun_out/pdl_bench2.log
