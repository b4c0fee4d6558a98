#!/bin/bash
# One GPU-box session, by named stage.  Usage (from the build container):
#     gpurun --timeout 1800 -- bash tools/gpu_round.sh [stage ...]
# Everything is written under gpurun_out/ (merged back by gpurun); the summaries worth keeping are copied to profiles/ by hand.
#   tests      pytest -m gpu                      smoke     __graft_entry__.smoke()
#   bench      bench.py (N=1) + reference arm     scaleN    bench.py under torchrun with N = $GPUS ranks (gpurun --gpus N)
#   ncu        launch list of the bench command + `--set full` capture of the headline kernel (+ profiles/traffic.json)
#   ncudec     `--set full` capture of the decoder call       ncuvec   same for the all-global vector kernel (MSDA_B200_HP=0)
#   ncudtypes  `--set full` captures of the headline shape in bf16 and fp32      hpdtypes   their A/B timings + bf16 accuracy
#   ncuall     `--set full` captures of configs[1] and configs[4] (feeds profiles/traffic.json)
#   hp         head-pair kernel A/B against the vector kernel on every workload (tests/perf_hp.py)
#   hpsweep    warps-per-CTA / shared-memory sweep of the head-pair kernel + every tuning build under build_variants/
#   sweep      tests/perf_sweep.py (all configurations, reference CUDA kernel beside ours)
#   probe      tools/gather_probe.cu row-gather ceilings (+ ncu wavefront accounting)
#   api        host cost per binding: C ABI / plugin entry / torch op on three workloads
#   sanitizer  compute-sanitizer memcheck + racecheck over the small parity tests
#   vproj / module   projection-kernel timing + phase trace + ncu; module-level timing
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
STAGES="${@:-tests smoke bench ncu}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/nvidia_smi.csv 2>&1
ncu_full() {  # tag, env assignments..., then the workload args of tools/msda_once.py
  local tag=$1; shift
  env "$@" timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_fwd -s 5 -c 1 -f -o gpurun_out/prof_$tag \
    python tools/msda_once.py $NCU_ARGS > gpurun_out/ncu_$tag.log 2>&1; echo "ncu $tag exit $?"
  python tools/ncu_summary.py gpurun_out/prof_$tag.ncu-rep > gpurun_out/ncu_${tag}_summary.txt; head -3 gpurun_out/ncu_${tag}_summary.txt
}
for st in $STAGES; do
  case $st in
    tests)
      timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
      tail -5 gpurun_out/pytest_gpu.log ;;
    smoke)
      timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -6 gpurun_out/smoke.log ;;
    bench)
      timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-1500 gpurun_out/bench.json
      timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_reference.json ;;
    scale*)
      N=${GPUS:-2}
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N \
        > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N exit $?"; cut -c1-1200 gpurun_out/bench_n$N.json ;;
    ncu)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
        python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra-workloads > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
      NCU_ARGS="swinl_enc_1152x768 float16 1" ncu_full headline
      python tools/update_traffic.py gpurun_out/prof_headline.ncu-rep swinl_enc_1152x768/float16/b1 0 "profiles/r02_ncu_headline_summary.txt (gpurun_out/prof_headline.ncu-rep)" ;;
    ncudec)
      NCU_ARGS="swinl_dec_1152x768 float16 1" ncu_full decoder
      python tools/update_traffic.py gpurun_out/prof_decoder.ncu-rep swinl_dec_1152x768/float16/b1 0 "profiles/r02_ncu_decoder_summary.txt (gpurun_out/prof_decoder.ncu-rep)" ;;
    ncuall)  # the remaining encoder configurations (captures for profiles/traffic.json; two reports stay under gpurun's 64 MiB)
      NCU_ARGS="r50_enc_608 float16 1" ncu_full r50
      NCU_ARGS="swinl_enc_1920x1280 float16 2" ncu_full enc1920 ;;
    ncuvec)
      NCU_ARGS="swinl_enc_1152x768 float16 1" ncu_full vec MSDA_B200_HP=0 ;;
    ncudtypes)
      NCU_ARGS="swinl_enc_1152x768 bfloat16 1" ncu_full bf16 MSDA_B200_HP=1
      NCU_ARGS="swinl_enc_1152x768 float32 1" ncu_full f32 MSDA_B200_HP=1 ;;
    hpdtypes)
      HP_EXACT=1 HP_TAG=_f32 HP_SMEM_LIST=148,0 timeout 300 python tests/perf_hp.py f32 2>&1 | tee gpurun_out/perf_hp_f32.log | tail -14
      HP_EXACT=1 HP_TAG=_bf16 HP_SMEM_LIST=0 timeout 300 python tests/perf_hp.py bf16 2>&1 | tee gpurun_out/perf_hp_bf16.log | tail -14
      python tools/bf16_check.py 2>&1 | tee gpurun_out/bf16_check.log | tail -14 ;;
    hp)
      HP_SMEM_LIST=${SMEMS:-148,0} timeout 900 python tests/perf_hp.py all 2>&1 | tee gpurun_out/perf_hp_all.log | tail -30 ;;
    hpsweep)
      for w in 25 24 21 16; do
        echo "== default build (<= 800 threads), warps=$w"; MSDA_B200_HP_WARPS=$w HP_TAG=_w$w HP_SMEM_LIST=${SMEMS:-148,0} python tests/perf_hp.py headline 2>&1 | grep "hp smem"
      done
      # tuning builds (python tools/build_variant.py NAME -DMSDA_HP_THREADS=.. -DMSDA_HP_DEPTH=..): every libmsda_*.so found
      for lib in build_variants/libmsda_*.so; do
        [ -f "$lib" ] || continue
        v=$(basename $lib .so); echo "== $v"
        MSDA_B200_LIB=$PWD/$lib HP_TAG=_$v HP_SMEM_LIST=${SMEMS:-0} python tests/perf_hp.py ${HP_SET:-headline} 2>&1 | grep "hp smem\|hp split"
      done ;;
    sweep)
      timeout 1200 python tests/perf_sweep.py --out gpurun_out/sweep.json > gpurun_out/sweep.log 2>&1; echo "sweep exit $?"; tail -3 gpurun_out/sweep.log ;;
    probe)
      bash tools/run_gather_probe.sh > gpurun_out/gather_probe.log 2>&1; tail -8 gpurun_out/gather_probe.log ;;
    api)
      for api in cabi plugin torch_op; do
        for wl in "swinl_enc_1920x1280 2" "swinl_enc_1152x768 1" "swinl_dec_1152x768 1"; do
          set -- $wl
          python bench.py --workload $1 --batch $2 --api $api --steps 1000 --warmup 20 --no-cpu-baseline --no-e2e --no-batch-sweep --no-neighbours --no-reference-cuda | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$api', d['config']['workload'], 'b', d['config']['per_gpu_batch'], round(d['us_per_call'],2),'us/call', round(d['value']),'img/s | per-call', {k: round(v,2) if isinstance(v,float) else v for k,v in d['per_call_us'].items() if k!='note'}, d['roofline']['kernel'])"
        done
      done 2>&1 | tee gpurun_out/api_bench.log
      timeout 300 python tests/perf_host_overhead.py 2>&1 | tee gpurun_out/host_overhead.log | tail -12 ;;
    sanitizer)
      K='golden or half_matches or fused or packed or backward_fp64 or backward_lower or non_finite or plugin_enqueue or dynamic_unit'
      timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_msda_gpu.py tests/test_backward_gpu.py -x -q -m gpu -k "$K" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?"; tail -4 gpurun_out/sanitizer_memcheck.log
      timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_hp_gpu.py -x -q -m gpu -k "not 200" > gpurun_out/sanitizer_hp_memcheck.log 2>&1; echo "hp memcheck exit $?"; tail -4 gpurun_out/sanitizer_hp_memcheck.log
      timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_hp_gpu.py -x -q -m gpu -k "bit_identical and 148 and f16" > gpurun_out/sanitizer_hp_racecheck.log 2>&1; echo "hp racecheck exit $?"; tail -4 gpurun_out/sanitizer_hp_racecheck.log
      timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_msda_gpu.py -x -q -m gpu -k "staged and (edge_borders or codino_enc_tiny)" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck exit $?"; tail -4 gpurun_out/sanitizer_racecheck.log
      timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_value_proj_gpu.py -x -q -m gpu -k "not 18414 and not 40000 and not 20000" > gpurun_out/sanitizer_value_proj.log 2>&1; echo "value_proj memcheck exit $?"; tail -4 gpurun_out/sanitizer_value_proj.log ;;
    vproj)
      timeout 300 python tests/perf_value_proj.py > gpurun_out/value_proj.log 2>&1; echo "value_proj perf exit $?"; tail -2 gpurun_out/value_proj.log | cut -c1-300
      timeout 300 ncu --set full --clock-control none --import-source on -k regex:value_proj_persistent -s 2 -c 4 -f -o gpurun_out/prof_value_proj \
        python tools/vproj_once.py > gpurun_out/ncu_vproj.log 2>&1; echo "ncu value_proj exit $?" ;;
    module)
      timeout 300 python tests/perf_module.py > gpurun_out/module.log 2>&1; echo "module perf exit $?"; cut -c1-200 gpurun_out/module.log | tail -7 ;;
  esac
done
ls -la gpurun_out | tail -12
