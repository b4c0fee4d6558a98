#!/bin/bash
# One GPU-box session: parity tests, smoke, tuning sweep, bench line, ncu launch list + full capture.
# Usage (from the build container):  gpurun --timeout 1800 -- bash tools/gpu_round.sh [stages...]
# Everything is written under gpurun_out/ (merged back by gpurun).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
STAGES="${@:-tests smoke sweep bench ncu vproj module}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/nvidia_smi.csv 2>&1
for st in $STAGES; do
  case $st in
    tests)
      timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
      tail -5 gpurun_out/pytest_gpu.log ;;
    smoke)
      timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/smoke.log ;;
    variants)
      bash tests/perf_variants.sh > /dev/null 2>&1; echo "variants exit $?"; tail -5 gpurun_out/variants.log ;;
    sweep)
      timeout 900 python tests/perf_sweep.py --out gpurun_out/sweep.json > gpurun_out/sweep.log 2>&1; echo "sweep exit $?"; tail -3 gpurun_out/sweep.log ;;
    bench)
      timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json
      timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_reference.json ;;
    vproj)
      timeout 300 python tests/perf_value_proj.py > gpurun_out/value_proj.log 2>&1; echo "value_proj perf exit $?"; tail -2 gpurun_out/value_proj.log | cut -c1-300
      if [ -f build_variants/vproj_trace.so ]; then
        (export MSDA_B200_LIB=$PWD/build_variants/vproj_trace.so
         echo "== kernel phases, 18414 rows (single wave)"; timeout 100 python tools/vproj_trace.py 18414 | tail -9
         echo "== kernel phases, 102300 rows"; timeout 100 python tools/vproj_trace.py 102300 | tail -9
         export MSDA_B200_LIB=$PWD/build_variants/vproj_trace_epi.so VPROJ_TRACE_EPI=1
         echo "== epilogue detail of tile 0, 18414 rows"; timeout 100 python tools/vproj_trace.py 18414 | tail -7) > gpurun_out/vproj_trace.log 2>&1
      fi
      timeout 300 ncu --set full --clock-control none --import-source on -k regex:value_proj_persistent -s 2 -c 4 -f -o gpurun_out/prof_value_proj \
        python tools/vproj_once.py > gpurun_out/ncu_vproj.log 2>&1; echo "ncu value_proj exit $?" ;;
    module)
      timeout 300 python tests/perf_module.py > gpurun_out/module.log 2>&1; echo "module perf exit $?"; cut -c1-200 gpurun_out/module.log | tail -7 ;;
    ncu)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
        python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_fwd -s 6 -c 2 -f -o gpurun_out/prof_headline \
        python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_fwd -s 6 -c 2 -f -o gpurun_out/prof_headline_fhfma \
        python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --flags 4 > gpurun_out/ncu_full_fhfma.log 2>&1; echo "ncu full fhfma exit $?" ;;
  esac
done
ls -la gpurun_out | tail -20
