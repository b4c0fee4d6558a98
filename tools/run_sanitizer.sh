# compute-sanitizer over the small parity tests (memcheck: OOB / misaligned; racecheck: the TMA-staged shared memory path)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
K='golden or half_matches or fused or packed or backward_fp64 or backward_lower or non_finite or plugin_enqueue or dynamic_unit'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_msda_gpu.py tests/test_backward_gpu.py -x -q -m gpu -k "$K" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?"; tail -6 gpurun_out/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_msda_gpu.py -x -q -m gpu -k "staged and (edge_borders or codino_enc_tiny)" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck exit $?"; tail -6 gpurun_out/sanitizer_racecheck.log
# the tcgen05 / TMA projection kernel (small shapes; the 18,414-row and 40,000-row cases are skipped: the sanitizer serialises CTAs)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_value_proj_gpu.py -x -q -m gpu -k "not 18414 and not 40000 and not 20000" > gpurun_out/sanitizer_value_proj.log 2>&1; echo "value_proj memcheck exit $?"; tail -4 gpurun_out/sanitizer_value_proj.log
