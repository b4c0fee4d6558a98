# compute-sanitizer over the small parity tests (memcheck: OOB / misaligned; racecheck: the TMA-staged shared memory path)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
K='golden or half_matches or fused or packed or backward_fp64 or backward_lower or non_finite or plugin_enqueue'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_msda_gpu.py tests/test_backward_gpu.py -x -q -m gpu -k "$K" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?"; tail -6 gpurun_out/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_msda_gpu.py -x -q -m gpu -k "staged and (edge_borders or codino_enc_tiny)" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck exit $?"; tail -6 gpurun_out/sanitizer_racecheck.log
