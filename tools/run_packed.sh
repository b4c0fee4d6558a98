cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_msda_gpu.py -x -q -m gpu -k "packed or golden or half_matches or plugin" > gpurun_out/pytest_packed.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_packed.log
timeout 600 python tests/perf_sweep.py --only headline --no-probes --out gpurun_out/sweep_packed.json > gpurun_out/sweep_packed.log 2>&1; echo "sweep exit $?"; cut -c1-175 gpurun_out/sweep_packed.log
