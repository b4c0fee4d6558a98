cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python tests/perf_sweep.py --only headline --no-probes --out gpurun_out/sweep_quick.json > gpurun_out/sweep_quick.log 2>&1; echo "sweep exit $?"; cut -c1-175 gpurun_out/sweep_quick.log
