cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; rm -f gpurun_out/variants_dec.log
timeout 900 python -m pytest tests/test_msda_gpu.py -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
for lib in co-detr-tensorrt_b200/csrc/libmsda_b200.so build_variants/*.so; do
  name=$(basename $lib .so); echo "=== $name" >> gpurun_out/variants_dec.log
  MSDA_B200_LIB=$PWD/$lib timeout 600 python tests/perf_sweep.py --only decoder --no-probes --out gpurun_out/sweep_dec_$name.json 2>&1 | grep -v "^wrote" | cut -c1-150 >> gpurun_out/variants_dec.log
done
