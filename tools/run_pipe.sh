cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
MSDA_B200_LIB=$PWD/build_variants/libmsda_pipe1_minb3.so timeout 900 python -m pytest tests/test_msda_gpu.py -x -q -m gpu 2>&1 | tail -3
bash tests/perf_variants.sh ctas > /dev/null 2>&1; grep -E "^===| default " gpurun_out/variants.log | sed "s/hbm.*err/err/" | cut -c1-100
