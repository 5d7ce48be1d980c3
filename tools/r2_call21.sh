timeout 900 python -m pytest tests/test_gpu_fused.py -x -q > gpurun_out/r2_tests_c21.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/r2_tests_c21.log
timeout 900 python bench.py --configs > gpurun_out/r2_configs_c21.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/r2_configs_c21.log | cut -c1-260
