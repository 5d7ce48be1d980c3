timeout 900 python -m pytest tests/test_gpu_multifit.py tests/test_gpu_fused.py tests/test_gpu_estimator.py -q -m gpu > gpurun_out/r2_tests_c37.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2_tests_c37.log
NBMF_MULTIFIT_TIMING=1 timeout 600 python tools/small_fit_bench.py > gpurun_out/r2_small_fit7.log 2>&1; echo "rc=$?"; grep -E "^K=" gpurun_out/r2_small_fit7.log | cut -c1-300; grep "batch of 64, K=32\|batch of 64, K=64" gpurun_out/r2_small_fit7.log | tail -6 | cut -c1-250
timeout 900 python bench.py --configs > gpurun_out/r2_configs_c37.log 2>&1; echo "rc=$?"; grep cfg5 gpurun_out/r2_configs_c37.log | cut -c1-200
timeout 900 python bench.py --configs > gpurun_out/r2_configs_c37b.log 2>&1; echo "rc=$?"; grep cfg5 gpurun_out/r2_configs_c37b.log | cut -c1-200
