timeout 900 python -m pytest tests/test_gpu_multifit.py tests/test_gpu_estimator.py -x -q > gpurun_out/r2_tests_c20.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_tests_c20.log
NBMF_MULTIFIT_TIMING=1 timeout 600 python tools/small_fit_bench.py > gpurun_out/r2_small_fit5.log 2>&1; echo "rc=$?"; grep -E "K=" gpurun_out/r2_small_fit5.log | cut -c1-300 | tail -24
timeout 900 python bench.py --configs > gpurun_out/r2_configs_c20.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/r2_configs_c20.log
