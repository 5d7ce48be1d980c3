timeout 900 python -m pytest tests/test_gpu_multifit.py tests/test_gpu_estimator.py -q -m gpu > gpurun_out/r2_tests_c38.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_tests_c38.log
NBMF_MULTIFIT_TIMING=1 timeout 600 python tools/small_fit_bench.py > gpurun_out/r2_small_fit8.log 2>&1; echo "rc=$?"; grep -E "^K=" gpurun_out/r2_small_fit8.log | cut -c1-300
for i in 1 2 3; do timeout 900 python bench.py --configs > gpurun_out/r2_configs_c38_$i.log 2>&1; echo "rc=$?"; grep cfg5 gpurun_out/r2_configs_c38_$i.log | cut -c1-120; done
