timeout 900 python -m pytest tests/test_gpu_ingest_eval.py tests/test_gpu_estimator.py tests/test_gpu_ref_suite.py -q -m gpu > gpurun_out/r2_tests_c36.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2_tests_c36.log
timeout 900 python bench.py --configs > gpurun_out/r2_configs_c36.log 2>&1; echo "rc=$?"; cut -c1-330 gpurun_out/r2_configs_c36.log
