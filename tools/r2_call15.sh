timeout 600 python -m pytest tests/test_gpu_multifit.py tests/test_gpu_tensor_engine.py tests/test_gpu_edge_cases.py -q -m gpu -x > gpurun_out/r2_pytest15.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest15.log
timeout 600 python tools/small_fit_bench.py > gpurun_out/r2_small_fit.log 2>&1; echo "rc=$?"; cat gpurun_out/r2_small_fit.log
