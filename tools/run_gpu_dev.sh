timeout 1500 python -m pytest tests/test_gpu_multifit.py tests/test_gpu_estimator.py -q > gpurun_out/pytest_mf.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_mf.log
timeout 900 python tools/multifit_bench.py > gpurun_out/multifit_bench.log 2>&1; echo rc=$?; cat gpurun_out/multifit_bench.log | tail -8
