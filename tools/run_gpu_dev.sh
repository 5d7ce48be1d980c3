timeout 1500 python -m pytest tests/test_gpu_ingest_eval.py tests/test_gpu_estimator.py -q > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_new.log
