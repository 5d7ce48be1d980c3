timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_tests_c29.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2_tests_c29.log
