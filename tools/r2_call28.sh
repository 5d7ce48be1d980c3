timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_c28.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2_tests_c28.log
timeout 900 python bench.py --configs > gpurun_out/r2_configs_c28.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/r2_configs_c28.log | cut -c1-330
