timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_c19.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_tests_c19.log
NBMF_MULTIFIT_TIMING=1 timeout 600 python tools/small_fit_bench.py > gpurun_out/r2_small_fit4.log 2>&1; echo "rc=$?"; grep -E "K=" gpurun_out/r2_small_fit4.log | cut -c1-300 | tail -24
timeout 900 python bench.py --configs > gpurun_out/r2_configs_c19.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/r2_configs_c19.log
