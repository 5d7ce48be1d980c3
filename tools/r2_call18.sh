NBMF_MULTIFIT_TIMING=1 timeout 600 python tools/small_fit_bench.py > gpurun_out/r2_small_fit3.log 2>&1; echo "rc=$?"; grep -E "K=64|K=32" gpurun_out/r2_small_fit3.log | cut -c1-400 | tail -16
