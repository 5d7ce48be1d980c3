timeout 900 python -m pytest tests/test_gpu_fused.py -x -q > gpurun_out/r2_tests_c23.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2_tests_c23.log
timeout 600 python tools/fused_trace.py > gpurun_out/r2_fused_trace2.log 2>&1; echo "rc=$?"; grep -v "^\[nbmf" gpurun_out/r2_fused_trace2.log | cut -c1-250; grep "^\[nbmf" gpurun_out/r2_fused_trace2.log | awk 'NR%3==0' | cut -c1-250
