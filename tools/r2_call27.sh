timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_c27.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2_tests_c27.log
timeout 600 python tools/fused_trace.py > gpurun_out/r2_fused_trace5.log 2>&1; echo "rc=$?"; grep -v "^\[nbmf" gpurun_out/r2_fused_trace5.log | cut -c1-250; grep "^\[nbmf" gpurun_out/r2_fused_trace5.log | awk 'NR%6==5 || NR%6==0' | cut -c1-250
timeout 900 python tools/fused_batch_bench.py > gpurun_out/r2_fused_batch.log 2>&1; echo "rc=$?"; cat gpurun_out/r2_fused_batch.log
