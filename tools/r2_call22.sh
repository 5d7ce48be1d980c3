timeout 600 python tools/fused_trace.py > gpurun_out/r2_fused_trace.log 2>&1; echo "rc=$?"; cat gpurun_out/r2_fused_trace.log | cut -c1-250
