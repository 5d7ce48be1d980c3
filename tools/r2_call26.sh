timeout 1500 python tools/fused_sweep.py > gpurun_out/r2_fused_sweep.log 2>&1; echo "rc=$?"; cat gpurun_out/r2_fused_sweep.log | cut -c1-200
