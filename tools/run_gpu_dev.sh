timeout 900 python tools/e2e_breakdown.py > gpurun_out/e2e_breakdown.log 2>&1; echo rc=$?; tail -3 gpurun_out/e2e_breakdown.log
