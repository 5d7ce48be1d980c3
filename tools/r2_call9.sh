timeout 600 python tools/batch_hint_bench.py > gpurun_out/r2_batch_hint.log 2>&1; echo "rc=$?"; cat gpurun_out/r2_batch_hint.log
