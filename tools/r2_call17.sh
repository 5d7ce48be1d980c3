timeout 900 python -m pytest tests/test_gpu_multifit.py tests/test_gpu_tensor_engine.py tests/test_gpu_onestep.py tests/test_gpu_depth.py tests/test_gpu_trajectory.py -q -m gpu -x > gpurun_out/r2_pytest17.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest17.log
timeout 600 python tools/small_fit_bench.py > gpurun_out/r2_small_fit2.log 2>&1; echo "rc=$?"; cat gpurun_out/r2_small_fit2.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('cfg4 h_ms=%.2f w_ms=%.2f step=%.2f ms value=%.3e parity=%s' % (r['avg_launch_ms'], r['w_pass']['avg_launch_ms'], d['ms_per_step'], d['value'], d['parity_check']['ok']))"
