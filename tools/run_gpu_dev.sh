set -x
tools/bin/tc_trace 512 > gpurun_out/tc_trace.log 2>&1; echo rc=$?; head -12 gpurun_out/tc_trace.log
timeout 600 python -m pytest tests/test_gpu_tensor_engine.py -x -q > gpurun_out/pytest_tc.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_tc.log
timeout 600 python bench.py --rows 200000 --cols 100000 --steps 3 --warmup 2 --no-e2e --no-cpu > gpurun_out/bench_tc2.log 2> gpurun_out/bench_tc2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_tc2.log').read().strip().splitlines()[-1]); r=d['roofline']
print('h_ms=%.2f w_ms=%.2f step=%.2f ms value=%.3e loss=%s clocks=%s'%(r['avg_launch_ms'], r['w_pass']['avg_launch_ms'], d['ms_per_step'], d['value'], d['config']['loss_first_last'], d['clocks']))
PY
