# Kernel experiment loop on a GPU box: cycle trace, tensor-engine parity tests, full-size bench line (no e2e / CPU legs).
TAG=${1:-exp}
timeout 120 tools/bin/tc_trace 2048 > gpurun_out/tc_trace_$TAG.log 2>&1; echo "trace rc=$?"; head -12 gpurun_out/tc_trace_$TAG.log
timeout 600 python -m pytest tests/test_gpu_tensor_engine.py tests/test_gpu_large.py -q -m gpu -x > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$TAG.log').read().strip().splitlines()[-1]); r=d['roofline']
print('h_ms=%.2f w_ms=%.2f step=%.2f ms value=%.3e loss=%s clocks=%s'%(r['avg_launch_ms'], r['w_pass']['avg_launch_ms'], d['ms_per_step'], d['value'], d['config']['loss_first_last'], d['clocks']))
PY
