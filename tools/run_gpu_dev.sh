python tools/small_fit.py 32 200; python tools/small_fit.py 6 200; python tools/small_fit.py 64 200
timeout 900 python tools/multifit_bench.py > gpurun_out/multifit_bench.log 2>&1; echo rc=$?; cat gpurun_out/multifit_bench.log | tail -8
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest.log
timeout 1500 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_full_tc5.log 2> gpurun_out/bench_full_tc5.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_full_tc5.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_full_tc5.log').read().strip().splitlines()[-1]); r=d['roofline']
print('h_ms=%.2f w_ms=%.2f step=%.2f ms value=%.3e plan=%s clocks=%s'%(r['avg_launch_ms'], r['w_pass']['avg_launch_ms'], d['ms_per_step'], d['value'], d['config']['launch_plan'], d['clocks']))
PY
