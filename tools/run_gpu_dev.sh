# Development loop on a GPU box (`gpurun -- 'bash tools/run_gpu_dev.sh'`): parity tests, then the full-size bench line.
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
timeout 1500 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_dev.log 2> gpurun_out/bench_dev.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_dev.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_dev.log').read().strip().splitlines()[-1]); r=d['roofline']
print('h_ms=%.2f w_ms=%.2f step=%.2f ms value=%.3e loss=%s clocks=%s'%(r['avg_launch_ms'], r['w_pass']['avg_launch_ms'], d['ms_per_step'], d['value'], d['config']['loss_first_last'], d['clocks']))
PY
