# round 2, GPU call 5: whole suite, pass times, full bench with e2e (cold + warm) and the CPU leg
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2_pytest5.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2_pytest5.log
timeout 600 python tools/tc_ab.py 65536 32768 32,12,64 > gpurun_out/r2_tc_ab5.log 2>&1; echo "ab rc=$?"; tail -4 gpurun_out/r2_tc_ab5.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench5.log 2> gpurun_out/r2_bench5.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench5.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench5.log').read().strip().splitlines()[-1]); r=d['roofline']
print('h_ms=%.2f w_ms=%.2f step=%.2f ms value=%.3e frac=%.3f loss=%s clocks=%s'%(r['avg_launch_ms'], r['w_pass']['avg_launch_ms'], d['ms_per_step'], d['value'], r['frac'], d['config']['loss_first_last'], d['clocks']))
print('parity', {k:d['parity_check'][k] for k in ('ok','h_rel','w_rel','loss_rel')})
print('e2e', {k:d['e2e'][k] for k in ('value','seconds','cold_seconds','h2d_bytes_per_step','d2h_bytes_per_step')})
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
PY
