# round 2, GPU call 1: whole GPU suite (new depth / cancellation / ref-suite / binding tests), default bench with the parity check, MMA N=16 microbench
timeout 1700 python -m pytest tests -q -m gpu > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2_pytest1.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench1.log 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench1.log').read().strip().splitlines()[-1]); r=d['roofline']
print('h_ms=%.2f w_ms=%.2f step=%.2f ms value=%.3e e2e=%.3e loss=%s clocks=%s'%(r['avg_launch_ms'], r['w_pass']['avg_launch_ms'], d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['loss_first_last'], d['clocks']))
print('parity', d['parity_check'])
print('cpu', d['cpu_baseline'])
PY
timeout 120 tools/bin/tc_bench > gpurun_out/r2_tc_bench.log 2>&1; echo "tc_bench rc=$?"; cat gpurun_out/r2_tc_bench.log | head -60
