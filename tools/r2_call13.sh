# round 2, GPU call 13: final build -- whole suite, default bench, ncu counters + full-set capture for this build
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2_pytest13.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest13.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench13.log 2> gpurun_out/r2_bench13.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench13.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench13.log').read().strip().splitlines()[-1]); r=d['roofline']
print('h_ms=%.2f w_ms=%.2f step=%.2f ms value=%.3e frac=%.3f loss=%s clocks=%s'%(r['avg_launch_ms'], r['w_pass']['avg_launch_ms'], d['ms_per_step'], d['value'], r['frac'], d['config']['loss_first_last'], d['clocks']))
print('parity', {k:d['parity_check'][k] for k in ('ok','h_rel','w_rel','loss_rel')})
print('e2e', {k:d['e2e'][k] for k in ('value','seconds','cold_seconds')})
print('issue', r.get('issue_slots'))
PY
timeout 900 python tools/ncu_regen.py --tag r02 > gpurun_out/r2_ncu_regen13.log 2>&1; echo "regen rc=$?"; tail -3 gpurun_out/r2_ncu_regen13.log
cp profiles/dram_traffic.json gpurun_out/dram_traffic.json; cp profiles/r02_ncu_pass_kernel_counters_*.csv gpurun_out/ 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pass_tc_kernel -c 3 -o gpurun_out/r2_prof_final -f python bench.py --rows 65536 --cols 32768 --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/r2_ncu13.log 2>&1; echo "ncu rc=$?"
timeout 600 python smoke_run.py 2>/dev/null; python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke13.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r2_smoke13.log
