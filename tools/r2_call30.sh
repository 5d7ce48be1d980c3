# round 2, GPU call 30: final build -- default bench, reference arm, smoke, ncu counters + full-set capture + launch list
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench30.log 2> gpurun_out/r2_bench30.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench30.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench30.log').read().strip().splitlines()[-1]); r=d['roofline']
print('h_ms=%.2f w_ms=%.2f step=%.2f ms value=%.3e frac=%.3f loss=%s clocks=%s'%(r['avg_launch_ms'], r['w_pass']['avg_launch_ms'], d['ms_per_step'], d['value'], r['frac'], d['config']['loss_first_last'], d['clocks']))
print('parity', {k:d['parity_check'][k] for k in ('ok','h_rel','w_rel','loss_rel')})
print('e2e', {k:d['e2e'][k] for k in ('value','seconds','cold_seconds')})
print('cpu', d['cpu_baseline']['value'], 'launches', d['gpu_launches'])
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench30_ref.log 2>&1; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2_bench30_ref.log | tail -1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke30.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/r2_smoke30.log
timeout 900 python tools/ncu_regen.py --tag r02 > gpurun_out/r2_ncu_regen30.log 2>&1; echo "regen rc=$?"; tail -3 gpurun_out/r2_ncu_regen30.log
cp profiles/dram_traffic.json gpurun_out/dram_traffic.json; cp profiles/r02_ncu_pass_kernel_counters_*.csv gpurun_out/ 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pass_tc_kernel -c 3 -o gpurun_out/r2_prof_final2 -f python bench.py --rows 65536 --cols 32768 --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/r2_ncu30.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches30.csv python bench.py --rows 65536 --cols 32768 --steps 2 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/r2_ncu30b.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:fused_fit -c 2 -o gpurun_out/r2_fused_final_lastfm_f64 -f python tools/fused_one.py 1226 285 10 float64 300 > gpurun_out/ncu_fused30.log 2>&1; echo "ncu fused rc=$?"
