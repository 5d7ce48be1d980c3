set -x
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
timeout 900 python tools/compare_engines.py 200000 20000 32 6 > gpurun_out/compare_engines.log 2>&1; echo "compare rc=$?"; head -3 gpurun_out/compare_engines.log
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full_tc4.log 2> gpurun_out/bench_full_tc4.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_full_tc4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_full_tc4.log').read().strip().splitlines()[-1]); r=d['roofline']
print('h_ms=%.2f w_ms=%.2f step=%.2f ms value=%.3e frac=%.3f frac_exec=%.3f loss=%s clocks=%s'%(r['avg_launch_ms'], r['w_pass']['avg_launch_ms'], d['ms_per_step'], d['value'], r['frac'], r.get('frac_executed',0), d['config']['loss_first_last'], d['clocks']))
print('e2e', d['e2e']['value'], d['e2e']['seconds'])
PY
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2_tc.log 2> gpurun_out/bench_n2_tc.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/bench_n2_tc.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2_tc.log').read().strip().splitlines()[-1]); r=d['roofline']
print('N=2: h_ms=%.2f w_ms=%.2f step=%.2f ms value=%.3e clocks=%s e2e=%s'%(r['avg_launch_ms'], r['w_pass']['avg_launch_ms'], d['ms_per_step'], d['value'], d['clocks'], d['e2e'] and (d['e2e']['value'], d['e2e']['seconds'])))
PY
