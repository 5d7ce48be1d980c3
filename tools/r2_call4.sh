# round 2, GPU call 4: K <= 32 two-pipeline kernels (+ KB = 16, column flip) and K <= 64 three-pipeline kernels; whole suite, bench, K sweep
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2_pytest4.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2_pytest4.log
timeout 600 python tools/tc_ab.py 65536 32768 32,12,64 > gpurun_out/r2_tc_ab4.log 2>&1; echo "ab rc=$?"; tail -6 gpurun_out/r2_tc_ab4.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2_bench4.log 2> gpurun_out/r2_bench4.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench4.err
python - <<'PY'
import json
for f in ('gpurun_out/r2_bench4.log',):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f,'h_ms=%.2f w_ms=%.2f step=%.2f ms value=%.3e loss=%s clocks=%s parity=%s'%(r['avg_launch_ms'], r['w_pass']['avg_launch_ms'], d['ms_per_step'], d['value'], d['config']['loss_first_last'], d['clocks'], {k:d['parity_check'][k] for k in ('ok','h_rel','w_rel','loss_rel')}))
    except Exception as e:
        print(f, 'failed', e)
PY
for K in 8 16 32 64; do for E in tensor simt; do timeout 300 python bench.py --rows 200000 --cols 100000 --k $K --engine $E --steps 3 --warmup 3 --no-e2e --no-cpu --no-parity 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('K=$K engine=$E  %.3e updates/s  %.2f ms/step  H %.2f ms  W %.2f ms' % (d['value'], d['ms_per_step'], r['avg_launch_ms'], r['w_pass']['avg_launch_ms']))"; done; done > gpurun_out/r2_k_sweep.log 2>&1; cat gpurun_out/r2_k_sweep.log
timeout 600 python bench.py --configs > gpurun_out/r2_configs.log 2>&1; echo "configs rc=$?"; tail -12 gpurun_out/r2_configs.log
