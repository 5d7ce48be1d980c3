# round 2, GPU call 3: full GPU suite on the round-2 kernels (shared Theta slots), A/B vs round 1, bench, ncu captures
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2_pytest3.log
timeout 600 python tools/tc_ab.py 65536 32768 32,12,64 > gpurun_out/r2_tc_ab3.log 2>&1; echo "ab rc=$?"; tail -6 gpurun_out/r2_tc_ab3.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2_bench3.log 2> gpurun_out/r2_bench3.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench3.err
python - <<'PY'
import json
for f in ('gpurun_out/r2_bench3.log',):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f,'h_ms=%.2f w_ms=%.2f step=%.2f ms value=%.3e loss=%s clocks=%s parity=%s'%(r['avg_launch_ms'], r['w_pass']['avg_launch_ms'], d['ms_per_step'], d['value'], d['config']['loss_first_last'], d['clocks'], {k:d['parity_check'][k] for k in ('ok','h_rel','w_rel','loss_rel')}))
    except Exception as e:
        print(f, 'failed', e)
PY
# ncu: full-set capture of the two pass kernels (1 launch each) at 65536 x 32768, K = 32
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pass_tc_kernel -c 2 -o gpurun_out/r2_prof_tc -f python bench.py --rows 65536 --cols 32768 --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/r2_ncu3.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r2_ncu3.log
