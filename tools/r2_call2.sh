# round 2, GPU call 2: round-2 tensor kernels: parity tests, A/B against the round-1 kernels, full-size bench of both
timeout 600 python -m pytest tests/test_gpu_tensor_engine.py tests/test_gpu_depth.py tests/test_gpu_onestep.py tests/test_gpu_large.py -q -m gpu -x > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_pytest2.log
timeout 600 python tools/tc_ab.py 65536 32768 32,20,12,48,64 > gpurun_out/r2_tc_ab.log 2>&1; echo "ab rc=$?"; cat gpurun_out/r2_tc_ab.log | tail -12
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2_bench2.log 2> gpurun_out/r2_bench2.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench2.err
NBMF_TC_V1=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/r2_bench2_v1.log 2> gpurun_out/r2_bench2_v1.err; echo "bench v1 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2_bench2.log','gpurun_out/r2_bench2_v1.log'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f,'h_ms=%.2f w_ms=%.2f step=%.2f ms value=%.3e loss=%s clocks=%s parity=%s'%(r['avg_launch_ms'], r['w_pass']['avg_launch_ms'], d['ms_per_step'], d['value'], d['config']['loss_first_last'], d['clocks'], d.get('parity_check')))
    except Exception as e:
        print(f, 'failed', e)
PY
