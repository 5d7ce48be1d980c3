# round 2, GPU call 10: half-buffer release of the ratio region (split) vs one release per block (nosplit); new depth cases; multifit tests
for V in nosplit split nosplit split; do NBMF_B200_LIB=$PWD/tools/bin/libnbmf_$V.so timeout 300 python tools/tc_ab.py 65536 32768 32 2>&1 | tail -1 | sed "s/^/$V: /"; done
timeout 900 python -m pytest tests/test_gpu_depth.py tests/test_gpu_multifit.py tests/test_gpu_tensor_engine.py tests/test_gpu_onestep.py -q -m gpu > gpurun_out/r2_pytest10.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest10.log
for V in nosplit split; do NBMF_B200_LIB=$PWD/tools/bin/libnbmf_$V.so timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2_bench10_$V.log 2> gpurun_out/r2_bench10_$V.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench10_$V.log').read().strip().splitlines()[-1]); r=d['roofline']
print('$V h_ms=%.2f w_ms=%.2f step=%.2f ms value=%.3e clocks=%s parity=%s'%(r['avg_launch_ms'], r['w_pass']['avg_launch_ms'], d['ms_per_step'], d['value'], d['clocks']['sm_mhz'], d['parity_check']['ok']))
PY
done
