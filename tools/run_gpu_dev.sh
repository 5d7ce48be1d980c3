timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
timeout 900 python tools/dense_bench.py > gpurun_out/dense_bench.log 2> gpurun_out/dense_bench.err; echo rc=$?; tail -3 gpurun_out/dense_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/dense_bench.log'):
    d=json.loads(l); r=d['roofline']
    print(d['config']['workload'], '| value %.3e | W pass %.2f ms %.0f GB/s (%.2f) | H pass %.2f ms %.0f GB/s (%.2f) | loss %s'%(d['value'], r['avg_launch_ms'], r['achieved'], r['frac'], r['h_pass']['avg_launch_ms'], r['h_pass']['achieved'], r['h_pass']['frac'], d['loss_first_last']))
PY
