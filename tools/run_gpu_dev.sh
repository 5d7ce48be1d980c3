timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1_final.log 2> gpurun_out/bench_n1_final.err; echo "bench n1 rc=$?"
for n in 2 4 8; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n${n}_final.log 2> gpurun_out/bench_n${n}_final.err; echo "bench n$n rc=$?"; tail -1 gpurun_out/bench_n${n}_final.err | cut -c1-200
done
python - <<'PY'
import json
for n in (1,2,4,8):
    d=json.loads(open(f'gpurun_out/bench_n{n}_final.log').read().strip().splitlines()[-1]); r=d['roofline']
    print('N=%d: h_ms=%.2f w_ms=%.2f step=%.2f ms value=%.3e frac=%.3f clocks=%s e2e=%s'%(n, r['avg_launch_ms'], r['w_pass']['avg_launch_ms'], d['ms_per_step'], d['value'], r['frac'], d['clocks'], d['e2e'] and (d['e2e']['value'], d['e2e']['seconds'])))
PY
