nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n${n}_final.log 2> gpurun_out/bench_n${n}_final.err; echo "bench n$n rc=$?"; tail -2 gpurun_out/bench_n${n}_final.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n${n}_final.log').read().strip().splitlines()[-1]); r=d['roofline']
print('N=$n: h_ms=%.2f w_ms=%.2f step=%.2f ms value=%.3e clocks=%s e2e=%s'%(r['avg_launch_ms'], r['w_pass']['avg_launch_ms'], d['ms_per_step'], d['value'], d['clocks'], d['e2e'] and (d['e2e']['value'], d['e2e']['seconds'])))
PY
done
