# round 2, GPU call 32 (8 GPUs): N=8 bench as the driver runs it, config 5 restarts over 8 GPUs (final build)
N=8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench32_n$N.log 2> gpurun_out/r2_bench32_n$N.err; echo "bench n$N rc=$?"; tail -2 gpurun_out/r2_bench32_n$N.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench32_n$N.log').read().strip().splitlines()[-1]); r=d['roofline']
print('N=$N h_ms=%.2f w_ms=%.2f step=%.2f ms value=%.3e loss=%s clocks=%s'%(r['avg_launch_ms'], r['w_pass']['avg_launch_ms'], d['ms_per_step'], d['value'], d['config']['loss_first_last'], d['clocks']))
print('parity', {k:d['parity_check'][k] for k in ('ok','h_rel','w_rel','loss_rel')})
print('e2e', {k:d['e2e'][k] for k in ('value','seconds','cold_seconds','final_loss')})
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --configs --gpus 8 > gpurun_out/r2_cfg5_n8b.log 2> gpurun_out/r2_cfg5_n8b.err; echo "cfg5 n8 rc=$?"; grep '^{' gpurun_out/r2_cfg5_n8b.log | cut -c1-200; tail -2 gpurun_out/r2_cfg5_n8b.err | cut -c1-200
