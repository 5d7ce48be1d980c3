# round 2, GPU call 31 (2 GPUs): multi-GPU tests, N=2 bench as the driver runs it, config 5 restarts over 2 GPUs
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/r2_pytest31.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest31.log
N=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench31_n$N.log 2> gpurun_out/r2_bench31_n$N.err; echo "bench n$N rc=$?"; tail -2 gpurun_out/r2_bench31_n$N.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench31_n$N.log').read().strip().splitlines()[-1]); r=d['roofline']
print('N=$N h_ms=%.2f w_ms=%.2f step=%.2f ms value=%.3e loss=%s clocks=%s'%(r['avg_launch_ms'], r['w_pass']['avg_launch_ms'], d['ms_per_step'], d['value'], d['config']['loss_first_last'], d['clocks']))
print('parity', {k:d['parity_check'][k] for k in ('ok','h_rel','w_rel','loss_rel')})
print('e2e', {k:d['e2e'][k] for k in ('value','seconds','cold_seconds','final_loss')})
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --configs --gpus 2 > gpurun_out/r2_cfg5_n2b.log 2> gpurun_out/r2_cfg5_n2b.err; echo "cfg5 n2 rc=$?"; grep '^{' gpurun_out/r2_cfg5_n2b.log | cut -c1-260; tail -2 gpurun_out/r2_cfg5_n2b.err | cut -c1-200
timeout 600 python bench.py --configs --restarts-only > gpurun_out/r2_cfg5_n1b.log 2> gpurun_out/r2_cfg5_n1b.err; echo "cfg5 n1 rc=$?"; grep '^{' gpurun_out/r2_cfg5_n1b.log | cut -c1-260
