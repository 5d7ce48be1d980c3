# round 2, GPU call 6 (2 GPUs): multi-GPU tests, N=2 bench with e2e, config-5 restarts over 2 GPUs
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/r2_pytest6.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2_pytest6.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench6_n2.log 2> gpurun_out/r2_bench6_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/r2_bench6_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench6_n2.log').read().strip().splitlines()[-1]); r=d['roofline']
print('N=2 h_ms=%.2f w_ms=%.2f step=%.2f ms value=%.3e loss=%s'%(r['avg_launch_ms'], r['w_pass']['avg_launch_ms'], d['ms_per_step'], d['value'], d['config']['loss_first_last']))
print('parity', {k:d['parity_check'][k] for k in ('ok','h_rel','w_rel','loss_rel')})
print('e2e', {k:d['e2e'][k] for k in ('value','seconds','cold_seconds','final_loss')})
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --configs --gpus 2 > gpurun_out/r2_cfg5_n2.log 2> gpurun_out/r2_cfg5_n2.err; echo "cfg5 n2 rc=$?"; cat gpurun_out/r2_cfg5_n2.log | cut -c1-260; tail -3 gpurun_out/r2_cfg5_n2.err
timeout 300 python bench.py --configs --restarts-only > gpurun_out/r2_cfg5_n1.log 2>&1; echo "cfg5 n1 rc=$?"; cat gpurun_out/r2_cfg5_n1.log | cut -c1-260
