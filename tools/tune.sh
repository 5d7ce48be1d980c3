#!/bin/bash
# variant sweep for the K=32 fp32 bit-packed pass kernels (run on the GPU box)
R=${R:-200000}; C=${C:-100000}
for h in default s4c4 s2c2 s2c4n4 s4c8 s2c4n2; do
  if [ $h = default ]; then unset NBMF_TUNE_H; else export NBMF_TUNE_H=$h; fi
  python bench.py --rows $R --cols $C --steps 3 --warmup 2 --no-e2e --no-cpu 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('H=$h', 'h_ms=%.2f frac=%.3f | w_ms=%.2f frac=%.3f | step=%.2f ms value=%.3e'%(r['avg_launch_ms'], r['frac'], r['w_pass']['avg_launch_ms'], r['w_pass']['frac'], d['ms_per_step'], d['value']))"
done
unset NBMF_TUNE_H
for w in s4c4 s2c2 s2c4n4 s4c8 s4c4b3; do
  export NBMF_TUNE_W=$w
  python bench.py --rows $R --cols $C --steps 3 --warmup 2 --no-e2e --no-cpu 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('W=$w', 'h_ms=%.2f frac=%.3f | w_ms=%.2f frac=%.3f | step=%.2f ms value=%.3e'%(r['avg_launch_ms'], r['frac'], r['w_pass']['avg_launch_ms'], r['w_pass']['frac'], d['ms_per_step'], d['value']))"
done
