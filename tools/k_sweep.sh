#!/bin/bash
# SIMT (packed FFMA2) vs tensor (tcgen05 split precision) engine by K: the crossover BASELINE.json configs[4] asks about,
# on a problem large enough to fill the GPU (run on the GPU box; one line per (K, engine)).
R=${R:-200000}; C=${C:-100000}
for k in 8 16 24 32 64; do
  for e in simt tensor; do
    if [ $k -gt 32 ] && [ $e = tensor ]; then continue; fi
    python bench.py --rows $R --cols $C --k $k --engine $e --steps 3 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('K=%2d engine=%-6s h_ms=%7.2f w_ms=%7.2f step=%7.2f ms  %.3e updates/s  loss %s'%($k, r['engine'], r['avg_launch_ms'], r['w_pass']['avg_launch_ms'], d['ms_per_step'], d['value'], d['config']['loss_first_last']))"
  done
done
