#!/bin/bash
# A/B on one box: bench.py with an environment switch off / on, alternating.  usage: VAR=WG_PACK4 bash tools/gpu_ab.sh
mkdir -p gpurun_out
V=${VAR:-WG_PACK4}
for rep in 1 2; do
  for val in 0 1; do
    env $V=$val python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$V=$val', 'ms_per_step %.2f' % d['ms_per_step'], 'e2e %.2f' % d['e2e']['ms_per_step'], 'launch_ms %.4f' % d['roofline']['launch_ms'], d['clocks'])"
  done
done | tee gpurun_out/ab_$V.txt
