mkdir -p gpurun_out
for rep in 1 2 3; do for g in 6 8 10; do WG_G2AT=$g python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('WG_G2AT=$g ms_per_step %.2f' % d['ms_per_step'], d['clocks']['sm_mhz'])"; done; done | tee gpurun_out/g2at_sweep_long.txt
for rep in 1 2 3; do for t in 0 1; do WG_TRIM=$t python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('WG_TRIM=$t ms_per_step %.2f' % d['ms_per_step'], d['clocks']['sm_mhz'])"; done; done | tee gpurun_out/trim_long.txt
