mkdir -p gpurun_out
VAR=WG_TRIM bash tools/gpu_ab.sh
for g in 5 7 8; do WG_G2AT=$g python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('WG_G2AT=$g ms_per_step %.2f' % d['ms_per_step'], d['clocks'])"; done | tee gpurun_out/g2at_sweep.txt
