mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --timeout 600 -s -k "wide" 2>&1 | grep -E "max-abs|passed|failed|Error|assert" | cut -c1-200 | tail -20
timeout 900 python -m pytest tests -q -m gpu -x --timeout 600 2>&1 | tail -3
for pk in 0 1 0 1; do WG_PACK4=$pk python bench.py --steps 3 --warmup 3 --channels 512 --batch 32 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('WG_PACK4=$pk config3 ms_per_step %.1f  x_rt %.0f' % (d['ms_per_step'], d['rtf_x_realtime']), d['clocks'])"; done | tee gpurun_out/ab_pack4_wide.txt
