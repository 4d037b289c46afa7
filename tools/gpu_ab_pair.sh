#!/bin/bash
mkdir -p gpurun_out
for mode in 0 1 0 1; do
  echo "=== WG_PAIR=$mode"
  WG_PAIR=$mode timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('ms_per_step %.2f  xRT %.0f  layer_ms %.4f  exec_TF %.0f  clk %s %s' % (d['ms_per_step'], d['rtf_x_realtime'], r['launch_ms'], r['executed_tensor_tflops'], d['clocks']['sm_mhz'], d['clocks']['reasons']))"
done
echo "=== timeline WG_PAIR=1"
WG_PAIR=1 timeout 200 python tools/timeline.py 2>&1 | tail -22
