#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --channels 512 --batch 8 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_wide.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 800 -c 230 --csv --log-file gpurun_out/launches_wide.csv $CMD > gpurun_out/ncu_launches_wide.log 2>&1
echo "exit $?"; tail -2 gpurun_out/plain_wide.log | cut -c1-200
