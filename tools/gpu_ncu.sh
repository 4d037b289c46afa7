#!/bin/bash
# ncu evidence: (1) per-launch durations of one bench step, (2) full metric set of the fused WN layer kernel.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 260 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:wn_layer_pair -s 120 -c 3 -o gpurun_out/prof_layer -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full set exit $?"
tail -3 gpurun_out/plain.log; tail -5 gpurun_out/ncu_full.log; ls -la gpurun_out
