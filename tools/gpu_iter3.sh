#!/bin/bash
# whole GPU suite, bench (256 and 512 channels), ncu launch list + full set of the layer kernel (first layer and a generic one)
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc"; tail -n ${TAILN:-12} gpurun_out/$name.log | cut -c1-400; return $rc; }
TAILN=30 run t_stage python -m pytest tests/test_gpu_parity.py -q -m gpu -x -s --timeout 300 -k "stagewise and wide" || exit 1
TAILN=6 run t_all python -m pytest tests -q -m gpu -x --timeout 600 || exit 1
TAILN=2 run bench_a python bench.py --steps 5 --warmup 3
TAILN=2 run bench_c3 python bench.py --steps 3 --warmup 3 --channels 512 --batch 32 --no-cpu-baseline
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 260 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:wn_layer_pair -s 120 -c 2 -o gpurun_out/prof_layer -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full set exit $?"
ncu --set full --clock-control none -k regex:flow_step -s 14 -c 1 -o gpurun_out/prof_flow_step -f $CMD > gpurun_out/ncu_full_fs.log 2>&1
echo "flow step full set exit $?"
