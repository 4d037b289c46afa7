#!/bin/bash
# end-of-iteration evidence: GPU suite, bench (with CPU baseline), reference arm, ncu launch list + full sets
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc"; tail -n ${TAILN:-12} gpurun_out/$name.log | cut -c1-400; return $rc; }
TAILN=4 run t_all python -m pytest tests -q -m gpu -x --timeout 600
TAILN=3 run smoke python -c "import __graft_entry__ as g; g.smoke()"
TAILN=2 run bench_final python bench.py --steps 10 --warmup 3
TAILN=2 run bench_ref python bench.py --impl reference --steps 2 --warmup 1
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 260 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:wn_layer_pair -s 120 -c 2 -o gpurun_out/prof_layer -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full set exit $?"
