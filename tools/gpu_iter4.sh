#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-12} gpurun_out/$name.log; }
TAILN=6 run t_all python -m pytest tests -q -m gpu -x --timeout 600
TAILN=3 run bench_cont python bench.py --steps 5 --warmup 3 --no-cpu-baseline
TAILN=12 run timeline_cont python tools/timeline.py
TAILN=8 run configs python tools/measure_configs.py
