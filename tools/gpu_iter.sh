#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-12} gpurun_out/$name.log | cut -c1-300; }
TAILN=6 run t_all python -m pytest tests -q -m gpu -x --timeout 600
TAILN=2 run bench_final python bench.py --steps 5 --warmup 3
TAILN=2 run bench_final2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
TAILN=8 run configs python tools/measure_configs.py
TAILN=24 run timeline_v8 python tools/timeline.py
python tools/tile_starts.py 3 2>&1 | tail -3
