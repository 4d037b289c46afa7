#!/bin/bash
# A/B of the continuous-issue pair kernel (default) against the older pair kernel (WG_PAIR=1): parity, bench, timeline.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-12} gpurun_out/$name.log; }
TAILN=40 run t_stage python -m pytest tests/test_gpu_parity.py -q -m gpu -s -x -k "stagewise and fp16" --timeout 300
TAILN=6 run t_all python -m pytest tests -q -m gpu -x --timeout 600
TAILN=3 run bench_cont python bench.py --steps 5 --warmup 3 --no-cpu-baseline
TAILN=30 run timeline_cont python tools/timeline.py
export WG_PAIR=1
TAILN=3 run bench_pair1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline
unset WG_PAIR
TAILN=8 run configs python tools/measure_configs.py
