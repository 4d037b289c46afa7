#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc"; tail -n ${TAILN:-12} gpurun_out/$name.log | cut -c1-400; return $rc; }
TAILN=30 run t_pack python -m pytest tests/test_gpu_parity.py -q -m gpu -x -s --timeout 300 -k "packed or first_layer or segment_choice or medium" || exit 1
TAILN=6 run t_all python -m pytest tests -q -m gpu -x --timeout 600 || exit 1
TAILN=2 run bench_a python bench.py --steps 5 --warmup 3 --no-cpu-baseline
TAILN=2 run bench_b python bench.py --steps 10 --warmup 3 --no-cpu-baseline
TAILN=8 run configs python tools/measure_configs.py
