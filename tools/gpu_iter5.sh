#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-12} gpurun_out/$name.log | cut -c1-300; }
TAILN=6 run t_all python -m pytest tests -q -m gpu -x --timeout 600
TAILN=2 run bench_final python bench.py --steps 5 --warmup 3
WG_G2AT=11 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('g2at 11 ms_per_step', d['ms_per_step'], d['clocks'])"
WG_G2AT=7 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('g2at 7 ms_per_step', d['ms_per_step'], d['clocks'])"
TAILN=8 run configs python tools/measure_configs.py
bash tools/gpu_ncu.sh 2>&1 | tail -12 | cut -c1-200
