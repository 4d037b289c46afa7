#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-12} gpurun_out/$name.log; }
run t_fp16 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "not fp32" --timeout 400
TAILN=3 run bench_full python bench.py --steps 5 --warmup 3 --no-cpu-baseline
TAILN=30 run timeline python tools/timeline.py
export WG_DBG_FLAGS=4
TAILN=3 run bench_noprefetch python bench.py --steps 5 --warmup 3 --no-cpu-baseline
TAILN=30 run timeline_noprefetch python tools/timeline.py
export WG_DBG_FLAGS=0 WG_GATE=f32
TAILN=3 run bench_gatef32 python bench.py --steps 5 --warmup 3 --no-cpu-baseline
