#!/bin/bash
# quick iteration check: parity tests (fp16 path + gemm), bench, timeline
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-12} gpurun_out/$name.log; }
run t_fp16 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "not fp32" --timeout 400
TAILN=3 run bench_full python bench.py --steps 5 --warmup 3 --no-cpu-baseline
TAILN=30 run timeline python tools/timeline.py
