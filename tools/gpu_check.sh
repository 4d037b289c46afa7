#!/bin/bash
# Stage-isolated GPU check: each stage under its own timeout so a hang does not hide the others.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name" ; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n 25 gpurun_out/$name.log; }
run t_gemm python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "descriptor_gemm" --timeout 300
run t_fp32 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "fp32" --timeout 500
run t_fp16_stage python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "stagewise and fp16" --timeout 300
run t_fp16_rest python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "not fp32 and not stagewise and not descriptor_gemm" --timeout 500
run bench_small python bench.py --batch 2 --frames 200 --steps 2 --warmup 3 --no-cpu-baseline
run bench_full python bench.py --steps 3 --warmup 3
