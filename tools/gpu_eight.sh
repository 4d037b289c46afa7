#!/bin/bash
# N-GPU weak-scaling line (one process per GPU, NCCL all-gather of the audio): usage N=8 bash tools/gpu_eight.sh
mkdir -p gpurun_out
N=${N:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu.log 2>&1
echo "$N-GPU bench exit $?"; grep '^{' gpurun_out/bench_${N}gpu.log | cut -c1-330; tail -3 gpurun_out/bench_${N}gpu.log | cut -c1-200
