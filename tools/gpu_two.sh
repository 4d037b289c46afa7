#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_workflow.py -q -m gpu -x --timeout 300 -s -k "mel_front_end or synthesize_wav" 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1
echo "2-GPU bench exit $?"; grep '^{' gpurun_out/bench_2gpu.log | cut -c1-330
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_2gpu_ref.log 2>&1
echo "2-GPU reference arm exit $?"; grep '^{' gpurun_out/bench_2gpu_ref.log | cut -c1-200
