#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n ${TAILN:-12} gpurun_out/$name.log | cut -c1-260; }
TAILN=6 run t_stage python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "stagewise or golden or segment" --timeout 300
python tools/tile_starts.py 3 2>&1 | tail -3
for f in 0 0; do
export WG_DBG_FLAGS=$f
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('flags $f ms_per_step', d['ms_per_step'], 'launch_ms', d['roofline']['launch_ms'], d['clocks'])"
done
WG_DBG_FLAGS=0 timeout 300 python tools/timeline.py 2>&1 | tail -22
