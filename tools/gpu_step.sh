#!/bin/bash
# One GPU session on the B200 box: named steps, each under its own timeout, logs under gpurun_out/<tag>/.
# usage: bash tools/gpu_step.sh <tag> <step> [<step> ...]       steps: see the case statement
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() { name=$1; lim=$2; shift 2; echo "=== $name"; timeout $lim "$@" > $OUT/$name.log 2>&1; rc=$?; echo "exit $rc"; tail -n ${TAILN:-6} $OUT/$name.log | cut -c1-600; return $rc; }
for step in "$@"; do
  case $step in
    quick256)  TAILN=15 run quick256 300 python -m pytest tests/test_gpu_parity.py -q -x -s -k "gemm or (stagewise and 256) or golden and not 512" ;;
    big)       TAILN=30 run big 600 python -m pytest tests/test_gpu_baseline_shapes.py -q -s -k "beyond_4_gib" --timeout 400 ;;
    energy)    python tools/energy_ab.py 2>&1 | tee $OUT/energy_ab.txt ;;
    energy_chain) python tools/energy_ab.py --only 256 --match "production,one launch,diagnostic" 2>&1 | tee $OUT/energy_chain.txt ;;
    energy_st) python tools/energy_ab.py --only 256 --match "production,x' stores" 2>&1 | tee $OUT/energy_stores.txt ;;
    energy_b)  for b in 2 4 8 16 32; do echo "batch $b"; python tools/energy_ab.py --child 256 $b --seconds 4 2>/dev/null | tail -1; done | tee $OUT/energy_batch.txt ;;
    energy_nx) WG_CHAIN=0 WG_NX=2 python bench.py --batch 2 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('parity with 2 x buffers:', d['parity']['max_abs'], d['parity']['pass'])"
               for b in 2 16; do for nx in 8 2; do echo "batch $b, WG_CHAIN=0, x buffers $nx"; WG_CHAIN=0 WG_NX=$nx python tools/energy_ab.py --child 256 $b --seconds 4 2>/dev/null | tail -1; done; done | tee $OUT/energy_nx.txt ;;
    energy_lib) for ch in "256 16" "512 8"; do for rep in 1 2; do for v in "WG_B200_LIB=$PWD/waveglow_b200/libwaveglow_b200_prev.so" "WG_CHAIN=1"; do echo "channels / batch $ch, ${v##*/}"; env $v python tools/energy_ab.py --child $ch --seconds 4 2>/dev/null | tail -1; done; done; done | tee $OUT/energy_lib.txt ;;
    energy256) python tools/energy_ab.py --only 256 2>&1 | tee $OUT/energy_ab.txt ;;
    hparams)   TAILN=30 run hparams 600 python -m pytest tests/test_gpu_parity.py -q -s -k "non_default_hparams" --timeout 300 ;;
    quick512)  TAILN=30 run quick512 300 python -m pytest tests/test_gpu_parity.py -q -x -s -k "(stagewise and 512) or 512_channel" ;;
    tests)     TAILN=8 run tests 1500 python -m pytest tests -q -m gpu -x --timeout 600 --durations=12 ;;
    tests_all) TAILN=25 run tests_all 1500 python -m pytest tests -q -m gpu --timeout 600 --durations=12 ;;
    shapes)    TAILN=30 run shapes 1200 python -m pytest tests/test_gpu_baseline_shapes.py -q -s --timeout 600 --durations=12 ;;
    smoke)     TAILN=4 run smoke 300 python -c "import __graft_entry__ as g; g.smoke()" ;;
    flags)     for f in 0 1 4 8 16 128 29 0; do echo "WG_DBG_FLAGS=$f (256 ch, B16 x 861)"; WG_DBG_FLAGS=$f python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-parity 2>/dev/null | python tools/bench_brief.py; done | tee $OUT/flags256.txt
               for f in 0 1 4 8 16 29 0; do echo "WG_DBG_FLAGS=$f (512 ch, B8 x 861)"; WG_DBG_FLAGS=$f python bench.py --channels 512 --batch 8 --steps 4 --warmup 3 --no-cpu-baseline --no-parity 2>/dev/null | python tools/bench_brief.py; done | tee $OUT/flags512.txt ;;
    chain)     TAILN=30 run chain 600 python -m pytest tests/test_gpu_parity.py -q -x -s --timeout 300 -k "chained or (stagewise and fp16) or golden" ;;
    chain2)    TAILN=30 run chain2 600 python -m pytest tests/test_gpu_parity.py -q -x -s --timeout 300 -k "chained" ;;
    ab_chain)  for v in 0 1 0 1; do echo "WG_CHAIN=$v (256 ch, B16 x 861)"; WG_CHAIN=$v python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-parity 2>/dev/null | python tools/bench_brief.py; done | tee $OUT/ab_chain256.txt
               for v in 0 1 0 1; do echo "WG_CHAIN=$v (512 ch, B8 x 861)"; WG_CHAIN=$v python bench.py --channels 512 --batch 8 --steps 5 --warmup 3 --no-cpu-baseline --no-parity 2>/dev/null | python tools/bench_brief.py; done | tee $OUT/ab_chain512.txt
               for v in 0 1; do echo "WG_CHAIN=$v"; WG_CHAIN=$v python tools/measure_configs.py 2>/dev/null; done | tee $OUT/ab_chain_configs.txt ;;
    fwd)       for env in "WG_CHAIN=1" "WG_CHAIN=0"; do env $env python tools/forward_times.py 2>&1 | tail -3; env $env python tools/forward_times.py --after-big 2>&1 | tail -3; done | tee $OUT/forward_times.txt ;;
    ab_lib)    for rep in 1 2 3; do for v in "WG_B200_LIB=$PWD/waveglow_b200/libwaveglow_b200_prev.so" "WG_CHAIN=1"; do echo "$v"; env $v python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-parity 2>/dev/null | python tools/bench_brief.py; done; done | tee $OUT/ab_lib.txt ;;
    scale*)    N=${step#scale}
               TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
               TAILN=1 run weak_n$N 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-parity
               TAILN=1 run strong512_n$N 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --scaling strong --channels 512 --global-batch 32 --no-parity ;;
    sweep_g2)  for v in 7 5 3 2 1 7; do echo "WG_G2AT0=$v (256 ch, B16 x 861)"; WG_G2AT0=$v python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-parity 2>/dev/null | python tools/bench_brief.py; done | tee $OUT/sweep_g2at0_256.txt
               for v in "10 24" "6 20" "4 14" "8 18" "12 26" "10 24"; do set -- $v; echo "WG_G2AT=$1 WG_G2AT1=$2 (512 ch, B8 x 861)"; WG_G2AT=$1 WG_G2AT1=$2 python bench.py --channels 512 --batch 8 --steps 5 --warmup 3 --no-cpu-baseline --no-parity 2>/dev/null | python tools/bench_brief.py; done | tee $OUT/sweep_g2at_512.txt ;;
    ab_chain256) for v in 0 1 0 1 0 1; do echo "WG_CHAIN=$v (256 ch, B16 x 861)"; WG_CHAIN=$v python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-parity 2>/dev/null | python tools/bench_brief.py; done | tee $OUT/ab_chain256.txt ;;
    ab_carry)  for rep in 1 2 3; do for v in "WG_EXACT_CARRY=0" "WG_EXACT_CARRY=1"; do echo "$v (512 ch, B8 x 861)"; env $v python bench.py --channels 512 --batch 8 --steps 5 --warmup 3 --no-cpu-baseline --no-parity 2>/dev/null | python tools/bench_brief.py; done; done | tee $OUT/ab_carry.txt ;;
    bench)     TAILN=1 run bench 600 python bench.py --steps 10 --warmup 3 ;;
    bench_fast) TAILN=1 run bench_fast 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity ;;
    bench512)  TAILN=1 run bench512 600 python bench.py --channels 512 --scaling strong --global-batch 32 --steps 5 --warmup 3 --no-cpu-baseline ;;
    bench512_b8) TAILN=1 run bench512_b8 600 python bench.py --channels 512 --batch 8 --steps 5 --warmup 3 --no-cpu-baseline --no-parity ;;
    ref)       TAILN=1 run ref 900 python bench.py --impl reference --steps 2 --warmup 1 ;;
    callers)   TAILN=12 run callers 900 python tools/profile_callers.py ;;
    configs)   TAILN=8 run configs 600 python tools/measure_configs.py ;;
    timeline256) TAILN=40 run timeline256 300 env WG_CHAIN=0 python tools/timeline.py ;;
    timeline_first) TAILN=40 run timeline_first 300 env WG_CHAIN=0 python tools/timeline.py --layer 0 ;;
    timeline512) TAILN=40 run timeline512 300 env WG_CHAIN=0 python tools/timeline.py --channels 512 --batch 8 ;;
    layers)    TAILN=12 run layers 300 python tools/layer_times.py ;;
    ncu_list)  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity > $OUT/ncu_launches.log 2>&1; echo "launch list exit $?" ;;
    ncu_full)  timeout 900 ncu --set full --clock-control none --import-source on -k regex:wn_layer_pair -s 48 -c 3 -o $OUT/prof_layer -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity > $OUT/ncu_full.log 2>&1; echo "full set exit $?" ;;
    ncu_list512) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches512.csv python bench.py --channels 512 --batch 8 --steps 1 --warmup 3 --no-cpu-baseline --no-parity > $OUT/ncu_launches512.log 2>&1; echo "launch list 512 exit $?" ;;
    ncu_full512) timeout 900 ncu --set full --clock-control none --import-source on -k regex:wn_layer_pair -s 48 -c 3 -o $OUT/prof_layer512 -f python bench.py --channels 512 --batch 8 --steps 1 --warmup 3 --no-cpu-baseline --no-parity > $OUT/ncu_full512.log 2>&1; echo "full set 512 exit $?" ;;
    *) echo "unknown step $step" ;;
  esac
done
