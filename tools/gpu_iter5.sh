mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc"; tail -n ${TAILN:-12} gpurun_out/$name.log | cut -c1-400; return $rc; }
TAILN=6 run t_all python -m pytest tests -q -m gpu -x --timeout 600 || exit 1
VAR=WG_PDL bash tools/gpu_ab.sh
TAILN=8 run configs python tools/measure_configs.py
