mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --timeout 600 -s -k "golden or first_layer or stagewise" 2>&1 | grep -E "max-abs|max err|passed|failed|Error|assert" | cut -c1-200 | tail -60
timeout 900 python -m pytest tests -q -m gpu -x --timeout 600 2>&1 | tail -3
python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-300
