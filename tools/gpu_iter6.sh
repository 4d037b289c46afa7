mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_workflow.py -q -m gpu -x --timeout 600 -s -k "packed_weight" 2>&1 | tail -5
