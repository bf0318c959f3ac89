set -x
timeout 300 python tools/mega_trace.py 2>&1 | tail -13
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -4
