timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/stage_times.py 2>&1 | tail -5
timeout 300 python tools/dit_bench.py 1 2>&1 | tail -1
