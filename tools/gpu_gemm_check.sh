#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fulldims.py -m gpu -q -x --timeout=300 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -x --timeout=300 -k "dit or encoder or forward" 2>&1 | tail -2
timeout 300 python tools/dit_bench.py 1 2>&1 | tail -1
timeout 300 python tools/stage_times.py 2>&1 | head -3
