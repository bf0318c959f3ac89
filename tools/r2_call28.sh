#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout=180 -k "attention" > gpurun_out/c28_pytest_attn.log 2>&1; echo "attention pytest rc=$?"; tail -5 gpurun_out/c28_pytest_attn.log
timeout 300 python tools/attn_tc_bench.py > gpurun_out/c28_attn_bench.txt 2>&1; cat gpurun_out/c28_attn_bench.txt
