#!/usr/bin/env bash
# whole GPU suite + the full default bench (211 windows + DiT + CPU baseline + whole-song oracle check)
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout=300 ) > gpurun_out/c24_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -6 gpurun_out/c24_pytest_all.log
timeout 1200 python bench.py > gpurun_out/c24_bench_full.json 2> gpurun_out/c24_bench_full.err; echo "bench rc=$?"; tail -c 3500 gpurun_out/c24_bench_full.json; tail -3 gpurun_out/c24_bench_full.err
