#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 3 -c 1 -f -o gpurun_out/c27_attn_tc python tools/attn_tc_bench.py > gpurun_out/c27_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/c27_ncu.log
