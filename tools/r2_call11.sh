#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for dbg in 0 1 5; do
  echo "=== ll_debug=$dbg"
  MB200_LL_DEBUG=$dbg timeout 200 python tools/mega3_trace.py > gpurun_out/c11_trace_dbg$dbg.txt 2>&1; head -13 gpurun_out/c11_trace_dbg$dbg.txt
done
