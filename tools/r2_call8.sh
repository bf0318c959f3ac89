#!/usr/bin/env bash
# round-2 GPU call 8: fine timeline of the K-split dataflow kernel + ablations (no weight stream / no waiting / no math)
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for dbg in 0 1 2 3 4 7; do
  echo "=== ll_debug=$dbg"
  MB200_LL_DEBUG=$dbg timeout 200 python tools/mega3_trace.py > gpurun_out/c8_trace_dbg$dbg.txt 2>&1; head -14 gpurun_out/c8_trace_dbg$dbg.txt
done
for reps in 1 2 4 16; do
  echo "=== reps=$reps"
  MB200_LL_REPS=$reps timeout 200 python tools/mega3_trace.py > gpurun_out/c8_trace_reps$reps.txt 2>&1; head -14 gpurun_out/c8_trace_reps$reps.txt
done
