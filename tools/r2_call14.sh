#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for cta in 0 1 13; do
  MB200_TRACE_CTA=$cta timeout 200 python tools/mega3_trace.py > gpurun_out/c14_trace_cta$cta.txt 2>&1; tail -16 gpurun_out/c14_trace_cta$cta.txt
done
