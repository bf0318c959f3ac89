#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
MB200_TRACE_CTA=0 timeout 200 python tools/mega3_trace.py > gpurun_out/c21_trace_cta0.txt 2>&1; tail -15 gpurun_out/c21_trace_cta0.txt
