#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python tools/attn_tc_bench.py > gpurun_out/c26_attn_bench.txt 2>&1; cat gpurun_out/c26_attn_bench.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"attention|attn_prep" -c 40 --csv --log-file gpurun_out/c26_attn_launches.csv python tools/attn_tc_bench.py > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/c26_attn_launches.csv 2>/dev/null | head -12
timeout 300 python tools/dit_bench.py 1 > gpurun_out/c26_dit_bench.txt 2>&1; tail -2 gpurun_out/c26_dit_bench.txt
