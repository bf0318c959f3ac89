#!/usr/bin/env bash
# decode attention at batch (north-star: >= 0.70 of the HBM roofline by ncu): event-timed sweep + one full ncu capture per kernel at B = 64
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -x --timeout=180 -k "graph or batch or lock" 2>&1 | tail -2; timeout 600 python tools/attn_bench.py 8 32 64 > gpurun_out/r2c_attn_batch.txt 2>&1; cat gpurun_out/r2c_attn_batch.txt
timeout 900 ncu --set full --clock-control none -k regex:decode_attention -s 60 -c 4 -f -o gpurun_out/r2c_decode_attention_b64 python tools/attn_bench.py 64 > gpurun_out/r2c_ncu_attn.log 2>&1; echo "ncu rc=$?"
