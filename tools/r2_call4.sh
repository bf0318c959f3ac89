#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_slider.py -m gpu -q --timeout=200 > gpurun_out/c4_pytest_slider.log 2>&1; echo "slider rc=$?"; tail -12 gpurun_out/c4_pytest_slider.log
timeout 200 python tools/mega2_trace.py > gpurun_out/c4_mega2_trace.txt 2>&1; cat gpurun_out/c4_mega2_trace.txt | tail -60
timeout 200 python tools/mega_trace.py > gpurun_out/c4_mega1_trace.txt 2>&1; tail -14 gpurun_out/c4_mega1_trace.txt | cut -c1-140
for sl in 0 100 300; do
  MB200_LL_SLEEP=$sl timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --mega 2 --dit 0 --windows 40 > gpurun_out/c4_bench_sleep$sl.json 2> gpurun_out/c4_bench_sleep$sl.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/c4_bench_sleep$sl.json") if l.startswith("{")][-1])
    print("ll_sleep $sl: us/token", d.get("roofline", {}).get("us_per_token"), "sha", d.get("token_stream_sha1"), d.get("error", "")[:300])
except Exception as e:
    print("sleep $sl: no json", e)
PY
done
