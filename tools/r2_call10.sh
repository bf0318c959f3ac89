#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -x --timeout=180 -k "(greedy or sampl) and not graph" > gpurun_out/c10_pytest_ks.log 2>&1
rc=$?; echo "ksplit pytest rc=$rc"; tail -8 gpurun_out/c10_pytest_ks.log
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --oracle-check sample --mega 2 --dit 0 --windows 40 > gpurun_out/c10_bench_mega3.json 2> gpurun_out/c10_bench_mega3.err
tail -3 gpurun_out/c10_bench_mega3.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/c10_bench_mega3.json") if l.startswith("{")][-1])
    print("mega 3: value", d.get("value"), "e2e", d.get("e2e", {}).get("value"), "us/token", d.get("roofline", {}).get("us_per_token"), "sha", d.get("token_stream_sha1"), d.get("stages_ms_per_song"), d.get("error", "")[:300])
except Exception as e:
    print("mega 3: no json", e)
PY
for dbg in 0 4; do
  echo "=== ll_debug=$dbg"
  MB200_LL_DEBUG=$dbg timeout 200 python tools/mega3_trace.py > gpurun_out/c10_trace_dbg$dbg.txt 2>&1; head -26 gpurun_out/c10_trace_dbg$dbg.txt
done
