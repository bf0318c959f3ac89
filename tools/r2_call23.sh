#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -x --timeout=180 -k "not dit" > gpurun_out/c23_pytest.log 2>&1
rc=$?; echo "pytest rc=$rc"; tail -5 gpurun_out/c23_pytest.log
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --oracle-check sample --mega 2 --dit 0 --windows 40 > gpurun_out/c23_bench.json 2> gpurun_out/c23_bench.err
tail -3 gpurun_out/c23_bench.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/c23_bench.json") if l.startswith("{")][-1])
    print("value", d.get("value"), "e2e", d.get("e2e", {}).get("value"), "us/token", d.get("roofline", {}).get("us_per_token"), "sha", d.get("token_stream_sha1"), d.get("stages_ms_per_song"), d.get("error", "")[:300])
except Exception as e:
    print("no json", e)
PY
for cta in 0 1; do
  MB200_TRACE_CTA=$cta timeout 200 python tools/mega3_trace.py > gpurun_out/c23_trace_cta$cta.txt 2>&1; tail -15 gpurun_out/c23_trace_cta$cta.txt
done
