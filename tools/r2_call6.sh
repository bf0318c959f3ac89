#!/usr/bin/env bash
# round-2 GPU call 6 (after container re-creation): where do we stand? whole GPU suite, token loop drivers side by side, phase trace
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout=300 --durations=15 ) > gpurun_out/c6_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -40 gpurun_out/c6_pytest_all.log
for m in 1 2; do
  timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --oracle-check none --mega $m --dit 0 --windows 40 > gpurun_out/c6_bench_mega$m.json 2> gpurun_out/c6_bench_mega$m.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/c6_bench_mega$m.json") if l.startswith("{")][-1])
    print("mega $m: value", d.get("value"), "e2e", d.get("e2e", {}).get("value"), "us/token", d.get("roofline", {}).get("us_per_token"), "sha", d.get("token_stream_sha1"), d.get("stages_ms_per_song"), d.get("error", "")[:300])
except Exception as e:
    print("mega $m: no json", e)
PY
done
timeout 200 python tools/mega2_trace.py > gpurun_out/c6_mega2_trace.txt 2>&1; head -50 gpurun_out/c6_mega2_trace.txt
timeout 300 python tools/stage_times.py > gpurun_out/c6_stage_times.txt 2>&1; tail -30 gpurun_out/c6_stage_times.txt
