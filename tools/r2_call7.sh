#!/usr/bin/env bash
# round-2 GPU call 7: K-split GEMV phases of the dataflow megakernel — parity, then speed + phase trace
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -x --timeout=180 -k "greedy and dataflow_ksplit" > gpurun_out/c7_pytest_ks.log 2>&1
rc=$?; echo "ksplit pytest rc=$rc"; tail -8 gpurun_out/c7_pytest_ks.log
timeout 600 python -m pytest tests/test_gpu_fulldims.py -m gpu -q --timeout=300 > gpurun_out/c7_pytest_full.log 2>&1; echo "fulldims rc=$?"; tail -5 gpurun_out/c7_pytest_full.log
for m in 3; do
  timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --oracle-check sample --mega $m --dit 0 --windows 40 > gpurun_out/c7_bench_mega$m.json 2> gpurun_out/c7_bench_mega$m.err
  tail -3 gpurun_out/c7_bench_mega$m.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/c7_bench_mega$m.json") if l.startswith("{")][-1])
    print("mega $m: value", d.get("value"), "e2e", d.get("e2e", {}).get("value"), "us/token", d.get("roofline", {}).get("us_per_token"), "sha", d.get("token_stream_sha1"), d.get("stages_ms_per_song"), d.get("oracle_check"), d.get("error", "")[:300])
except Exception as e:
    print("mega $m: no json", e)
PY
done
MB200_MEGA=3 timeout 200 python tools/mega2_trace.py > gpurun_out/c7_mega3_trace.txt 2>&1; head -60 gpurun_out/c7_mega3_trace.txt
