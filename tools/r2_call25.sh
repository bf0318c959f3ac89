#!/usr/bin/env bash
# same box, back to back: grid-barrier megakernel vs dataflow megakernel on the whole song (decode only); then the tensor-core attention tests
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout=180 -k "attention" > gpurun_out/c25_pytest_attn.log 2>&1; echo "attention pytest rc=$?"; tail -12 gpurun_out/c25_pytest_attn.log
for m in 1 2 1 2; do
  timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --oracle-check none --mega $m --dit 0 > gpurun_out/c25_bench_mega$m.json 2> gpurun_out/c25_bench_mega$m.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/c25_bench_mega$m.json") if l.startswith("{")][-1])
    print("mega $m: value", d.get("value"), "e2e", d.get("e2e", {}).get("value"), "us/token", d.get("roofline", {}).get("us_per_token"), "sha", d.get("token_stream_sha1"), d.get("stages_ms_per_song"), d.get("error", "")[:300])
except Exception as e:
    print("mega $m: no json", e)
PY
done
