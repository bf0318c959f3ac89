#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_fulldims.py -m gpu -q -x --timeout=180 -k "not dit" > gpurun_out/c29_pytest.log 2>&1
rc=$?; echo "pytest rc=$rc"; tail -5 gpurun_out/c29_pytest.log
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --oracle-check full --mega 2 --dit 0 > gpurun_out/c29_bench.json 2> gpurun_out/c29_bench.err
tail -3 gpurun_out/c29_bench.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/c29_bench.json") if l.startswith("{")][-1])
    print("value", d.get("value"), "e2e", d.get("e2e", {}).get("value"), "us/token", d.get("roofline", {}).get("us_per_token"), "sha", d.get("token_stream_sha1"), d.get("stages_ms_per_song"), d.get("oracle_check"), d.get("self_consistency"), d.get("error", "")[:300])
except Exception as e:
    print("no json", e)
PY
