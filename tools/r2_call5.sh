#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -q -x --timeout=180 -k "greedy and dataflow and torchaudio" > gpurun_out/c5_pytest.log 2>&1; echo "dataflow rc=$?"; tail -3 gpurun_out/c5_pytest.log
for cfgs in "1 0" "4 0" "8 0" "16 0" "8 300" "16 300" "16 1000"; do
  set -- $cfgs
  MB200_LL_REPS=$1 MB200_LL_SLEEP=$2 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --mega 2 --dit 0 --windows 40 > gpurun_out/c5_bench_r$1_s$2.json 2> gpurun_out/c5_bench_r$1_s$2.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/c5_bench_r$1_s$2.json") if l.startswith("{")][-1])
    print("reps $1 sleep $2: us/token", d.get("roofline", {}).get("us_per_token"), "sha", d.get("token_stream_sha1"), d.get("error", "")[:300])
except Exception as e:
    print("reps $1 sleep $2: no json", e)
PY
done
MB200_LL_REPS=16 timeout 200 python tools/mega2_trace.py > gpurun_out/c5_mega2_trace.txt 2>&1; head -28 gpurun_out/c5_mega2_trace.txt
