#!/usr/bin/env bash
# short validation: whole GPU suite, the default bench line, the reference arm, the token-loop timeline
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 900 python -m pytest tests -m gpu -q --timeout=300 ) > gpurun_out/r2c_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2c_pytest_gpu.log
timeout 1200 python bench.py > gpurun_out/r2c_bench_full_n1.json 2> gpurun_out/r2c_bench_full_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/r2c_bench_full_n1.err
timeout 300 python tools/mega3_trace.py > gpurun_out/r2c_mega_trace.txt 2>&1
timeout 300 python tools/stage_times.py > gpurun_out/r2c_stage_times.txt 2>&1
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2c_bench_full_n1.json") if l.startswith("{")][-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "us/token", d["roofline"]["us_per_token"], "frac", d["roofline"]["frac"], d["stages_ms_per_song"],
      "cpu", d["cpu_baseline"]["value"], "oracle", d["oracle_check"]["match"], d["self_consistency"], d["dit_parity"], d["token_stream_sha1"])
PY
tail -4 gpurun_out/r2c_stage_times.txt; head -16 gpurun_out/r2c_mega_trace.txt
