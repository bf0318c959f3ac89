#!/usr/bin/env bash
# round-2 GPU call 2: dataflow megakernel parity first; only if green: the whole GPU suite and the bench
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -x --timeout=180 -k "greedy and (dataflow or megakernel) and torchaudio" > gpurun_out/c2_pytest_dataflow.log 2>&1
rc=$?; echo "dataflow pytest rc=$rc"; tail -15 gpurun_out/c2_pytest_dataflow.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 1200 python -m pytest tests -m gpu -q --maxfail=8 --timeout=300 > gpurun_out/c2_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -15 gpurun_out/c2_pytest_all.log
timeout 400 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --mega 1 > gpurun_out/c2_bench_mega1.json 2> gpurun_out/c2_bench_mega1.err
timeout 400 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --mega 2 > gpurun_out/c2_bench_mega2.json 2> gpurun_out/c2_bench_mega2.err
python - <<'PY'
import json
for m in (1, 2):
    try:
        d = json.loads([l for l in open(f"gpurun_out/c2_bench_mega{m}.json") if l.startswith("{")][-1])
        print("mega", m, "value", d.get("value"), "e2e", d.get("e2e", {}).get("value"), "us/token", d.get("roofline", {}).get("us_per_token"), "sha", d.get("token_stream_sha1"),
              d.get("self_consistency"), d.get("stages_ms_per_song"), d.get("error", "")[:500])
    except Exception as e:
        print("mega", m, "no json", e)
PY
timeout 900 python bench.py > gpurun_out/c2_bench_full.json 2> gpurun_out/c2_bench_full.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/c2_bench_full.json
