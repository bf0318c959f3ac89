#!/usr/bin/env bash
# round-2 GPU call 1: parity at full dims, seed sweep of the bench's self-consistency, DiT step graph, first full bench line
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
tail -5 gpurun_out/c1_pytest.log
timeout 300 python tools/dit_bench.py 1 > gpurun_out/c1_dit_bench.txt 2>&1; tail -2 gpurun_out/c1_dit_bench.txt
for s in 1 2 3 4 5 6 7; do
  timeout 300 python bench.py --steps 1 --warmup 1 --song-seed $s --no-cpu-baseline > gpurun_out/c1_bench_seed$s.json 2> gpurun_out/c1_bench_seed$s.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/c1_bench_seed$s.json") if l.startswith("{")][-1])
    print("seed $s", d.get("value"), d.get("e2e",{}).get("value"), d.get("self_consistency"), d.get("error","")[:300])
except Exception as e:
    print("seed $s: no json", e)
PY
done
timeout 900 python bench.py > gpurun_out/c1_bench_full.json 2> gpurun_out/c1_bench_full.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/c1_bench_full.json
