#!/usr/bin/env bash
# BASELINE config[3] per-GPU shape: 8 songs decoded in lock-step as one batch on one GPU (decode only, 40 windows per song)
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python bench.py --songs-per-gpu 8 --windows 40 --steps 1 --warmup 1 --no-cpu-baseline --oracle-check none --dit 0 > gpurun_out/r2c_bench_songs8.json 2> gpurun_out/r2c_bench_songs8.err; echo "rc=$?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2c_bench_songs8.json") if l.startswith("{")][-1])
    print("songs/gpu 8: value", d.get("value"), "e2e", d.get("e2e", {}).get("value"), "ms/step", d.get("ms_per_step"), d.get("self_consistency"), d.get("stages_ms_per_song"), d.get("error", "")[:300])
except Exception as e:
    print("no json", e)
PY
tail -2 gpurun_out/r2c_bench_songs8.err
