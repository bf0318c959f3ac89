#!/usr/bin/env bash
# BASELINE config[3] shape at N GPUs: 8 songs per GPU decoded in lock-step, songs sharded over ranks, one NCCL gather of the token streams;
# usage: gpurun --gpus N -- 'bash tools/gpu_songs8_scale_check.sh N'
set -u
N=${1:-2}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --songs-per-gpu 8 --windows 40 \
    --steps 1 --warmup 1 --dit 0 > gpurun_out/r2d_bench_songs8_n$N.json 2> gpurun_out/r2d_bench_songs8_n$N.err; echo "bench songs8 N=$N rc=$?"
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2d_bench_songs8_n$N.json") if l.startswith("{")][-1])
    print("N=$N songs/gpu 8: value", d.get("value"), "e2e", d.get("e2e", {}).get("value"), "ms/step", d.get("ms_per_step"), d.get("self_consistency"), d.get("clocks"), d.get("error", "")[:400])
except Exception as e:
    print("no json", e)
PY
tail -3 gpurun_out/r2d_bench_songs8_n$N.err
