#!/usr/bin/env bash
# N-GPU run exactly as the driver launches it (torchrun, one rank per GPU, NCCL), then smoke(); usage: gpurun --gpus N -- 'bash tools/gpu_scale_check.sh N'
set -u
N=${1:-2}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 1 \
    > gpurun_out/r2c_bench_n$N.json 2> gpurun_out/r2c_bench_n$N.err; echo "bench N=$N rc=$?"
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2c_bench_n$N.json") if l.startswith("{")][-1])
    print("N=$N value", d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], d["self_consistency"], d["clocks"])
except Exception as e:
    print("no json", e)
PY
tail -3 gpurun_out/r2c_bench_n$N.err
timeout 300 python tools/stage_times.py 2>&1 | head -4
