set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python tools/stage_times.py 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --windows 6 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_2gpu_smoke.json 2> gpurun_out/bench_2gpu_smoke.err; tail -c 700 gpurun_out/bench_2gpu_smoke.json; tail -5 gpurun_out/bench_2gpu_smoke.err
