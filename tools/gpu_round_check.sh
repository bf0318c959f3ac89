#!/usr/bin/env bash
# The measurement set behind profiles/r1_*.md, as run on the GPU box:  gpurun --timeout 3000 -- 'bash tools/gpu_round_check.sh'
# Everything lands in gpurun_out/ (scratch); the summaries that are judged are copied into profiles/ by hand.
set -x
mkdir -p gpurun_out build_tools
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 1200 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
# launch list of the bench command (small song: ncu serialises every launch) and one full capture of the dominant kernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --windows 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_megakernel -s 2 -c 1 -o gpurun_out/mega_full \
    python bench.py --windows 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_mega.log 2>&1
# per-phase timeline of the token loop, stage times, DiT loop, decode attention at batch, the L2 / barrier micro-benchmark
timeout 300 python tools/mega_trace.py > gpurun_out/mega_trace.txt 2>&1
timeout 300 python tools/stage_times.py > gpurun_out/stage_times.txt 2>&1
timeout 300 python tools/dit_bench.py 1 > gpurun_out/dit_bench.txt 2>&1
timeout 600 python tools/attn_bench.py 1 8 32 64 > gpurun_out/attn_bench.txt 2>&1
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build_tools/l2_hotspot_bench tools/l2_hotspot_bench.cu && timeout 120 ./build_tools/l2_hotspot_bench > gpurun_out/l2_hotspot.txt 2>&1
tail -n 3 gpurun_out/*.txt
