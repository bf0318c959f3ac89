set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/stage_times.py 2>&1 | tail -8
timeout 300 python tools/dit_bench.py 1 2>&1 | tail -6
timeout 300 python tools/dit_bench.py 0 2>&1 | tail -6
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tf32x3_kernel|mel_kernel" -c 8 -o gpurun_out/tc_full2 python bench.py --windows 16 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_tc2.log 2>&1; tail -2 gpurun_out/ncu_tc2.log
