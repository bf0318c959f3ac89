#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_megakernel_ll -s 2 -c 1 -f -o gpurun_out/c17_mega_ll \
    python bench.py --windows 4 --steps 1 --warmup 1 --no-cpu-baseline --oracle-check none --dit 0 > gpurun_out/c17_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/c17_ncu.log; ls -la gpurun_out/
