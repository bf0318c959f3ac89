set -x
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/bench_r1_full.json 2> gpurun_out/bench_r1_full.err; tail -2 gpurun_out/bench_r1_full.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1_full.json')); print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step']); print(d['roofline']); print(d['cpu_baseline']); print(d['clocks'])"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_reference.json 2> gpurun_out/bench_r1_reference.err; tail -c 400 gpurun_out/bench_r1_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --windows 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_d.log 2>&1; wc -l gpurun_out/launches_r1d.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_megakernel -s 2 -c 1 -o gpurun_out/mega_full3 python bench.py --windows 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_mega3.log 2>&1; tail -2 gpurun_out/ncu_mega3.log
