set -x
mkdir -p gpurun_out
timeout 900 python bench.py --windows 24 --steps 2 --warmup 3 --cpu-windows 2 > gpurun_out/bench_mega24d.json 2> gpurun_out/bench_mega24d.err; tail -3 gpurun_out/bench_mega24d.err; python -c "
import json; d=json.load(open('gpurun_out/bench_mega24d.json')); print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step']); print(d['roofline']); print(d['cpu_baseline'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --windows 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_b.log 2>&1; tail -2 gpurun_out/ncu_launch_b.log; wc -l gpurun_out/launches_r1b.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_megakernel -s 2 -c 1 -o gpurun_out/mega_full python bench.py --windows 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_mega.log 2>&1; tail -3 gpurun_out/ncu_mega.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_f32_kernel|mel_kernel|attention_kernel" -s 1 -c 6 -o gpurun_out/enc_full python bench.py --windows 16 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_enc.log 2>&1; tail -3 gpurun_out/ncu_enc.log
ls -la gpurun_out
