set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
python tools/stage_times.py 2>&1 | tail -5
timeout 900 python bench.py --windows 24 --steps 2 --warmup 3 --cpu-windows 2 --cpu-threads 16 > gpurun_out/bench_mega24e.json 2> gpurun_out/bench_mega24e.err; tail -3 gpurun_out/bench_mega24e.err; python -c "
import json; d=json.load(open('gpurun_out/bench_mega24e.json')); print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step']); print(d['roofline']['us_per_token'], d['roofline']['frac'], d['roofline']['share_of_step']); print(d['cpu_baseline'])"
