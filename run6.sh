set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 900 python bench.py --windows 24 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mega24b.json 2> gpurun_out/bench_mega24b.err; tail -3 gpurun_out/bench_mega24b.err
timeout 900 python bench.py --windows 24 --steps 2 --warmup 3 --no-cpu-baseline --mega 0 > gpurun_out/bench_graph24b.json 2> gpurun_out/bench_graph24b.err; tail -3 gpurun_out/bench_graph24b.err
python - <<'PY'
import json
for f in ('mega24b','graph24b'):
    try:
        d=json.load(open(f'gpurun_out/bench_{f}.json'))
        print(f, 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],1), 'us/token(all)', round(d['ms_per_step']*1000/(24*64),1), d['roofline']['per_token_us'])
    except Exception as e: print(f, 'ERR', e)
PY
