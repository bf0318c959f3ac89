set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python tools/mega_trace.py 2>&1 | tail -14
timeout 900 python bench.py --windows 24 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mega24c.json 2> gpurun_out/bench_mega24c.err; tail -3 gpurun_out/bench_mega24c.err
python - <<'PY'
import json
for f in ('mega24c',):
    try:
        d=json.load(open(f'gpurun_out/bench_{f}.json'))
        print(f, 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],1), 'us/token(all)', round(d['ms_per_step']*1000/(24*64),1))
    except Exception as e: print(f, 'ERR', e)
PY
