timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --windows 24 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'],'e2e',d['e2e']['value'],'us/token',d['roofline']['us_per_token'],'frac',d['roofline']['frac'],'sha',d['token_stream_sha1'])"
