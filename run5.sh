set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
# CPU thread scaling of the oracle decode (which thread count is the fair CPU baseline?)
python - <<'PY' 2>&1 | tail -8
import os, sys, time, torch
sys.path.insert(0, '.')
from mapperatorinator_b200 import TokenLayout, v29_model_config
from mapperatorinator_b200.weights import init_model_state_dict
from oracle import generate as G
import bench
cfg = v29_model_config(); lay = TokenLayout.from_json('tests/golden/tokenizer_v29.json'); sd = init_model_state_dict(cfg, 0)
w, _, _ = bench.segment(bench.synth_song(0), cfg)
for th in (8, 16, 32, 64, 128):
    torch.set_num_threads(th)
    p = torch.tensor([bench.prompt_for(0, [])])
    gk = bench.gen_kwargs(0, 211, p.shape[1]); gk['max_length'] = p.shape[1] + 24; gk['min_new_tokens'] = 24
    t = time.perf_counter()
    with torch.no_grad():
        ids, st = G.model_generate(sd, cfg, lay, dict(inputs=w[0:1], decoder_input_ids=p, decoder_attention_mask=p.ne(0)), gk)
    print('threads', th, 'tok/s', 24 / (time.perf_counter() - t), flush=True)
PY
timeout 900 python bench.py --windows 24 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mega24.json 2> gpurun_out/bench_mega24.err; tail -c 1200 gpurun_out/bench_mega24.json; tail -3 gpurun_out/bench_mega24.err
timeout 900 python bench.py --windows 24 --steps 2 --warmup 3 --no-cpu-baseline --mega 0 > gpurun_out/bench_graph24.json 2> gpurun_out/bench_graph24.err; tail -c 1200 gpurun_out/bench_graph24.json; tail -3 gpurun_out/bench_graph24.err
timeout 900 python bench.py --windows 24 --steps 2 --warmup 3 --no-cpu-baseline --mega 0 --pdl 1 > gpurun_out/bench_pdl24.json 2> gpurun_out/bench_pdl24.err; tail -c 600 gpurun_out/bench_pdl24.json; tail -3 gpurun_out/bench_pdl24.err
