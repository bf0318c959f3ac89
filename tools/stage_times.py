#!/usr/bin/env python
"""Where does a window's time go? encode / prefill(+first token) / token loop at v29 dimensions, device vs host wall time."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mapperatorinator_b200 import TokenLayout, v29_model_config  # noqa: E402
from mapperatorinator_b200.modeling import B200Mapperatorinator  # noqa: E402
from mapperatorinator_b200.weights import init_model_state_dict  # noqa: E402

cfg = v29_model_config()
layout = TokenLayout.from_json(os.path.join(ROOT, "tests", "golden", "tokenizer_v29.json"))
model = B200Mapperatorinator(cfg, init_model_state_dict(cfg, 0), max_windows=16, max_batch=2)
windows, _, _ = bench.segment(bench.synth_song(0, 60.0), cfg)
w = windows[:16].cuda()


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1000 / n


print("encode 16 windows        device %.2f ms  wall %.2f ms" % timed(lambda: model.engine.encode(w, 0)))
print("encode 1 window          device %.2f ms  wall %.2f ms   (CUDA-graph replay)" % timed(lambda: model.engine.encode(w[:1], 0)))
model.engine.set_option("enc_graph", 0)
print("encode 1 window          device %.2f ms  wall %.2f ms   (eager launches)" % timed(lambda: model.engine.encode(w[:1], 0)))
model.engine.set_option("enc_graph", 1)
prompt = torch.tensor([bench.prompt_for(1, [list(range(100, 164))])])
P = prompt.shape[1]
for new in (1, 2, 64):
    gk = bench.gen_kwargs(1, 211, P)
    gk["max_length"] = P + new
    gk["min_new_tokens"] = new
    d, wl = timed(lambda: model.engine.generate([1], prompt, prompt.ne(0), layout, gk))
    print(f"generate P={P} new={new:3d}   device {d:.2f} ms  wall {wl:.2f} ms")
