#!/usr/bin/env python
"""Per-phase timeline of the DATAFLOW token-loop megakernel at v29 dimensions (token 9 of a window; CTAs 0, 1, 100, 140):
   wait  = phase start -> input staged (the dataflow wait for the producers + LayerNorm / copy into shared memory)
   wts   = input staged -> weight slice landed (0 when the bulk prefetch was early enough)
   work  = weights -> rows / attention units done (incl. the end-of-phase CTA barrier)
clock64 of the CTA's SM at 1.965 GHz."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mapperatorinator_b200 import TokenLayout, _lib, v29_model_config  # noqa: E402
from mapperatorinator_b200.modeling import B200Mapperatorinator  # noqa: E402
from mapperatorinator_b200.weights import init_model_state_dict  # noqa: E402

cfg = v29_model_config()
layout = TokenLayout.from_json(os.path.join(ROOT, "tests", "golden", "tokenizer_v29.json"))
model = B200Mapperatorinator(cfg, init_model_state_dict(cfg, 0), max_windows=2, max_batch=2)
windows, _, _ = bench.segment(bench.synth_song(0, 20.0), cfg)
model.engine.encode(windows[:2].cuda(), 0)
model.engine.set_option("mega", int(os.environ.get("MB200_MEGA", "3")))
model.engine.set_option("mega_trace", 1)
if os.environ.get("MB200_LL_REPS"):
    model.engine.set_option("ll_reps", int(os.environ["MB200_LL_REPS"]))
if os.environ.get("MB200_LL_SLEEP"):
    model.engine.set_option("ll_sleep", int(os.environ["MB200_LL_SLEEP"]))
prompt = torch.tensor([bench.prompt_for(0, [])])
for _ in range(3):
    model.engine.generate([0], prompt, prompt.ne(0), layout, bench.gen_kwargs(0, 211, prompt.shape[1]))
n = 12 * 8 + 2
raw = np.zeros((128, 16), dtype=np.uint64)
_lib.check(_lib.load().mb200_model_read_trace(model.engine.handle, raw.ctypes.data, 128))
t = raw.reshape(-1)[: 4 * n * 4].reshape(4, n, 4).astype(np.float64) / 1.965e3      # us
names = ["qkv", "self_attn", "out", "q_c", "cross_attn", "out_c", "fc1", "fc2"]
for ci, cta in enumerate((0, 1, 100, 140)):
    tt = t[ci]
    total = tt[-1, 3] - tt[0, 0]
    print(f"CTA {cta}: token total {total:.1f} us")
    agg = {}
    for i in range(n):
        nm = names[i % 8] if i < 96 else ("proj_out" if i == 96 else "sample")
        s0, s1, s2, s3 = tt[i]
        if nm in ("self_attn", "cross_attn", "sample"):
            agg.setdefault(nm, []).append((0.0, 0.0, s3 - s0))
        else:
            agg.setdefault(nm, []).append((s1 - s0, s2 - s1, s3 - s2))
    print("  phase          n    wait    wts    work   total")
    for nm, v in agg.items():
        a = np.array(v).mean(0)
        print(f"  {nm:12s} {len(v):3d} {a[0]:7.2f} {a[1]:6.2f} {a[2]:7.2f} {a.sum():7.2f}")
