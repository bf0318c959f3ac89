#!/usr/bin/env python
"""Fine timeline of the dataflow megakernel with K-split GEMV phases (token 9 of a window, CTAs 0 and 1), 8 stamps per phase:
   poll  = phase start -> thread 0's tagged input arrived          (fails = unsuccessful poll rounds of thread 0)
   ln    = polled -> input ready in registers (LayerNorm statistics / shared-memory hand-over, 1-2 CTA barriers)
   wts   = input ready -> weight slice landed
   fma   = multiply + warp butterfly + partials written
   bar   = the phase's CTA barrier
   epi   = barrier -> output row 0 stored (epilogue thread)
MB200_LL_DEBUG bits: 1 no weight copies, 2 polls never wait, 4 no multiply-reduce (timing diagnostics; tokens are garbage)."""
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
model.engine.set_option("mega", 3)
model.engine.set_option("mega_trace", 1)
for name, env in (("ll_reps", "MB200_LL_REPS"), ("ll_sleep", "MB200_LL_SLEEP"), ("ll_debug", "MB200_LL_DEBUG")):
    if os.environ.get(env):
        model.engine.set_option(name, int(os.environ[env]))
prompt = torch.tensor([bench.prompt_for(0, [])])
for _ in range(3):
    try:
        model.engine.generate([0], prompt, prompt.ne(0), layout, bench.gen_kwargs(0, 211, prompt.shape[1]))
    except RuntimeError as e:          # diagnostics modes may trip the engine's own checks
        print("generate raised:", str(e)[:200])
n = 12 * 8 + 2
raw = np.zeros((128, 16), dtype=np.uint64)
_lib.check(_lib.load().mb200_model_read_trace(model.engine.handle, raw.ctypes.data, 128))
t = raw.reshape(-1)[: 2 * n * 8].reshape(2, n, 8)
names = ["qkv", "self_attn", "out", "q_c", "cross_attn", "out_c", "fc1", "fc2"]
us = 1.0 / 1.965e3
for cta in (0, 1):
    tt = t[cta].astype(np.float64)
    total = (tt[-1, 5] - tt[0, 0]) * us
    print(f"CTA {cta}: token total {total:.1f} us")
    agg = {}
    for i in range(n):
        nm = names[i % 8] if i < 96 else ("proj_out" if i == 96 else "sample")
        s = tt[i]
        nxt = tt[i + 1, 0] if i + 1 < n else s[5]
        if nm in ("self_attn", "cross_attn", "sample"):
            agg.setdefault(nm, []).append((0, 0, 0, 0, 0, 0, 0, (s[5] - s[0]) * us, (nxt - s[0]) * us))
        else:
            agg.setdefault(nm, []).append(((s[1] - s[0]) * us, t[cta][i, 7], (s[2] - s[1]) * us, (s[3] - s[2]) * us, (s[4] - s[3]) * us, (s[5] - s[4]) * us,
                                           (s[6] - s[5]) * us, (s[5] - s[0]) * us, (nxt - s[0]) * us))
    print("  phase          n    poll  fails     ln    wts    fma    bar    epi | to-barrier  to-next-phase")
    for nm, v in agg.items():
        a = np.array(v, dtype=np.float64).mean(0)
        print(f"  {nm:12s} {len(v):3d} {a[0]:7.2f} {a[1]:6.1f} {a[2]:6.2f} {a[3]:6.2f} {a[4]:6.2f} {a[5]:6.2f} {a[6]:6.2f} | {a[7]:8.2f} {a[8]:8.2f}")
