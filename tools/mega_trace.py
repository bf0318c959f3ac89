#!/usr/bin/env python
"""Per-phase timeline of the persistent decode megakernel (CTA 0, 9th token) at v29 dimensions: where does a token's time go?"""
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
model.engine.set_option("mega_trace", 1)
prompt = torch.tensor([bench.prompt_for(0, [])])
for _ in range(3):
    model.engine.generate([0], prompt, prompt.ne(0), layout, bench.gen_kwargs(0, 211, prompt.shape[1]))
n = 12 * 8 + 2
out = np.zeros((n, 16), dtype=np.uint64)
_lib.check(_lib.load().mb200_model_read_trace(model.engine.handle, out.ctypes.data, n))
t = (out.astype(np.float64) / 1.965).astype(np.int64)      # SM cycles (clock64, 1965 MHz under load) -> ns
names = ["qkv", "self_attn", "out", "q_c", "cross_attn", "out_c", "fc1", "fc2"]
tot = t[-1, 5] - t[0, 0]
print(f"token total {tot / 1e3:.1f} us over {n} phases")
print("phase            own_stage (issue, loads, rest)  sync   pref+wait   math   barrier    total (us)")
agg, warm, own, attn, rowt = {}, {}, {}, {}, {}
for i in range(n):
    nm = names[i % 8] if i < 96 else ("proj_out" if i == 96 else "sample")
    s0, s1, s2, s3, s4, s5 = (t[i, j] for j in range(6))
    if nm in ("self_attn", "cross_attn", "sample"):
        seg = (0, 0, 0, 0, 0, 0, s4 - s0, s5 - s4)
        own.setdefault(nm, []).append((t[i, 11] - s0, 0))
        if nm != "sample":
            attn.setdefault(nm, []).append((t[i, 6] - s0, t[i, 7] - t[i, 6], t[i, 8] - t[i, 7], t[i, 9] - t[i, 8]))
    else:
        s6, s7, c0, c1, c2 = t[i, 6], t[i, 7], t[i, 8], t[i, 9], t[i, 10]
        # trace mode stages twice: cold pass (c0 -> c2) then warm pass (s6 -> s1); report the cold one as the stage, warm beside it
        ld = (c0 - s0, (c1 - c0) if c1 > 0 else 0, (c2 - c1) if c1 > 0 else c2 - c0)
        warm.setdefault(nm, []).append(((s7 - s6) if s7 > 0 else 0, (s1 - s7) if s7 > 0 else s1 - s6))
        seg = (c2 - s0,) + ld + (s2 - s1, s3 - s2, s4 - s3, s5 - s4)
        own.setdefault(nm, []).append((t[i, 12] - s3, t[i, 11] - t[i, 12]))
        rowt.setdefault(nm, []).append((t[i, 13] - t[i, 12], t[i, 14] - t[i, 13], t[i, 11] - t[i, 14]))
    agg.setdefault(nm, []).append(seg + (s5 - s0,))
for nm, v in agg.items():
    a = np.array(v, dtype=np.float64).mean(0) / 1e3
    print(f"{nm:12s} x{len(v):3d} {a[0]:8.2f} ({a[1]:5.2f} {a[2]:5.2f} {a[3]:5.2f}) {a[4]:8.2f} {a[5]:8.2f} {a[6]:7.2f} {a[7]:9.2f} {a[8]:8.2f}"
          + f"   warp 0's rows cold {np.mean([o[0] for o in own[nm]]) / 1e3:.2f} warm {np.mean([o[1] for o in own[nm]]) / 1e3:.2f}"
          + (f"   [K/V loaded+scores {np.mean([x[0] for x in attn[nm]]) / 1e3:.2f} softmax {np.mean([x[1] for x in attn[nm]]) / 1e3:.2f} PV+partials {np.mean([x[2] for x in attn[nm]]) / 1e3:.2f} ticket(+merge) {np.mean([x[3] for x in attn[nm]]) / 1e3:.2f}]" if nm in attn else "")
          + (f"   first row (warm): dot {np.mean([x[0] for x in rowt[nm]]) / 1e3:.2f} shuffle-sum {np.mean([x[1] for x in rowt[nm]]) / 1e3:.2f} epilogue(+more rows) {np.mean([x[2] for x in rowt[nm]]) / 1e3:.2f}" if nm in rowt else "")
          + (f"   warm re-run of the staging: loads {np.mean([w[0] for w in warm[nm]]) / 1e3:.2f} rest {np.mean([w[1] for w in warm[nm]]) / 1e3:.2f}" if nm in warm else ""))
