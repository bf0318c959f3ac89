#!/usr/bin/env python
"""Timeline of the dataflow token-loop megakernel at v29 dimensions: token 9 of a window, one CTA (MB200_TRACE_CTA, default 0), 16 clock64
stamps per phase taken by thread 0 (epilogue stamp: thread 192).
GEMV phase columns (us):   pre   phase top -> first poll loads issued
                           misc  -> weight prefetch issued / LayerNorm weights requested / epilogue operands fetched
                           wts   -> weight slice of this phase (requested a phase ago) has landed
                           poll  -> thread 0's tagged input arrived          (fails = unsuccessful poll rounds of thread 0)
                           ln1/ln2  -> first / second LayerNorm statistic known (named barrier over the column warps)
                           fma   -> partial sums written;  bar -> CTA barrier passed;  epi -> output row 0 stored (thread 192)
attention phase columns:   load  phase top -> K/V cache rows requested;  q -> q (and the new k|v row) polled + CTA barrier
                           score -> scores in shared memory (barrier);  pv -> softmax + PV partials (barrier);  out -> stored;  merge -> split merge
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
model.engine.set_option("mega", 2)
model.engine.set_option("mega_trace", 1)
for name, env in (("ll_reps", "MB200_LL_REPS"), ("ll_sleep", "MB200_LL_SLEEP"), ("ll_debug", "MB200_LL_DEBUG"), ("trace_cta", "MB200_TRACE_CTA")):
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
t = raw[:n].astype(np.float64)
names = ["qkv", "self_attn", "out", "q_c", "cross_attn", "out_c", "fc1", "fc2"]
us = 1.0 / 1.965e3
total = (t[-1, 8] - t[0, 0]) * us
print(f"CTA {os.environ.get('MB200_TRACE_CTA', '0')}: token total {total:.1f} us")
gem, att = {}, {}
for i in range(n):
    nm = names[i % 8] if i < 96 else ("proj_out" if i == 96 else "sample")
    s = t[i]
    nxt = t[i + 1, 0] if i + 1 < n else s[8]
    d = lambda a, b: (s[a] - s[b]) * us if s[a] > 0 and s[b] > 0 else 0.0
    if nm in ("self_attn", "cross_attn"):
        att.setdefault(nm, []).append((d(1, 0), d(2, 1), d(3, 2), d(4, 3), d(5, 4), d(6, 5), (s[8] - s[0]) * us, (nxt - s[0]) * us))
    elif nm == "sample":      # logits polled | processor chain | argmax | append + embedding | state + header
        att.setdefault(nm, []).append((d(1, 0), d(2, 1), d(3, 2), d(4, 3), d(5, 4), 0, (s[8] - s[0]) * us, (nxt - s[0]) * us))
        print(f"  sample detail: state loaded {d(6, 0):.2f}  first poll chunk {d(7, 6):.2f} (failed rounds {int(raw[i, 10])})  second chunk {d(1, 7):.2f}")
    else:
        if s[3] > s[6]:      # row-per-warp phase: weights are waited for after the barrier that publishes the activation
            row = (d(1, 0), d(2, 1), d(3, 6), d(4, 2), raw[i, 10], d(5, 4), d(6, 5) if s[5] > 0 else d(6, 4), d(7, 3), 0.0, d(8, 7))
        else:                # K-split phase (fc2)
            row = (d(1, 0), d(2, 1), d(3, 2), d(4, 3), raw[i, 10], 0.0, d(6, 4), d(7, 6), d(8, 7), d(9, 8))
        gem.setdefault(nm, []).append(row + ((s[8] - s[0]) * us, (nxt - s[0]) * us))
print("  GEMV phase     n     pre   misc    wts   poll  fails    ln1    ln2    fma    bar    epi | to-barrier  to-next")
for nm, v in gem.items():
    a = np.array(v, dtype=np.float64).mean(0)
    print(f"  {nm:12s} {len(v):3d} " + " ".join(f"{x:6.2f}" for x in a[:10]) + f" | {a[10]:8.2f} {a[11]:8.2f}")
print("  attention      n    load      q  score     pv    out  merge | to-barrier  to-next")
for nm, v in att.items():
    a = np.array(v, dtype=np.float64).mean(0)
    print(f"  {nm:12s} {len(v):3d} " + " ".join(f"{x:6.2f}" for x in a[:6]) + f" | {a[6]:8.2f} {a[7]:8.2f}")
