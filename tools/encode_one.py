#!/usr/bin/env python
"""Encode ONE window a few times (the per-call encoder of the drop-in `server.model_generate` path) — target for an ncu launch list."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mapperatorinator_b200 import v29_model_config  # noqa: E402
from mapperatorinator_b200.modeling import B200Mapperatorinator  # noqa: E402
from mapperatorinator_b200.weights import init_model_state_dict  # noqa: E402

cfg = v29_model_config()
model = B200Mapperatorinator(cfg, init_model_state_dict(cfg, 0), max_windows=2, max_batch=1)
windows, _, _ = bench.segment(bench.synth_song(0, 20.0), cfg)
w = windows[:1].cuda()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    model.engine.encode(w, 0)
torch.cuda.synchronize()
