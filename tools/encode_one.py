#!/usr/bin/env python
"""Encode ONE window (the per-call encoder of the drop-in `server.model_generate` path) or a chunk of N windows (the resident path) a few
times — target for ncu launch lists / captures:  python tools/encode_one.py [repetitions] [windows]"""
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
nw = int(sys.argv[2]) if len(sys.argv) > 2 else 1
model = B200Mapperatorinator(cfg, init_model_state_dict(cfg, 0), max_windows=max(2, nw), max_batch=1)
windows, _, _ = bench.segment(bench.synth_song(0, 20.0 + 8.2 * nw), cfg)
w = windows[:nw].cuda()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    model.engine.encode(w, 0)
torch.cuda.synchronize()
