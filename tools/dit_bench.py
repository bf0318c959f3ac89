#!/usr/bin/env python
"""Stage (iii) timing: DiT-B 100-step refinement of a 1024-point chunk (BASELINE config[2] building block), fused on-device loop.
Prints ms/step and achieved TFLOP/s against the dense-attention FLOP count of SURVEY §8d (425 GFLOP/step at T=1024, B=2)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mapperatorinator_b200 import _lib, dit_b_config  # noqa: E402
from mapperatorinator_b200.diffusion import B200DiT, InpaintDenoiser, band_attention_mask, create_diffusion  # noqa: E402
from mapperatorinator_b200.weights import init_dit_state_dict  # noqa: E402

tc = int(sys.argv[1]) if len(sys.argv) > 1 else 1
_lib.load().mb200_set_tensor_cores(tc)
dc = dit_b_config(600)
dit = B200DiT(dc, init_dit_state_dict(dc, 1), max_seq_len=1024)
T = 1024
g = torch.Generator().manual_seed(0)
x = torch.rand(1, 2, T, generator=g) * 2 - 1
x = torch.cat([x, x]).cuda()
c = torch.randn(1, dc.context_size, T, generator=g); c = torch.cat([c, c]).cuda()
y = (torch.rand(2, dc.class_size, generator=g) < 0.05).float().cuda()
diff = create_diffusion([100, 0, 0, 0, 0, 0, 0, 0, 0, 0], "squaredcos_cap_v2", 1000)
noise = torch.randn(100, 2, 2, T, generator=g).cuda()
mask = torch.ones_like(x, dtype=torch.bool)
am = band_attention_mask(T, 128, "cuda")
mk = dict(c=c, y=y, cfg_scale=1.0, attn_mask=am, key_padding_mask=None)
run = lambda: diff.p_sample_loop(dit.forward_with_cfg, x.shape, x, denoised_fn=InpaintDenoiser(mask, x), model_kwargs=mk, step_noise=noise)
out = run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(2):
    out = run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 2
d, L = dc.hidden, dc.depth
flop_step = 2 * T * L * (2 * (4 * d * d + 2 * d * 4 * d) + 2 * 2 * d * T)      # dense-attention accounting
print(f"tensor_cores={tc} DiT-B T={T} 100 steps: {ms:.1f} ms total, {ms / 100:.2f} ms/step, {flop_step * 100 / (ms * 1e-3) / 1e12:.1f} TFLOP/s "
      f"(algorithmic {flop_step / 1e9:.0f} GFLOP/step), finite={bool(torch.isfinite(out).all())}")
