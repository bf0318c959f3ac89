"""CPU emulation of a single-pass TF32 DiT: the oracle's 100-step loop on the bench's T = 256 parity chunk with every token-wise GEMM operand
(mode "g") and the attention operands (mode "ga") rounded to TF32 first, compared with the fp32 oracle.  Measured: 1.03e-3 / 1.10e-3 max abs error,
i.e. over the 1e-3 north-star tolerance -- the reason the DiT GEMMs stay 3xTF32.  Test infrastructure: imports oracle/.

    python tools/dit_tf32_emulation.py g|ga
"""
import os, sys, time, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch.nn.functional as F
import bench
from bench import *
from oracle import dit as D
torch.set_num_threads(16)
def tf32(x):
    # round to nearest (ties away, like cvt.rna.tf32.f32) to 10 explicit mantissa bits
    i = x.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF
    return i.view(torch.float32)
mode = sys.argv[1]
orig_lin, orig_mm = F.linear, torch.matmul
def lin(x, w, b=None):
    if x.dim() >= 3:   # only the token-wise GEMMs (the [N,T,*] activations); the conditioning MLPs stay fp32
        return orig_lin(tf32(x), tf32(w), b)
    return orig_lin(x, w, b)
def mm(a, b):
    return orig_mm(tf32(a), tf32(b))
dc = dit_b_config(DIT_CLASSES); dsd = init_dit_state_dict(dc, 1)
seq_x, seq_c, y, y_null = synth_hit_objects(0)
Tp = 256
g = torch.Generator().manual_seed(5)
x = torch.cat([seq_x[None, :, :Tp]] * 2); c = torch.cat([seq_c[None, :, :Tp]] * 2); yy = torch.stack([y, y_null], 0)
noise = torch.randn(DIT_STEPS, 2, 2, Tp, generator=g)
ipm = torch.ones_like(x, dtype=torch.bool); ipm[:, :, :32] = False
def run():
    with torch.no_grad():
        return D.p_sample_loop(dsd, dc, D.Schedule(), x, c, yy, 1.0, D.band_mask(Tp, 128), noise, inpaint_mask=ipm)
ref = run()
if 'g' in mode: D.F.linear = lin; F.linear = lin
if 'a' in mode: D.torch.matmul = mm
t0=time.time(); out = run(); print(mode, 'max abs err', (out-ref).abs().max().item(), 'mean', (out-ref).abs().mean().item(), time.time()-t0)
