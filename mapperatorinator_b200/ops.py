"""Kernel-level wrappers over the C ABI (`mb200_op_*`) for parity tests; operands are CUDA torch tensors."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib

ACT = {"none": 0, "gelu": 1, "gelu_tanh": 2, "silu": 3}
MASK = {"none": 0, "causal": 1, "band": 2, "dense": 3}


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32c(t: torch.Tensor) -> torch.Tensor:
    assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
    return t


def gemm(a, w, bias=None, act="none", alpha=1.0, residual=None, gate=None, gate_rows_per_batch=1):
    """act(a @ w.T + bias) * alpha [* gate[row // rpb]] [+ residual]."""
    lib = _lib.load()
    a, w = _f32c(a), _f32c(w)
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty(M, N, device=a.device, dtype=torch.float32)
    _lib.check(lib.mb200_op_gemm(_ptr(a), K, _ptr(w), K, _ptr(out), N, _ptr(bias), ACT[act], float(alpha), _ptr(residual), N,
                                 _ptr(gate), (gate.shape[1] if gate is not None else 0), int(gate_rows_per_batch), M, N, K, _stream()))
    return out


def gemm_tc(a, w, bias=None, act="none", alpha=1.0, residual=None):
    """Same contract as `gemm`, forced through the tcgen05 3xTF32 kernel."""
    lib = _lib.load()
    a, w = _f32c(a), _f32c(w)
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty(M, N, device=a.device, dtype=torch.float32)
    _lib.check(lib.mb200_op_gemm_tc(_ptr(a), K, _ptr(w), K, _ptr(out), N, _ptr(bias), ACT[act], float(alpha), _ptr(residual), N, M, N, K, _stream()))
    return out


def layernorm(x, weight=None, bias=None, shift=None, scale=None, rows_per_batch=1, eps=1e-5):
    lib = _lib.load()
    x = _f32c(x)
    rows, dim = x.shape
    y = torch.empty_like(x)
    _lib.check(lib.mb200_op_layernorm(_ptr(x), _ptr(y), _ptr(weight), _ptr(bias), _ptr(shift), _ptr(scale), int(rows_per_batch), rows, dim,
                                      float(eps), _stream()))
    return y


def attention(q, k, v, heads, scale=1.0, mask="none", q_pos0=0, key_valid=None, band=0, dense_mask=None):
    """q (B, Tq, H*64), k/v (B, Tk, H*64) token-major -> (B, Tq, H*64)."""
    lib = _lib.load()
    q, k, v = _f32c(q), _f32c(k), _f32c(v)
    B, Tq, _ = q.shape
    Tk = k.shape[1]
    o = torch.empty_like(q)
    _lib.check(lib.mb200_op_attention(_ptr(q), _ptr(k), _ptr(v), _ptr(o), B, heads, Tq, Tk, float(scale), MASK[mask], int(q_pos0),
                                      _ptr(key_valid), int(band), _ptr(dense_mask), _stream()))
    return o
