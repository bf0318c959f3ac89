"""Oracle: the osuT5 encoder / KV-cached decoder as the reference runs it at v29 (stock HF Whisper backbone).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Plain CPU torch fp32, weights by reference state_dict names.

Follows:
  * `OsuTEncoder.forward` (osuT5/osuT5/model/modeling_mapperatorinator.py:392-443): mel -> encoder_embedder Linear
    (:433) -> swapaxes (:436) -> HF WhisperEncoder.
  * HF `WhisperEncoder.forward` (transformers 5.5.0 models/whisper/modeling_whisper.py:592-640): gelu(conv1 k3 p1),
    gelu(conv2 k3 s2 p1), + embed_positions, pre-LN layers, final LN.
  * HF `WhisperAttention.forward` (:286-358): q = (q_proj(x)) * head_dim**-0.5 BEFORE the attention call (scaling=1.0),
    k_proj has no bias.
  * HF `WhisperDecoder.forward` (:700-790): inputs_embeds (= decoder_embedder(ids), modeling_mapperatorinator.py:205-207)
    + embed_positions[arange(len) + past_len] (pad tokens consume position indices, SURVEY §7 hard parts), causal mask
    AND key-padding mask from decoder_attention_mask.
  * `WhisperForConditionalGeneration`: proj_out (no bias).
third-party arithmetic: transformers==5.5.0 as installed (reference pins 4.57.3; see SURVEY §7 version skew).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from . import mel as mel_oracle

W = Dict[str, torch.Tensor]


def _ln(x, w: W, prefix: str, eps: float = 1e-5):
    return F.layer_norm(x, (x.shape[-1],), w[prefix + "weight"], w[prefix + "bias"], eps)


def _split_heads(x, heads):  # (B, T, D) -> (B, H, T, hd)
    B, T, D = x.shape
    return x.view(B, T, heads, D // heads).transpose(1, 2)


def _attn(q, k, v, mask_add: Optional[torch.Tensor]):
    """softmax(q k^T + mask) v with scaling 1.0 (the query is pre-scaled)."""
    s = torch.matmul(q, k.transpose(2, 3))
    if mask_add is not None:
        s = s + mask_add
    p = torch.softmax(s, dim=-1)
    p = torch.nan_to_num(p, nan=0.0)  # fully-masked (pad) query rows: SDPA returns 0 for them
    o = torch.matmul(p, v)
    B, H, T, hd = o.shape
    return o.transpose(1, 2).reshape(B, T, H * hd)


def encoder_forward(w: W, cfg, mel_frames: torch.Tensor) -> torch.Tensor:
    """(B, 1024, n_mels) -> (B, 512, d)."""
    x = F.linear(mel_frames, w["encoder_embedder.weight"], w["encoder_embedder.bias"])   # (B, 1024, d)
    x = x.swapaxes(1, 2)                                                                  # (B, d, 1024)
    p = "transformer.model.encoder."
    x = F.gelu(F.conv1d(x, w[p + "conv1.weight"], w[p + "conv1.bias"], padding=1))
    x = F.gelu(F.conv1d(x, w[p + "conv2.weight"], w[p + "conv2.bias"], stride=2, padding=1))
    x = x.permute(0, 2, 1) + w[p + "embed_positions.weight"]
    scale = cfg.head_dim ** -0.5
    for i in range(cfg.encoder_layers):
        lp = f"{p}layers.{i}."
        h = _ln(x, w, lp + "self_attn_layer_norm.")
        q = F.linear(h, w[lp + "self_attn.q_proj.weight"], w[lp + "self_attn.q_proj.bias"]) * scale
        k = F.linear(h, w[lp + "self_attn.k_proj.weight"])
        v = F.linear(h, w[lp + "self_attn.v_proj.weight"], w[lp + "self_attn.v_proj.bias"])
        a = _attn(_split_heads(q, cfg.heads), _split_heads(k, cfg.heads), _split_heads(v, cfg.heads), None)
        x = x + F.linear(a, w[lp + "self_attn.out_proj.weight"], w[lp + "self_attn.out_proj.bias"])
        h = _ln(x, w, lp + "final_layer_norm.")
        h = F.gelu(F.linear(h, w[lp + "fc1.weight"], w[lp + "fc1.bias"]))
        x = x + F.linear(h, w[lp + "fc2.weight"], w[lp + "fc2.bias"])
    return _ln(x, w, p + "layer_norm.")


def encode(w: W, cfg, pcm: torch.Tensor) -> torch.Tensor:
    """raw PCM (B, 130944) -> encoder states (B, 512, d): MelSpectrogram (always fp32) then the encoder."""
    return encoder_forward(w, cfg, mel_oracle.mel_forward(pcm, cfg.mel))


class DecoderState:
    """Static KV cache like `cache_utils.get_cache` (osuT5/osuT5/inference/cache_utils.py:23-35): self K/V grow with
    every fed token (pads included), cross K/V are computed once from the encoder states."""

    def __init__(self, w: W, cfg, enc: torch.Tensor):
        self.w, self.cfg = w, cfg
        B = enc.shape[0]
        self.B = B
        self.len = 0
        self.key_valid = torch.zeros(B, 0, dtype=torch.bool)      # decoder_attention_mask, extended with True
        self.k: List[Optional[torch.Tensor]] = [None] * cfg.decoder_layers
        self.v: List[Optional[torch.Tensor]] = [None] * cfg.decoder_layers
        self.ck, self.cv = [], []
        for i in range(cfg.decoder_layers):
            lp = f"transformer.model.decoder.layers.{i}.encoder_attn."
            self.ck.append(_split_heads(F.linear(enc, w[lp + "k_proj.weight"]), cfg.heads))
            self.cv.append(_split_heads(F.linear(enc, w[lp + "v_proj.weight"], w[lp + "v_proj.bias"]), cfg.heads))


def decoder_forward(st: DecoderState, ids: torch.Tensor, attn_mask: Optional[torch.Tensor],
                    position_rule: str = "arange", last_only: bool = False) -> torch.Tensor:
    """Feed `ids` (B, n) after the `st.len` tokens already cached; returns logits (B, n, V_out) f32.
    attn_mask (B, n) bool marks real (non-pad) tokens among the NEW ids (None = all real)."""
    w, cfg = st.w, st.cfg
    B, n = ids.shape
    if attn_mask is None:
        attn_mask = torch.ones(B, n, dtype=torch.bool)
    key_valid = torch.cat([st.key_valid, attn_mask.bool()], dim=1)            # (B, L)
    L = key_valid.shape[1]
    p = "transformer.model.decoder."
    if position_rule == "arange":            # transformers 5.5.0: arange(n) + past_len for every row
        pos = (torch.arange(n) + st.len)[None, :].expand(B, n)
    elif position_rule == "mask_cumsum":     # transformers 4.5x Whisper prepare_inputs_for_generation rule
        pos = (key_valid.long().cumsum(-1) - 1).clamp(min=0)[:, st.len:]
    else:
        raise ValueError(position_rule)
    x = w["decoder_embedder.weight"][ids] + w[p + "embed_positions.weight"][pos]
    qi = torch.arange(st.len, L)[:, None]
    kj = torch.arange(L)[None, :]
    allowed = (kj <= qi)[None, :, :] & key_valid[:, None, :]                  # (B, n, L)
    mask_add = torch.zeros(B, 1, n, L).masked_fill(~allowed[:, None], float("-inf"))
    scale = cfg.head_dim ** -0.5
    for i in range(cfg.decoder_layers):
        lp = f"{p}layers.{i}."
        h = _ln(x, w, lp + "self_attn_layer_norm.")
        q = F.linear(h, w[lp + "self_attn.q_proj.weight"], w[lp + "self_attn.q_proj.bias"]) * scale
        k = _split_heads(F.linear(h, w[lp + "self_attn.k_proj.weight"]), cfg.heads)
        v = _split_heads(F.linear(h, w[lp + "self_attn.v_proj.weight"], w[lp + "self_attn.v_proj.bias"]), cfg.heads)
        st.k[i] = k if st.k[i] is None else torch.cat([st.k[i], k], dim=2)
        st.v[i] = v if st.v[i] is None else torch.cat([st.v[i], v], dim=2)
        a = _attn(_split_heads(q, cfg.heads), st.k[i], st.v[i], mask_add)
        x = x + F.linear(a, w[lp + "self_attn.out_proj.weight"], w[lp + "self_attn.out_proj.bias"])
        h = _ln(x, w, lp + "encoder_attn_layer_norm.")
        q = F.linear(h, w[lp + "encoder_attn.q_proj.weight"], w[lp + "encoder_attn.q_proj.bias"]) * scale
        a = _attn(_split_heads(q, cfg.heads), st.ck[i], st.cv[i], None)
        x = x + F.linear(a, w[lp + "encoder_attn.out_proj.weight"], w[lp + "encoder_attn.out_proj.bias"])
        h = _ln(x, w, lp + "final_layer_norm.")
        h = F.gelu(F.linear(h, w[lp + "fc1.weight"], w[lp + "fc1.bias"]))
        x = x + F.linear(h, w[lp + "fc2.weight"], w[lp + "fc2.bias"])
    st.len = L
    st.key_valid = key_valid
    if last_only:
        x = x[:, -1:, :]
    x = _ln(x, w, p + "layer_norm.")
    return F.linear(x, w["transformer.proj_out.weight"]).float()


def forward_logits(w: W, cfg, pcm: torch.Tensor, ids: torch.Tensor, attn_mask: Optional[torch.Tensor],
                   position_rule: str = "arange") -> torch.Tensor:
    """Teacher-forced logits, `Mapperatorinator.forward` (modeling_mapperatorinator.py:139-228) without loss."""
    st = DecoderState(w, cfg, encode(w, cfg, pcm))
    return decoder_forward(st, ids, attn_mask, position_rule)
