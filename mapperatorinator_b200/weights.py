"""Weight containers for the hot path: reference `state_dict()` key layout.

The engine ingests weights by their reference names, so a real checkpoint
(`Mapperatorinator.state_dict()`, `DiT.state_dict()`) loads unchanged.  There is no network / checkpoint in this
environment, so benchmarks and tests use `init_model_state_dict` / `init_dit_state_dict`: seeded random weights of
the exact reference shapes (CPU `torch.Generator`, reproducible on every box with the same torch build).
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from .config import DiTConfig, ModelConfig


def _randn(g: torch.Generator, *shape: int, std: float = 1.0) -> torch.Tensor:
    return torch.randn(*shape, generator=g, dtype=torch.float32) * std


def whisper_sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> torch.Tensor:
    """Frozen encoder positions (HF modeling_whisper.py `sinusoids`): cat[sin, cos] over log-spaced timescales."""
    log_inc = math.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-log_inc * torch.arange(channels // 2, dtype=torch.float32))
    t = torch.arange(length, dtype=torch.float32)[:, None] * inv[None, :]
    return torch.cat([t.sin(), t.cos()], dim=1)


def model_weight_shapes(cfg: ModelConfig) -> Dict[str, tuple]:
    """Every tensor of `Mapperatorinator.state_dict()` that the inference hot path reads (v29: no cond embedders)."""
    d, f = cfg.d_model, cfg.ffn_dim
    s: Dict[str, tuple] = {
        "encoder_embedder.weight": (d, cfg.mel.n_mels), "encoder_embedder.bias": (d,),
        "decoder_embedder.weight": (cfg.vocab_size_in, d),
        "transformer.model.encoder.conv1.weight": (d, d, 3), "transformer.model.encoder.conv1.bias": (d,),
        "transformer.model.encoder.conv2.weight": (d, d, 3), "transformer.model.encoder.conv2.bias": (d,),
        "transformer.model.encoder.embed_positions.weight": (cfg.max_source_positions, d),
        "transformer.model.encoder.layer_norm.weight": (d,), "transformer.model.encoder.layer_norm.bias": (d,),
        "transformer.model.decoder.embed_positions.weight": (cfg.max_target_positions, d),
        "transformer.model.decoder.layer_norm.weight": (d,), "transformer.model.decoder.layer_norm.bias": (d,),
        "transformer.proj_out.weight": (cfg.vocab_size_out, d),
    }

    def attn(prefix: str):
        s[prefix + "q_proj.weight"] = (d, d); s[prefix + "q_proj.bias"] = (d,)
        s[prefix + "k_proj.weight"] = (d, d)                                   # k has no bias (modeling_whisper.py:279)
        s[prefix + "v_proj.weight"] = (d, d); s[prefix + "v_proj.bias"] = (d,)
        s[prefix + "out_proj.weight"] = (d, d); s[prefix + "out_proj.bias"] = (d,)

    def ln(prefix: str):
        s[prefix + "weight"] = (d,); s[prefix + "bias"] = (d,)

    for side, n in (("encoder", cfg.encoder_layers), ("decoder", cfg.decoder_layers)):
        for i in range(n):
            p = f"transformer.model.{side}.layers.{i}."
            attn(p + "self_attn."); ln(p + "self_attn_layer_norm.")
            if side == "decoder":
                attn(p + "encoder_attn."); ln(p + "encoder_attn_layer_norm.")
            s[p + "fc1.weight"] = (f, d); s[p + "fc1.bias"] = (f,)
            s[p + "fc2.weight"] = (d, f); s[p + "fc2.bias"] = (d,)
            ln(p + "final_layer_norm.")
    return s


def init_model_state_dict(cfg: ModelConfig, seed: int = 0, std: float = 0.02) -> Dict[str, torch.Tensor]:
    """Seeded random weights with the reference key layout.  LayerNorm affine and biases are perturbed away from
    (1, 0) so that a kernel which drops them fails parity."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for name, shape in model_weight_shapes(cfg).items():
        if name.endswith("encoder.embed_positions.weight"):
            out[name] = whisper_sinusoids(*shape)
        elif "layer_norm.weight" in name:
            out[name] = 1.0 + _randn(g, *shape, std=0.1)
        elif name.endswith(".bias"):
            out[name] = _randn(g, *shape, std=0.02)
        elif name == "decoder_embedder.weight":
            out[name] = _randn(g, *shape, std=1.0)     # strong token identity -> varied greedy streams on random weights
        elif name.endswith("decoder.embed_positions.weight"):
            out[name] = _randn(g, *shape, std=0.3)
        elif name == "encoder_embedder.weight":
            out[name] = _randn(g, *shape, std=0.05)
        elif name == "transformer.proj_out.weight":
            out[name] = _randn(g, *shape, std=0.3)
        else:
            fan_in = shape[1] * (shape[2] if len(shape) == 3 else 1)
            out[name] = _randn(g, *shape, std=1.0 / math.sqrt(fan_in))
    return out


def dit_weight_shapes(cfg: DiTConfig) -> Dict[str, tuple]:
    """`DiT.state_dict()` (osu_diffusion/utils/models.py:213-279)."""
    d = cfg.hidden
    f = d * cfg.mlp_ratio
    s: Dict[str, tuple] = {
        "context_embedder.mlp.0.weight": (d, cfg.in_channels * cfg.pos_freq_dim + cfg.context_size),
        "context_embedder.mlp.0.bias": (d,),
        "t_embedder.mlp.0.weight": (d, cfg.t_freq_dim), "t_embedder.mlp.0.bias": (d,),
        "t_embedder.mlp.2.weight": (d, d), "t_embedder.mlp.2.bias": (d,),
        "y_embedder.class_embedding.0.weight": (d, cfg.class_size), "y_embedder.class_embedding.0.bias": (d,),
        "y_embedder.class_embedding.2.weight": (d, d), "y_embedder.class_embedding.2.bias": (d,),
        "final_layer.adaLN_modulation.1.weight": (2 * d, d), "final_layer.adaLN_modulation.1.bias": (2 * d,),
        "final_layer.linear.weight": (cfg.out_channels, d), "final_layer.linear.bias": (cfg.out_channels,),
    }
    for i in range(cfg.depth):
        p = f"blocks.{i}."
        s[p + "attn.in_proj_weight"] = (3 * d, d); s[p + "attn.in_proj_bias"] = (3 * d,)
        s[p + "attn.out_proj.weight"] = (d, d); s[p + "attn.out_proj.bias"] = (d,)
        s[p + "mlp.fc1.weight"] = (f, d); s[p + "mlp.fc1.bias"] = (f,)
        s[p + "mlp.fc2.weight"] = (d, f); s[p + "mlp.fc2.bias"] = (d,)
        s[p + "adaLN_modulation.1.weight"] = (6 * d, d); s[p + "adaLN_modulation.1.bias"] = (6 * d,)
    return s


def init_dit_state_dict(cfg: DiTConfig, seed: int = 1) -> Dict[str, torch.Tensor]:
    """Seeded DiT weights.  The reference zero-initialises adaLN and the output layer (models.py:270-279), which
    makes an untrained model output exactly 0; like SURVEY §8c.2 we re-randomise them so parity is meaningful."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for name, shape in dit_weight_shapes(cfg).items():
        if name.endswith(".bias") or name.endswith("in_proj_bias"):
            out[name] = _randn(g, *shape, std=0.02)
        elif "adaLN_modulation" in name:
            out[name] = _randn(g, *shape, std=0.02)
        elif name == "final_layer.linear.weight":
            out[name] = _randn(g, *shape, std=0.02)
        else:
            out[name] = _randn(g, *shape, std=1.0 / math.sqrt(shape[1]))
    return out
