"""Drop-in for the model-call boundary of the reference: `osuT5.osuT5.inference.server.model_generate` / `model_forward`
(osuT5/osuT5/inference/server.py:72-181).  Same signature, same returns (CPU LongTensor of prompt+generated ids, stats dict
with the reference's token accounting, :50-69), so `Processor.model_generate` (processor.py:155-176) can call it unchanged
with a `B200Mapperatorinator` as `model`.
"""
from __future__ import annotations

import time

import torch

from .token_layout import TokenLayout


def get_eos_token_id(tokenizer, lookback_time: float = 0, lookahead_time: float = 0, context_type=None):
    """server.py:72-80."""
    return TokenLayout.from_tokenizer(tokenizer).eos_token_ids(lookback_time, lookahead_time, context_type)


def _build_generation_stats(result: torch.Tensor, model_kwargs: dict, pad_token_id, elapsed_seconds: float) -> dict:
    """server.py:50-69: generated = non-pad output tokens minus non-pad prompt tokens, per row, clamped at 0."""
    mask = model_kwargs.get("decoder_attention_mask")
    ids = model_kwargs.get("decoder_input_ids")
    if isinstance(mask, torch.Tensor):
        prompt_counts = mask.to(torch.long).sum(dim=-1).cpu()
    elif pad_token_id is None:
        prompt_counts = torch.full((ids.shape[0],), ids.shape[1], dtype=torch.long)
    else:
        prompt_counts = ids.ne(pad_token_id).to(torch.long).sum(dim=-1).cpu()
    if pad_token_id is None:
        out_counts = torch.full((result.shape[0],), result.shape[1], dtype=torch.long)
    else:
        out_counts = result.ne(pad_token_id).to(torch.long).sum(dim=-1)
    gen = torch.clamp(out_counts - prompt_counts, min=0)
    n = int(gen.sum().item())
    return {"generated_tokens": n, "generated_tokens_per_sample": gen.tolist(), "elapsed_seconds": float(elapsed_seconds),
            "tokens_per_second": n / elapsed_seconds if elapsed_seconds > 0 else 0.0}


@torch.no_grad()
def model_generate(model, tokenizer, model_kwargs, generate_kwargs):
    """`model_generate(model, tokenizer, model_kwargs, generate_kwargs) -> (LongTensor[B, P+N] on CPU, stats)`.
    `precision` is accepted for signature parity; the engine computes in fp32 (the parity contract of the hot path)."""
    generate_kwargs = dict(generate_kwargs)
    generate_kwargs.pop("precision", None)
    layout = TokenLayout.from_tokenizer(tokenizer)
    pad_token_id = generate_kwargs.get("pad_token_id", getattr(tokenizer, "pad_id", None))
    start = time.perf_counter()
    result = model.generate(
        inputs=model_kwargs["inputs"], decoder_input_ids=model_kwargs["decoder_input_ids"],
        decoder_attention_mask=model_kwargs.get("decoder_attention_mask"), negative_prompt=model_kwargs.get("negative_prompt"),
        negative_prompt_attention_mask=model_kwargs.get("negative_prompt_attention_mask"), tokenizer=layout,
        generate_kwargs=generate_kwargs)
    elapsed = time.perf_counter() - start
    result = result.cpu()
    return result, _build_generation_stats(result, model_kwargs, pad_token_id, elapsed)


@torch.no_grad()
def model_forward(model, model_kwargs, generate_kwargs):
    """server.py:159-181 (cfg_scale == 1 path): teacher-forced fp32 logits on the CPU."""
    out = model.forward(frames=model_kwargs["inputs"], decoder_input_ids=model_kwargs["decoder_input_ids"],
                        decoder_attention_mask=model_kwargs.get("decoder_attention_mask"))
    return out.logits.to(torch.float32).cpu()
