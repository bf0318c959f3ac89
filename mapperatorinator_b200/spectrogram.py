"""Drop-in for `osuT5.osuT5.model.spectrogram.MelSpectrogram` (same constructor arguments, same forward contract)."""
from __future__ import annotations

import torch

from .config import MelConfig
from .engine import MelEngine


class MelSpectrogram:
    """(B, n_samples) f32 -> (B, n_samples // hop + 1, n_mels) f32; always fp32 like the reference (`_apply`, :85-92)."""

    def __init__(self, implementation: str = "nnAudio", log_scale: bool = False, sample_rate: int = 16000, n_ftt: int = 2048,
                 n_mels: int = 512, hop_length: int = 128, f_min: int = 0, f_max: int = 8000, pad_mode: str = "constant",
                 mel_basis=None):
        assert implementation in ["torchaudio", "nnAudio"], f"Unsupported implementation: {implementation}"
        self.cfg = MelConfig(implementation, log_scale, sample_rate, n_ftt, n_mels, hop_length, f_min, f_max, pad_mode)
        self.log_scale = log_scale
        self._engine = MelEngine(self.cfg, mel_basis)

    def forward(self, samples: torch.Tensor) -> torch.Tensor:
        return self._engine.forward(samples.to("cuda", torch.float32))

    __call__ = forward
