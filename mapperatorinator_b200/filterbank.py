"""Host-side mel filterbanks for the two front ends of the reference `MelSpectrogram`
(osuT5/osuT5/model/spectrogram.py:38-61).  A real checkpoint carries these as buffers
(`spectrogram.transform.mel_basis` / `...mel_scale.fb`); without one they are regenerated here.
Parameter preparation only — the transform itself runs in csrc/mel.cu.
"""
from __future__ import annotations

import math

import numpy as np

from .config import MelConfig


def _slaney_hz_to_mel(f: np.ndarray) -> np.ndarray:
    f = np.asarray(f, dtype=np.float64)
    lin = f / (200.0 / 3)
    log_region = f >= 1000.0
    out = lin.copy()
    out[log_region] = 15.0 + np.log(f[log_region] / 1000.0) / (math.log(6.4) / 27.0)
    return out


def _slaney_mel_to_hz(m: np.ndarray) -> np.ndarray:
    m = np.asarray(m, dtype=np.float64)
    out = (200.0 / 3) * m
    log_region = m >= 15.0
    out[log_region] = 1000.0 * np.exp((math.log(6.4) / 27.0) * (m[log_region] - 15.0))
    return out


def slaney_filterbank(cfg: MelConfig) -> np.ndarray:
    """nnAudio / librosa `mel(htk=False, norm=1)`: (n_mels, n_fft//2+1) float32, area-normalised Slaney triangles."""
    n_freq = cfg.n_fft // 2 + 1
    bins = np.linspace(0.0, cfg.sample_rate / 2.0, n_freq)
    edges = _slaney_mel_to_hz(np.linspace(_slaney_hz_to_mel(np.array([cfg.f_min]))[0],
                                          _slaney_hz_to_mel(np.array([cfg.f_max]))[0], cfg.n_mels + 2))
    width = np.diff(edges)
    fb = np.zeros((cfg.n_mels, n_freq), dtype=np.float32)
    for i in range(cfg.n_mels):
        rise = (bins - edges[i]) / width[i]
        fall = (edges[i + 2] - bins) / width[i + 1]
        fb[i] = np.maximum(0.0, np.minimum(rise, fall))
    fb *= (2.0 / (edges[2:] - edges[:-2]))[:, None]
    return fb


def htk_filterbank(cfg: MelConfig) -> np.ndarray:
    """torchaudio `melscale_fbanks(mel_scale='htk', norm=None)` transposed to (n_mels, n_freq), float32 arithmetic."""
    import torch
    n_freq = cfg.n_fft // 2 + 1
    freqs = torch.linspace(0, cfg.sample_rate // 2, n_freq)
    m_lo = 2595.0 * math.log10(1.0 + cfg.f_min / 700.0)
    m_hi = 2595.0 * math.log10(1.0 + cfg.f_max / 700.0)
    pts = 700.0 * (10 ** (torch.linspace(m_lo, m_hi, cfg.n_mels + 2) / 2595.0) - 1.0)
    diff = pts[1:] - pts[:-1]
    slopes = pts.unsqueeze(0) - freqs.unsqueeze(1)
    fb = torch.clamp(torch.min(-slopes[:, :-2] / diff[:-1], slopes[:, 2:] / diff[1:]), min=0.0)
    return fb.T.contiguous().numpy().astype(np.float32)


def mel_filterbank(cfg: MelConfig) -> np.ndarray:
    return slaney_filterbank(cfg) if cfg.implementation == "nnAudio" else htk_filterbank(cfg)
