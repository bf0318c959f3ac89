"""Independent pins for the v29 (nnAudio 0.3.4) mel front end.  nnAudio itself is not installed and not vendored in the reference
tree (requirements.txt:3), so the oracle RESTATES its published algorithm (oracle/mel.py).  These tests anchor the two halves of
that restatement on independent implementations that ARE present:

  * the Slaney / area-normalised filterbank (nnAudio vendors librosa's `filters.mel(htk=False, norm=1)`) against torchaudio's own
    implementation of the same published definition, `melscale_fbanks(norm="slaney", mel_scale="slaney")`;
  * the windowed-DFT-as-conv1d STFT against an explicit float64 DFT (`X[k, t] = sum_n w[n] x[t*hop + n] e^{-2 pi i k n / N}` over the
    zero-padded signal, `center=True`), i.e. the textbook definition nnAudio's kernels encode, and against `torch.stft`.
Together with the torchaudio-flavour pin (tests/test_oracle_vs_reference.py, same STFT core, reference module imported) this
bounds what "parity unpinned vs nnAudio" can still hide to nnAudio-internal dtype choices (float32 kernels, `sqrt(.)**2`).
"""
import numpy as np
import pytest
import torch

from mapperatorinator_b200 import MelConfig
from mapperatorinator_b200.filterbank import mel_filterbank, slaney_filterbank
from oracle import mel as mel_oracle

V29 = MelConfig()      # nnAudio, 388 mels, 0-8000 Hz, n_fft 1024, hop 128, zero padding, no log


def test_slaney_filterbank_matches_torchaudio():
    import torchaudio
    ref = torchaudio.functional.melscale_fbanks(V29.n_fft // 2 + 1, V29.f_min, V29.f_max, V29.n_mels, V29.sample_rate,
                                                norm="slaney", mel_scale="slaney").T.numpy()          # (n_mels, n_freq)
    for name, fb in (("product", slaney_filterbank(V29)), ("oracle", mel_oracle.slaney_mel_basis(16000, 1024, 388, 0.0, 8000.0))):
        assert fb.shape == ref.shape == (388, 513)
        scale = np.abs(ref).max()
        # torchaudio evaluates the same formulas in float32 (librosa / nnAudio and this repo: float64, then cast): 2e-5 relative
        assert np.abs(fb - ref).max() <= 3e-5 * scale, (name, np.abs(fb - ref).max(), scale)
        assert np.abs(fb - ref)[(fb > 0) != (ref > 0)].size == 0 or np.abs(fb - ref)[(fb > 0) != (ref > 0)].max() <= 3e-5 * scale      # same support
    assert np.array_equal(mel_filterbank(V29), slaney_filterbank(V29))


def test_nnaudio_stft_matches_fp64_explicit_dft():
    g = torch.Generator().manual_seed(7)
    n = 130944
    t = torch.arange(n) / 16000.0
    x = (0.3 * torch.sin(2 * np.pi * 440 * t) + 0.1 * torch.sin(2 * np.pi * 3111 * t) + 0.05 * torch.randn(n, generator=g))[None]
    # oracle restatement: power spectrogram (B, n_freq, frames) before the filterbank
    wcos, wsin = mel_oracle.nnaudio_fourier_kernels(V29.n_fft)
    xp = torch.nn.functional.pad(x[:, None, :], (512, 512))
    re = torch.nn.functional.conv1d(xp, wcos, stride=128)
    im = torch.nn.functional.conv1d(xp, wsin, stride=128)
    power = (torch.sqrt(re.pow(2) + im.pow(2)) ** 2.0)[0].double()                                       # (513, 1024)
    assert power.shape == (513, 1024)
    # explicit float64 DFT of a sample of frames (all bins)
    w = torch.from_numpy(mel_oracle.hann_window(1024))
    xp64 = xp[0, 0].double()
    k = torch.arange(513, dtype=torch.float64)[:, None]
    nn_ = torch.arange(1024, dtype=torch.float64)[None, :]
    ang = 2 * np.pi * k * nn_ / 1024
    frames = [0, 1, 3, 4, 100, 511, 777, 1020, 1023]
    worst = 0.0
    for f in frames:
        seg = xp64[f * 128:f * 128 + 1024] * w
        ref = (torch.cos(ang) @ seg) ** 2 + (torch.sin(ang) @ seg) ** 2
        worst = max(worst, float((power[:, f] - ref).abs().max() / ref.abs().max()))
    assert worst <= 5e-6, worst                                                                           # fp32 conv vs fp64 DFT
    # and torch.stft (the core the torchaudio flavour is pinned through) gives the same spectrum
    st = torch.stft(x.double(), 1024, hop_length=128, window=w, center=True, pad_mode="constant", return_complex=True)[0]
    p2 = st.real ** 2 + st.imag ** 2
    assert float((power - p2).abs().max() / p2.abs().max()) <= 5e-6


def test_nnaudio_mel_forward_is_filterbank_times_fp64_power():
    """End to end: oracle `mel_forward` (v29 flavour) == torchaudio's Slaney filterbank applied to the float64 `torch.stft` power
    spectrum, to fp32 accuracy."""
    import torchaudio
    g = torch.Generator().manual_seed(8)
    x = torch.randn(2, 130944, generator=g) * 0.1
    out = mel_oracle.mel_forward(x, V29).double()                                                          # (B, frames, mels)
    w = torch.from_numpy(mel_oracle.hann_window(1024))
    st = torch.stft(x.double(), 1024, hop_length=128, window=w, center=True, pad_mode="constant", return_complex=True)
    fb = torchaudio.functional.melscale_fbanks(513, 0.0, 8000.0, 388, 16000, norm="slaney", mel_scale="slaney").double()   # (513, 388)
    ref = torch.matmul((st.real ** 2 + st.imag ** 2).transpose(1, 2), fb)
    assert out.shape == ref.shape == (2, 1024, 388)
    assert float((out - ref).abs().max() / ref.abs().max()) <= 5e-5, float((out - ref).abs().max() / ref.abs().max())
