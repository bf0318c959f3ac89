"""Audio ingest: the oracle's per-sample restatement vs CPython's audioop (the C code pydub calls for the reference's
`load_audio_file`, data_utils.py:80-101), and the CUDA kernel vs both — bit-exact (integer / byte arithmetic)."""
import numpy as np
import pytest
import torch

from oracle import audio as ao

CASES = [  # (frames, channels, file rate, model rate, speed, normalize)
    (10007, 2, 44100, 16000, 1.0, True), (5000, 1, 44100, 16000, 1.0, True), (4096, 2, 48000, 16000, 1.0, True),
    (3000, 2, 44100, 16000, 1.5, True), (1, 2, 44100, 16000, 1.0, True), (2, 1, 44100, 16000, 1.0, False),
    (777, 2, 16000, 16000, 1.0, True), (2000, 2, 22050, 16000, 1.0, False), (999, 1, 8000, 16000, 1.0, True),
]


def _pcm(n, ch, seed):
    rng = np.random.default_rng(seed)
    x = rng.integers(-32768, 32768, size=(n, ch), dtype=np.int64).astype(np.int16)
    if n > 10:                       # extremes next to each other: the (int) truncation, the >> 16 floor and tomono's clamp all get exercised
        x[3] = -32768; x[4] = 32767; x[5] = -32768; x[6, 0] = 32767; x[7] = -1
    return x


@pytest.mark.parametrize("case", CASES)
def test_closed_form_equals_audioop(case):
    n, ch, ir, orate, sp, norm = case
    pcm = _pcm(n, ch, n)
    ref = ao.ingest_reference(pcm, ir, orate, sp, norm)
    got = ao.ingest_closed_form(pcm, ir, orate, sp, norm)
    assert ref.shape == (ao.out_frames(n, int(ir * sp), orate),)
    assert np.array_equal(ref, got)


def test_silence_and_empty():
    z = np.zeros((500, 2), dtype=np.int16)
    assert np.array_equal(ao.ingest_reference(z, 44100, 16000), ao.ingest_closed_form(z, 44100, 16000))       # peak 0: no division
    assert ao.out_frames(0, 44100, 16000) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_gpu_ingest_bit_exact(case):
    from mapperatorinator_b200.audio import load_pcm
    n, ch, ir, orate, sp, norm = case
    pcm = _pcm(n, ch, n + 1)
    ref = ao.ingest_reference(pcm, ir, orate, sp, norm)
    got = load_pcm(pcm, ir, orate, sp, norm).cpu().numpy()
    assert got.shape == ref.shape
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), f"first diff at {np.nonzero(got != ref)[0][:5]}"


@pytest.mark.gpu
def test_gpu_ingest_full_song_and_mono_1d():
    """BASELINE's workload: 180 s of 44.1 kHz stereo (7 938 000 frames -> 2 880 000 samples), bit-exact against audioop; a 1-D mono array;
    silence (peak 0 leaves the samples untouched)."""
    from mapperatorinator_b200.audio import load_pcm
    rng = np.random.default_rng(7)
    n = 180 * 44100
    t = np.arange(n) / 44100.0
    sig = (0.4 * np.sin(2 * np.pi * 220 * t) + 0.1 * rng.standard_normal(n))
    pcm = np.stack([np.clip(sig * 20000, -32768, 32767), np.clip(sig * 15000 + 300, -32768, 32767)], 1).astype(np.int16)
    ref = ao.ingest_reference(pcm, 44100, 16000)
    got = load_pcm(torch.from_numpy(pcm), 44100, 16000).cpu().numpy()
    assert got.shape == (2880000,) and np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    mono = pcm[:50000, 0].copy()
    assert np.array_equal(load_pcm(mono, 44100, 16000).cpu().numpy(), ao.ingest_reference(mono[:, None], 44100, 16000))
    z = np.zeros((4000, 2), dtype=np.int16)
    assert np.array_equal(load_pcm(z, 44100, 16000).cpu().numpy(), ao.ingest_reference(z, 44100, 16000))
