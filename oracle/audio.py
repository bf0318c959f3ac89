"""TEST INFRASTRUCTURE ONLY — CPU oracle of the audio ingest: PCM at the file's rate -> mono float32 at the model rate.

Reference: `load_audio_file` (osuT5/osuT5/dataset/data_utils.py:80-101, called by `Preprocessor.load`, preprocessor.py:39):
    audio = AudioSegment.from_file(file)                  # ffmpeg decode -> interleaved PCM            (OUT OF SCOPE: file decoding)
    audio.frame_rate = int(audio.frame_rate * speed)
    audio = audio.set_frame_rate(sample_rate)             # pydub -> audioop.ratecv(data, width, channels, in_rate, out_rate, None)
    audio = audio.set_channels(1)                         # pydub -> audioop.tomono(data, width, 0.5, 0.5) for stereo input
    samples = np.array(audio.get_array_of_samples()).astype(np.float32)
    return normalize_audio_samples(samples)               # samples / max|samples|  (data_utils.py:132-137)

pydub (requirements.txt: pydub==0.25.1) is NOT installed here and not vendored by the reference; its two methods are thin wrappers
(pydub/audio_segment.py `set_frame_rate`, `set_channels`) around the C functions `audioop.ratecv` / `audioop.tomono` of CPython's
standard library, which IS present (3.12).  `ingest_reference` calls those C functions directly — it is the arithmetic the reference runs.
`ingest_closed_form` restates them per output sample (what the CUDA kernel implements) and is held to `ingest_reference` bit for bit by
tests/test_audio_ingest.py, so the pin is "outputs of the reference's own dependency run here".

audioop.ratecv (Modules/audioop.c), 16-bit samples widened to 32 bits (s << 16), weights (1, 0), initial state d = -outrate:
    consume an input frame while d < 0 (prev <- cur, cur <- frame, d += outrate); emit while d >= 0:
        out = (int)(((double)prev * d + (double)cur * (outrate - d)) / (double)outrate) >> 16 ;  d -= inrate
(rates divided by their gcd first).  Output frame k therefore reads input frames n - 2 and n - 1 with n = 1 + ceil(k * inrate / outrate)
and d = (n - 1) * outrate - k * inrate; frames are emitted while n <= n_in.
audioop.tomono: floor(l * 0.5 + r * 0.5) after clamping to the sample range.
"""
import math
import warnings

import numpy as np


def out_frames(n_in: int, in_rate: int, out_rate: int) -> int:
    if in_rate == out_rate:
        return n_in
    g = math.gcd(in_rate, out_rate)
    i, o = in_rate // g, out_rate // g
    return 0 if n_in <= 0 else (n_in - 1) * o // i + 1


def ingest_reference(pcm: np.ndarray, frame_rate: int, sample_rate: int, speed: float = 1.0, normalize: bool = True) -> np.ndarray:
    """pcm: int16 [n_frames, channels] (channels 1 or 2).  The reference's arithmetic through CPython's audioop."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", DeprecationWarning)
        import audioop
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    n, ch = pcm.shape
    data = pcm.tobytes()
    in_rate = int(frame_rate * speed)
    if in_rate != sample_rate:                                       # pydub set_frame_rate: no-op when the rates agree
        data, _ = audioop.ratecv(data, 2, ch, in_rate, sample_rate, None)
    if ch == 2:                                                      # pydub set_channels(1)
        data = audioop.tomono(data, 2, 0.5, 0.5)
    samples = np.frombuffer(data, dtype=np.int16).astype(np.float32)
    if normalize:
        peak = np.max(np.abs(samples)) if samples.size > 0 else 0
        if peak > 0:
            samples = samples / peak
    return samples


def ingest_closed_form(pcm: np.ndarray, frame_rate: int, sample_rate: int, speed: float = 1.0, normalize: bool = True) -> np.ndarray:
    """The same result computed independently per output frame (the CUDA kernel's formulation)."""
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    n, ch = pcm.shape
    in_rate = int(frame_rate * speed)
    if in_rate != sample_rate:
        g = math.gcd(in_rate, sample_rate)
        i, o = in_rate // g, sample_rate // g
        K = out_frames(n, in_rate, sample_rate)
        k = np.arange(K, dtype=np.int64)
        nn = 1 + (k * i + o - 1) // o                                # frames consumed when output k is emitted
        d = ((nn - 1) * o - k * i).astype(np.float64)
        x = pcm.astype(np.int64) << 16
        cur = x[nn - 1].astype(np.float64)
        prev = np.where((nn >= 2)[:, None], x[np.maximum(nn - 2, 0)], 0).astype(np.float64)
        val = (prev * d[:, None] + cur * (o - d)[:, None]) / float(o)
        res = np.trunc(val).astype(np.int64) >> 16                   # (int) cast truncates toward zero, >> is arithmetic
    else:
        res = pcm.astype(np.int64)
    if ch == 2:
        v = res[:, 0].astype(np.float64) * 0.5 + res[:, 1].astype(np.float64) * 0.5
        v = np.where(v > 32767.0, 32767.0, np.where(v < -32767.0, -32768.0, v))          # audioop's fbound
        res = np.floor(v).astype(np.int64)
    else:
        res = res[:, 0]
    samples = res.astype(np.int16).astype(np.float32)
    if normalize:
        peak = np.max(np.abs(samples)) if samples.size > 0 else 0
        if peak > 0:
            samples = samples / peak
    return samples
