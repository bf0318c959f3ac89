"""Audio ingest on the GPU: the drop-in for the CPU tail of `load_audio_file`
(osuT5/osuT5/dataset/data_utils.py:80-101; `Preprocessor.load`, osuT5/osuT5/inference/preprocessor.py:39).

    audio = AudioSegment.from_file(path)                                  # ffmpeg decode: stays with the caller
    samples = load_pcm(np.array(audio.get_array_of_samples()).reshape(-1, audio.channels), audio.frame_rate, sample_rate, speed, normalize)

does what `audio.set_frame_rate(sample_rate).set_channels(1)` + float32 + `normalize_audio_samples` do, bit for bit, in two small
kernels (`csrc/audio.cu`).  No CPU fallback: raises without the built library or without a GPU."""
from typing import Union

import numpy as np
import torch

from . import _lib


def load_pcm(pcm: Union[np.ndarray, torch.Tensor], frame_rate: int, sample_rate: int, speed: float = 1.0, normalize: bool = True,
             device: str = "cuda") -> torch.Tensor:
    """pcm: int16 [n_frames, channels] (channels 1 or 2; a 1-D array is mono), host or device.  Returns float32 [n_out] on `device`."""
    lib = _lib.load()
    if not torch.cuda.is_available():
        raise RuntimeError("mapperatorinator_b200 needs a CUDA device (there is no CPU fallback)")
    t = torch.as_tensor(pcm)
    if t.dtype != torch.int16:
        raise TypeError("pcm must be int16 (AudioSegment.get_array_of_samples() of a 16-bit file)")
    if t.dim() == 1:
        t = t[:, None]
    t = t.contiguous().to(device, non_blocking=True)
    n, ch = t.shape
    in_rate = int(frame_rate * speed)                                     # data_utils.py:96
    n_out = int(lib.mb200_audio_out_frames(n, in_rate, int(sample_rate)))
    out = torch.empty(n_out, dtype=torch.float32, device=t.device)
    scratch = torch.zeros(1, dtype=torch.int32, device=t.device)
    if n_out > 0:
        _lib.check(lib.mb200_audio_ingest(t.data_ptr(), n, ch, in_rate, int(sample_rate), 1 if normalize else 0, out.data_ptr(), scratch.data_ptr(),
                                          torch.cuda.current_stream(t.device).cuda_stream))
    return out
