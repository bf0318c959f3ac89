"""Static shapes of the hot path (the v29 chain) and the tiny plumbing config.

Mirrors the fields of the reference that the hot path reads:
  * `MapperatorinatorConfig` (osuT5/osuT5/model/configuration_mapperatorinator.py:7-177): n_fft/hop/n_mels,
    vocab sizes, max_source_positions = src_seq_len // 2 (:101), max_target_positions = tgt_seq_len.
  * HF `WhisperConfig` dims of openai/whisper-small (d_model 768, 12+12 layers, 12 heads, FFN 3072).
  * `DiT_models['DiT-B']` (osu_diffusion/utils/models.py:392-393): depth 12, hidden 768, 12 heads.
"""
from __future__ import annotations

import dataclasses


@dataclasses.dataclass(frozen=True)
class MelConfig:
    """spectrogram.py:7-61 constructor arguments (`n_ftt` sic)."""
    implementation: str = "nnAudio"   # "nnAudio" (v29: Slaney mel, norm=1, no log) | "torchaudio" (v30+: HTK, norm=None)
    log_scale: bool = False
    sample_rate: int = 16000
    n_fft: int = 1024
    n_mels: int = 388
    hop_length: int = 128
    f_min: float = 0.0
    f_max: float = 8000.0
    pad_mode: str = "constant"        # "constant" (v29) | "reflect" (v30+)


@dataclasses.dataclass(frozen=True)
class ModelConfig:
    """Dimensions of the osuT5 (Whisper backbone) model as the hot path sees them."""
    d_model: int = 768
    encoder_layers: int = 12
    decoder_layers: int = 12
    heads: int = 12
    ffn_dim: int = 3072
    src_seq_len: int = 1024           # mel frames per window
    tgt_seq_len: int = 2048           # max_target_positions
    vocab_size_out: int = 3667
    vocab_size_in: int = 4340
    mel: MelConfig = dataclasses.field(default_factory=MelConfig)

    @property
    def head_dim(self) -> int:
        return self.d_model // self.heads

    @property
    def max_source_positions(self) -> int:
        return self.src_seq_len // 2

    @property
    def max_target_positions(self) -> int:
        return self.tgt_seq_len

    @property
    def samples_per_window(self) -> int:
        # preprocessor.py:14-17: (src_seq_len - 1) * hop
        return (self.src_seq_len - 1) * self.mel.hop_length


@dataclasses.dataclass(frozen=True)
class DiTConfig:
    """osu_diffusion/utils/models.py:213-317 constructor arguments."""
    hidden: int = 768
    depth: int = 12
    heads: int = 12
    mlp_ratio: int = 4
    in_channels: int = 2
    context_size: int = 272
    class_size: int = 600
    pos_freq_dim: int = 128            # FirstLayer.frequency_embedding_size
    t_freq_dim: int = 256              # TimestepEmbedder.frequency_embedding_size

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads

    @property
    def out_channels(self) -> int:
        return self.in_channels * 2


def v29_model_config(vocab_size_in: int = 4340) -> ModelConfig:
    return ModelConfig(vocab_size_in=vocab_size_in)


def tiny_model_config(vocab_size_in: int = 4340, mel: MelConfig | None = None) -> ModelConfig:
    """BASELINE.json configs[0] (plumbing-size model). SURVEY §8d row 1 suggests 64-d / 4 heads; this build keeps
    head_dim = 64 (128-d / 2 heads / FFN 256, 2+2 layers) so the tiny case runs the SAME kernel instantiations as
    whisper-small instead of a second head_dim specialisation that production never uses."""
    return ModelConfig(d_model=128, encoder_layers=2, decoder_layers=2, heads=2, ffn_dim=256,
                       vocab_size_in=vocab_size_in, mel=mel or MelConfig())


def dit_b_config(class_size: int = 600) -> DiTConfig:
    return DiTConfig(class_size=class_size)


def tiny_dit_config(class_size: int = 40) -> DiTConfig:
    return DiTConfig(hidden=128, depth=2, heads=2, class_size=class_size)
