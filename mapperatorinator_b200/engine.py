"""Python host of the engine: thin owners of the C-ABI handles.  torch is used for device memory and streams only."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .config import DiTConfig, MelConfig, ModelConfig
from .filterbank import mel_filterbank
from .token_layout import TokenLayout

VF_EOS, VF_TIMED, VF_SOS, VF_LB_EOS, VF_BEAT, VF_MANIA, VF_SCROLL = 1, 2, 4, 8, 16, 32, 64


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _mel_c(cfg: MelConfig) -> _lib.MelConfigC:
    return _lib.MelConfigC(cfg.n_fft, cfg.hop_length, cfg.n_mels, 1 if cfg.pad_mode == "reflect" else 0, 1 if cfg.log_scale else 0)


class MelEngine:
    """Stage (i).  `forward` == reference `MelSpectrogram.forward` (spectrogram.py:63-83)."""

    def __init__(self, cfg: MelConfig, mel_basis: Optional[np.ndarray] = None):
        self.cfg = cfg
        self.lib = _lib.load()
        basis = np.ascontiguousarray(mel_filterbank(cfg) if mel_basis is None else mel_basis, dtype=np.float32)
        assert basis.shape == (cfg.n_mels, cfg.n_fft // 2 + 1)
        self.handle = C.c_void_p()
        cc = _mel_c(cfg)
        _lib.check(self.lib.mb200_mel_create(C.byref(self.handle), C.byref(cc), basis.ctypes.data))

    def forward(self, samples: torch.Tensor) -> torch.Tensor:
        assert samples.is_cuda and samples.dtype == torch.float32 and samples.dim() == 2
        samples = samples.contiguous()
        B, n = samples.shape
        out = torch.empty(B, n // self.cfg.hop_length + 1, self.cfg.n_mels, device=samples.device, dtype=torch.float32)
        _lib.check(self.lib.mb200_mel_forward(self.handle, samples.data_ptr(), B, n, out.data_ptr(), _stream()))
        return out

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.mb200_mel_destroy(self.handle)
        except Exception:
            pass


def build_vflags(layout: TokenLayout, eos_ids: Sequence[int]) -> np.ndarray:
    """Per-token flag byte consumed by the fused decode step (kernels.h VF_*)."""
    f = np.zeros(layout.vocab_size_in, dtype=np.uint8)
    for ids, bit in ((eos_ids, VF_EOS), (layout.timed_token_ids(), VF_TIMED), (layout.sos_ids(), VF_SOS),
                     (layout.lookback_eos_ids(), VF_LB_EOS), (layout.beat_type_tokens(), VF_BEAT),
                     (layout.mania_type_tokens(), VF_MANIA), (layout.scroll_speed_tokens(), VF_SCROLL)):
        if len(ids):
            f[np.asarray(list(ids), dtype=np.int64)] |= bit
    return f


class ModelEngine:
    """Stage (ii): weights + resident encoder slots + KV arena behind `mb200_model_*`."""

    def __init__(self, cfg: ModelConfig, state_dict: Dict[str, torch.Tensor], max_windows: int = 32, max_batch: int = 16,
                 mel_basis: Optional[np.ndarray] = None, device: str = "cuda:0"):
        if not torch.cuda.is_available():
            raise RuntimeError("mapperatorinator_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.cfg = cfg
        self.device = torch.device(device)
        self.lib = _lib.load()
        self.max_windows, self.max_batch = max_windows, max_batch
        basis = np.ascontiguousarray(mel_filterbank(cfg.mel) if mel_basis is None else mel_basis, dtype=np.float32)
        cc = _lib.ModelConfigC(cfg.d_model, cfg.encoder_layers, cfg.decoder_layers, cfg.heads, cfg.ffn_dim, cfg.src_seq_len,
                               cfg.tgt_seq_len, cfg.vocab_size_in, cfg.vocab_size_out, _mel_c(cfg.mel), max_windows, max_batch)
        self.handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mb200_model_create(C.byref(self.handle), C.byref(cc), basis.ctypes.data))
            for name, t in state_dict.items():
                if not isinstance(t, torch.Tensor) or not t.is_floating_point():
                    continue
                a = t.detach().to("cpu", torch.float32).contiguous().numpy()
                _lib.check(self.lib.mb200_model_set_weight(self.handle, name.encode(), a.ctypes.data, a.size))
            _lib.check(self.lib.mb200_model_finalize(self.handle))

    # ---- encoder ---------------------------------------------------------------------------------------------------
    def encode(self, pcm: torch.Tensor, slot_begin: int = 0, return_states: bool = False) -> Optional[torch.Tensor]:
        """OsuTEncoder.forward + cross-K/V for `pcm` (n, samples_per_window) CUDA f32, into slots [slot_begin, +n)."""
        assert pcm.is_cuda and pcm.dtype == torch.float32 and pcm.dim() == 2 and pcm.shape[1] == self.cfg.samples_per_window
        pcm = pcm.contiguous()
        n = pcm.shape[0]
        out = None
        if return_states:
            out = torch.empty(n, self.cfg.max_source_positions, self.cfg.d_model, device=pcm.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mb200_model_encode(self.handle, pcm.data_ptr(), n, slot_begin, out.data_ptr() if out is not None else None,
                                                   _stream()))
        return out

    # ---- decoder ---------------------------------------------------------------------------------------------------
    def _generate_params(self, layout: TokenLayout, generate_kwargs: dict, position_rule: str = "arange"):
        """generate_kwargs (server.py:83-134 names) -> the C struct + the EOS id set."""
        gk = dict(generate_kwargs)
        t = float(gk.get("temperature", 1.0))
        types_first = bool(gk.get("types_first", False))
        lookback_time = float(gk.get("lookback_time", 0.0))
        lookahead_time = float(gk.get("lookahead_time", 0.0))
        ctx = gk.get("context_type")
        eos_ids = layout.eos_token_ids(lookback_time, lookahead_time, ctx)
        p = _lib.GenerateParamsC()
        p.cfg_scale = float(gk.get("cfg_scale", 1.0))
        p.timeshift_bias = float(gk.get("timeshift_bias", 0))
        p.types_first = int(types_first)
        p.temperature = t
        p.timing_temperature = float(gk.get("timing_temperature", t))
        p.mania_column_temperature = float(gk.get("mania_column_temperature", t))
        p.taiko_hit_temperature = float(gk.get("taiko_hit_temperature", t))
        conds = []
        if types_first:  # logit_processors.py:62-71
            if p.timing_temperature != t and layout.beat_type_tokens():
                conds.append((p.timing_temperature, 1, VF_BEAT))
            if p.mania_column_temperature != t and layout.mania_type_tokens():
                conds.append((p.mania_column_temperature, 3, VF_MANIA))
            if p.taiko_hit_temperature != t and layout.scroll_speed_tokens():
                conds.append((p.taiko_hit_temperature, 1, VF_SCROLL))
        p.n_cond = len(conds)
        for i, (ct, off, fl) in enumerate(conds):
            p.cond_temp[i], p.cond_offset[i], p.cond_flag[i] = ct, off, fl
        p.lookback_on = int(lookback_time > 0)
        p.lookback_start = layout.time_shift_start
        p.lookback_end = layout.lookback_end(lookback_time) if lookback_time > 0 else layout.time_shift_start
        p.do_sample = int(bool(gk.get("do_sample", False)))
        p.top_k = int(gk.get("top_k", 0) or 0)
        p.top_p = float(gk.get("top_p", 1.0) if gk.get("top_p") is not None else 1.0)
        p.top_p_cut = 1.0 - float(gk.get("top_p", 1.0) if gk.get("top_p") is not None else 1.0)      # double arithmetic, then one rounding to f32
        p.max_length = int(gk.get("max_length", self.cfg.tgt_seq_len))
        p.min_new_tokens = int(gk.get("min_new_tokens") or 0)
        pad = gk.get("pad_token_id", layout.pad_id)
        p.pad_token_id = int(layout.pad_id if pad is None else pad)
        # The reference draws from torch's global RNG (HF _sample -> torch.multinomial), i.e. a fresh stream per call that
        # `torch.manual_seed` controls.  Same contract here: without an explicit `seed` every call takes a new 62-bit seed from
        # torch's default generator (an explicit seed pins the device RNG for tests).
        seed = gk.get("seed")
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if p.do_sample else 0
        p.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        p.time_shift_start, p.time_shift_end = layout.time_shift_start, layout.time_shift_end
        p.position_rule = {"arange": 0, "mask_cumsum": 1}[position_rule]
        return p, eos_ids, gk

    def generate(self, slots: Sequence[int], prompt: torch.Tensor, prompt_mask: Optional[torch.Tensor], layout: TokenLayout,
                 generate_kwargs: dict, negative_prompt: Optional[torch.Tensor] = None,
                 negative_mask: Optional[torch.Tensor] = None, position_rule: str = "arange") -> torch.Tensor:
        """The token loop of `server.model_generate` for rows whose encoder states already sit in `slots`.
        Returns a CPU LongTensor (B, L) = prompt + generated, like the reference."""
        p, eos_ids, gk = self._generate_params(layout, generate_kwargs, position_rule)
        B, P = prompt.shape
        if int(gk.get("num_beams", 1) or 1) != 1:
            raise NotImplementedError("beam search is outside the hot path (SURVEY §8: greedy / sampling only)")
        use_cfg = negative_prompt is not None and p.cfg_scale > 1.0

        ids = np.ascontiguousarray(prompt.detach().cpu().numpy().astype(np.int64))
        msk = None if prompt_mask is None else np.ascontiguousarray(prompt_mask.detach().cpu().numpy().astype(np.uint8))
        neg = nmsk = None
        if use_cfg:
            neg_full = ids.copy()      # prepare_inputs_for_generation: ids.repeat(2); [:B, :neg_len] = negative prompt
            npn = negative_prompt.detach().cpu().numpy().astype(np.int64)
            neg_full[:, :npn.shape[1]] = npn
            neg = np.ascontiguousarray(neg_full)
            # the reference's negative_prompt_attention_mask is swallowed by HF generate()'s own parameter of that name
            # (transformers generation/utils.py:2142): the negative rows run with the conditional prompt's mask.
            nmsk = np.ascontiguousarray(msk.copy() if msk is not None else np.ones_like(ids, dtype=np.uint8))
        vflags = build_vflags(layout, eos_ids)
        slots_a = np.ascontiguousarray(np.asarray(list(slots), dtype=np.int32))
        assert slots_a.shape[0] == B
        out = np.zeros((B, p.max_length), dtype=np.int64)
        out_len = C.c_int32(0)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mb200_model_generate(
                self.handle, slots_a.ctypes.data, B, ids.ctypes.data, None if msk is None else msk.ctypes.data, P,
                None if neg is None else neg.ctypes.data, None if nmsk is None else nmsk.ctypes.data, vflags.ctypes.data,
                C.byref(p), out.ctypes.data, C.byref(out_len), _stream()))
        L = out_len.value
        return torch.from_numpy(out.reshape(-1)[: B * L].reshape(B, L).copy())

    def logits_chain(self, logits: torch.Tensor, ids: torch.Tensor, prompt_len: int, layout: TokenLayout, generate_kwargs: dict,
                     step: int = 0, has_last_scores: bool = False, use_cfg: bool = False):
        """Parity hook: one selection step of the fused logits-processor chain on given logits (rows, V) CUDA f32 and ids (B, L).
        Returns (scores (B, V) CUDA — what the selection sees, -inf = removed —, chosen (B,) CPU)."""
        p, eos_ids, gk = self._generate_params(layout, generate_kwargs)
        B, L = ids.shape
        a = np.ascontiguousarray(ids.detach().cpu().numpy().astype(np.int64))
        vflags = build_vflags(layout, eos_ids)
        logits = logits.contiguous().float()
        scores = torch.empty(B, self.cfg.vocab_size_out, device=logits.device, dtype=torch.float32)
        chosen = np.zeros(B, dtype=np.int64)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mb200_model_logits_chain(self.handle, logits.data_ptr(), B, int(use_cfg), a.ctypes.data, L, int(prompt_len),
                                                         vflags.ctypes.data, C.byref(p), int(step), int(has_last_scores), scores.data_ptr(),
                                                         chosen.ctypes.data, _stream()))
        return scores, torch.from_numpy(chosen)

    def forward_logits(self, slots: Sequence[int], ids: torch.Tensor, mask: Optional[torch.Tensor],
                       position_rule: str = "arange") -> torch.Tensor:
        B, L = ids.shape
        a = np.ascontiguousarray(ids.detach().cpu().numpy().astype(np.int64))
        m = None if mask is None else np.ascontiguousarray(mask.detach().cpu().numpy().astype(np.uint8))
        slots_a = np.ascontiguousarray(np.asarray(list(slots), dtype=np.int32))
        out = torch.empty(B, L, self.cfg.vocab_size_out, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mb200_model_forward_logits(self.handle, slots_a.ctypes.data, B, a.ctypes.data,
                                                           None if m is None else m.ctypes.data, L,
                                                           {"arange": 0, "mask_cumsum": 1}[position_rule], out.data_ptr(), _stream()))
        return out

    def set_option(self, name: str, value: int) -> None:
        _lib.check(self.lib.mb200_model_set_option(self.handle, name.encode(), int(value)))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.mb200_model_destroy(self.handle)
        except Exception:
            pass


class DiTEngine:
    """Stage (iii) behind `mb200_dit_*`."""

    def __init__(self, cfg: DiTConfig, state_dict: Dict[str, torch.Tensor], max_seq_len: int = 1024, max_batch: int = 2,
                 device: str = "cuda:0"):
        if not torch.cuda.is_available():
            raise RuntimeError("mapperatorinator_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.cfg = cfg
        self.device = torch.device(device)
        self.lib = _lib.load()
        cc = _lib.DitConfigC(cfg.hidden, cfg.depth, cfg.heads, cfg.mlp_ratio, cfg.in_channels, cfg.context_size, cfg.class_size,
                             cfg.pos_freq_dim, cfg.t_freq_dim, max_seq_len, max_batch)
        self.handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mb200_dit_create(C.byref(self.handle), C.byref(cc)))
            for name, t in state_dict.items():
                a = t.detach().to("cpu", torch.float32).contiguous().numpy()
                _lib.check(self.lib.mb200_dit_set_weight(self.handle, name.encode(), a.ctypes.data, a.size))
            _lib.check(self.lib.mb200_dit_finalize(self.handle))
        self._mask_keep = None

    def _mask(self, mask_mode: str, band: int, dense: Optional[torch.Tensor]) -> _lib.DitMaskC:
        mm = {"none": 0, "band": 2, "dense": 3}[mask_mode]
        self._mask_keep = dense
        return _lib.DitMaskC(mm, int(band), None if dense is None else dense.data_ptr())

    def forward_with_cfg(self, x, t, c, y, cfg_scale: float, mask_mode: str = "none", band: int = 0, dense=None) -> torch.Tensor:
        N, _, T = x.shape
        x, c, y = x.contiguous().float(), c.contiguous().float(), y.contiguous().float()
        tt = np.ascontiguousarray(t.detach().cpu().numpy().astype(np.int32))
        out = torch.empty(N, self.cfg.out_channels, T, device=x.device, dtype=torch.float32)
        mk = self._mask(mask_mode, band, dense)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mb200_dit_forward_with_cfg(self.handle, x.data_ptr(), tt.ctypes.data, c.data_ptr(), y.data_ptr(), N, T,
                                                           float(cfg_scale), C.byref(mk), out.data_ptr(), _stream()))
        return out

    def sample_loop(self, z, c, y, cfg_scale: float, schedule_rows: np.ndarray, noise: torch.Tensor,
                    inpaint: Optional[torch.Tensor] = None, mask_mode: str = "none", band: int = 0, dense=None) -> torch.Tensor:
        """schedule_rows (steps, 8) in LOOP order (first row = highest timestep); noise (steps, N, 2, T)."""
        N, _, T = z.shape
        z, c, y, noise = z.contiguous().float(), c.contiguous().float(), y.contiguous().float(), noise.contiguous().float()
        sched = np.ascontiguousarray(schedule_rows, dtype=np.float32)
        steps = sched.shape[0]
        assert noise.shape == (steps, N, 2, T)
        ip = None if inpaint is None else inpaint.to(torch.uint8).contiguous()
        out = torch.empty_like(z)
        mk = self._mask(mask_mode, band, dense)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mb200_dit_sample_loop(self.handle, z.data_ptr(), c.data_ptr(), y.data_ptr(),
                                                      None if ip is None else ip.data_ptr(), N, T, float(cfg_scale), C.byref(mk),
                                                      sched.ctypes.data, steps, noise.data_ptr(), out.data_ptr(), _stream()))
        return out

    CURVE_TYPES = {"Bezier": 0, "PerfectCurve": 1, "Catmull": 2, "Linear": 3, None: 0}

    def set_sliders(self, sliders) -> int:
        """Register the sliders of the chunk about to be sampled: iterable of (curve_type, control-point indices, end index, length)
        with CHUNK-RELATIVE indices (empty / None clears).  Returns how many were registered."""
        sl = list(sliders or [])
        if not sl:
            _lib.check(self.lib.mb200_dit_set_sliders(self.handle, 0, None, None, None, None, None))
            return 0
        off = np.zeros(len(sl) + 1, dtype=np.int32)
        off[1:] = np.cumsum([len(s[1]) for s in sl])
        idx = np.ascontiguousarray(np.concatenate([np.asarray(s[1], dtype=np.int32) for s in sl]))
        end = np.ascontiguousarray([int(s[2]) for s in sl], dtype=np.int32)
        typ = np.ascontiguousarray([self.CURVE_TYPES.get(s[0], 0) for s in sl], dtype=np.int32)      # anything unknown flattens as a bezier (slider_path.py:114-115)
        length = np.ascontiguousarray([float(s[3]) for s in sl], dtype=np.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mb200_dit_set_sliders(self.handle, len(sl), off.ctypes.data, idx.ctypes.data, end.ctypes.data, typ.ctypes.data,
                                                      length.ctypes.data))
        return len(sl)

    def apply_sliders(self, x: torch.Tensor) -> torch.Tensor:
        """The slider half of the `denoised_fn` closure on x (N, 2, T) CUDA f32 (returns a new tensor)."""
        x = x.contiguous().float().clone()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mb200_dit_apply_sliders(self.handle, x.data_ptr(), x.shape[0], x.shape[2], _stream()))
        return x

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.mb200_dit_destroy(self.handle)
        except Exception:
            pass
