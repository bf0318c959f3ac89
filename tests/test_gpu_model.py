"""Stage-level parity through the reference-facing boundary: CUDA engine vs the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tiny():
    from mapperatorinator_b200 import tiny_model_config
    from mapperatorinator_b200.modeling import B200Mapperatorinator
    from mapperatorinator_b200.weights import init_model_state_dict
    cfg = tiny_model_config()
    sd = init_model_state_dict(cfg, 0)
    return cfg, sd, B200Mapperatorinator(cfg, sd, max_windows=8, max_batch=8)


def _pcm(cfg, B, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, cfg.samples_per_window, generator=g) * 0.1


def test_encoder_states(tiny):
    from oracle import whisper as wo
    cfg, sd, model = tiny
    pcm = _pcm(cfg, 3)
    ref = wo.encode(sd, cfg, pcm)
    out = model.engine.encode(pcm.cuda(), 0, return_states=True).cpu()
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-4), (out - ref).abs().max()


def test_teacher_forced_logits_left_padded(tiny):
    from oracle import whisper as wo
    cfg, sd, model = tiny
    pcm = _pcm(cfg, 2, 1)
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(17, cfg.vocab_size_in, (2, 21), generator=g)
    ids[1, :4] = 0
    mask = ids.ne(0)
    ref = wo.forward_logits(sd, cfg, pcm, ids, mask)
    out = model.forward(frames=pcm, decoder_input_ids=ids, decoder_attention_mask=mask).logits.cpu()
    real = mask[:, :, None].expand_as(ref)
    assert torch.allclose(out[real], ref[real], rtol=2e-4, atol=2e-4), (out[real] - ref[real]).abs().max()


GK = dict(precision="fp32", do_sample=False, num_beams=1, top_p=0.9, top_k=0, cfg_scale=1.0, timeshift_bias=0, types_first=True,
          temperature=0.9, timing_temperature=0.1, mania_column_temperature=0.5, taiko_hit_temperature=0.5)


@pytest.mark.parametrize("case", ["b1_first_window", "b2_leftpad_lookback", "b1_eos_stop", "b3_timeshift_bias"])
def test_greedy_generate_bit_exact(tiny, layout, case):
    from mapperatorinator_b200.server import model_generate
    from oracle import generate as go
    cfg, sd, model = tiny
    if case == "b1_first_window":
        prompt = torch.tensor([[3700, 3705, 3720, 1, 9]])
        gk = dict(GK, max_length=5 + 40, min_new_tokens=40, lookback_time=0.0, lookahead_time=3273.6, context_type="map")
    elif case == "b2_leftpad_lookback":
        prompt = torch.tensor([[0, 0, 3700, 3705, 1, 9, 3645, 30], [3700, 3701, 3702, 3703, 3704, 1, 9, 3655]])
        gk = dict(GK, max_length=8 + 48, min_new_tokens=48, lookback_time=4092.0, lookahead_time=3273.6, context_type="map")
    elif case == "b1_eos_stop":
        prompt = torch.tensor([[3700, 3705, 1, 9, 3645, 30]])
        gk = dict(GK, max_length=64, lookback_time=4092.0, lookahead_time=3273.6, context_type="map")     # natural EOS / max_length
    else:
        prompt = torch.tensor([[3700, 1, 5, 3657, 100], [3701, 1, 5, 3656, 90], [3702, 1, 5, 3655, 10]])
        gk = dict(GK, max_length=5 + 32, min_new_tokens=32, timeshift_bias=0.7, lookback_time=0.0, lookahead_time=0.0, context_type="timing")
    B = prompt.shape[0]
    pcm = _pcm(cfg, B, seed=B)
    mk = dict(inputs=pcm, decoder_input_ids=prompt, decoder_attention_mask=prompt.ne(0), negative_prompt=None,
              negative_prompt_attention_mask=None)
    want, wstats = go.model_generate(sd, cfg, layout, dict(mk), dict(gk))
    got, gstats = model_generate(model, layout, dict(mk), dict(gk))
    assert got.shape == want.shape, (got.shape, want.shape)
    if not torch.equal(got, want):
        diff = (got != want).nonzero()[0].tolist()
        pytest.fail(f"first divergence at row/col {diff}: got {got[diff[0], diff[1]].item()} want {want[diff[0], diff[1]].item()}")
    assert gstats["generated_tokens_per_sample"] == wstats["generated_tokens_per_sample"]


def test_cfg_generate_bit_exact(tiny, layout):
    from mapperatorinator_b200.server import model_generate
    from oracle import generate as go
    cfg, sd, model = tiny
    prompt = torch.tensor([[3700, 3705, 3710, 1, 9, 3645, 30], [3701, 3706, 3711, 1, 9, 3648, 55]])
    neg = torch.tensor([[0, 3700, 3712, 1, 9, 3645, 30], [0, 3701, 3713, 1, 9, 3648, 55]])
    gk = dict(GK, cfg_scale=2.0, max_length=7 + 32, lookback_time=0.0, lookahead_time=0.0, context_type="map")
    pcm = _pcm(cfg, 2, seed=7)
    mk = dict(inputs=pcm, decoder_input_ids=prompt, decoder_attention_mask=prompt.ne(0), negative_prompt=neg,
              negative_prompt_attention_mask=neg.ne(0))
    want, _ = go.model_generate(sd, cfg, layout, dict(mk), dict(gk))
    got, _ = model_generate(model, layout, dict(mk), dict(gk))
    assert torch.equal(got, want), (got.tolist(), want.tolist())


def test_sampling_is_valid_and_seeded(tiny, layout):
    from mapperatorinator_b200.server import model_generate
    cfg, sd, model = tiny
    prompt = torch.tensor([[3700, 3705, 1, 9]])
    gk = dict(GK, do_sample=True, top_p=0.9, max_length=4 + 32, min_new_tokens=32, lookback_time=0.0, lookahead_time=0.0,
              context_type="map", seed=5)
    mk = dict(inputs=_pcm(cfg, 1), decoder_input_ids=prompt, decoder_attention_mask=prompt.ne(0))
    a, _ = model_generate(model, layout, dict(mk), dict(gk))
    b, _ = model_generate(model, layout, dict(mk), dict(gk))
    c, _ = model_generate(model, layout, dict(mk), dict(gk, seed=6))
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert int(a[:, 4:].max()) < cfg.vocab_size_out and a.shape == (1, 36)


# ---- DiT --------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def tiny_dit():
    from mapperatorinator_b200 import tiny_dit_config
    from mapperatorinator_b200.diffusion import B200DiT
    from mapperatorinator_b200.weights import init_dit_state_dict
    dc = tiny_dit_config()
    sd = init_dit_state_dict(dc, 1)
    return dc, sd, B200DiT(dc, sd, max_seq_len=512)


def _dit_inputs(dc, T, seed=2):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(1, 2, T, generator=g) * 2 - 1
    c = torch.randn(1, dc.context_size, T, generator=g)
    y = (torch.rand(2, dc.class_size, generator=g) < 0.1).float()
    return torch.cat([x, x]), torch.cat([c, c]), y, g


@pytest.mark.parametrize("T,mask", [(300, "band"), (100, "none"), (130, "dense")])
def test_dit_forward_with_cfg(tiny_dit, T, mask):
    from oracle import dit as do
    dc, sd, dit = tiny_dit
    x, c, y, g = _dit_inputs(dc, T)
    am = {"band": do.band_mask(T, 128), "none": None, "dense": torch.rand(T, T, generator=g) < 0.2}[mask]
    if mask == "dense":
        am[torch.arange(T), torch.arange(T)] = False
    t = torch.tensor([37, 37])
    ref = do.dit_forward_with_cfg(sd, dc, x, t, c, y, 1.5, am)
    out = dit.forward_with_cfg(x.cuda(), t, c.cuda(), y.cuda(), 1.5, attn_mask=None if am is None else am.cuda()).cpu()
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-4), (out - ref).abs().max()


def test_dit_sample_loop_fused_and_python_seams(tiny_dit):
    from mapperatorinator_b200.diffusion import InpaintDenoiser, create_diffusion
    from oracle import dit as do
    dc, sd, dit = tiny_dit
    T = 200
    x, c, y, g = _dit_inputs(dc, T, seed=4)
    diff = create_diffusion([100, 0, 0, 0, 0, 0, 0, 0, 0, 0], "squaredcos_cap_v2", 1000)
    sched = do.Schedule()
    assert np.allclose(diff.schedule_rows()[::-1, 1:7], sched.table()[:, 1:7])
    noise = torch.randn(100, 2, 2, T, generator=g)
    ip = torch.ones_like(x, dtype=torch.bool)
    ip[:, :, :40] = False
    am = do.band_mask(T, 128)
    ref = do.p_sample_loop(sd, dc, sched, x, c, y, 1.0, am, noise, inpaint_mask=ip)
    mk = dict(c=c.cuda(), y=y.cuda(), cfg_scale=1.0, attn_mask=am.cuda(), key_padding_mask=None)
    fused = diff.p_sample_loop(dit.forward_with_cfg, x.shape, x.cuda(), denoised_fn=InpaintDenoiser(ip.cuda(), x.cuda()),
                               clip_denoised=True, model_kwargs=mk, step_noise=noise.cuda()).cpu()
    assert (fused - ref).abs().max() <= 1e-3, (fused - ref).abs().max()          # north_star tolerance: 1e-3 abs
    zc = x.cuda()
    closure = lambda xx: torch.where(ip.cuda(), xx, zc)                            # arbitrary host callable -> Python loop seam
    loop = diff.p_sample_loop(dit.forward_with_cfg, x.shape, x.cuda(), denoised_fn=closure, clip_denoised=True, model_kwargs=mk,
                              step_noise=noise.cuda()).cpu()
    assert (loop - ref).abs().max() <= 1e-3
