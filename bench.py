#!/usr/bin/env python
"""bench.py — event tokens/sec of the Mapperatorinator inference hot path on B200 (contract: see README "Measurement").

Workload (BASELINE.json configs[1]): osuT5 v29 dimensions (whisper-small, 213 M params, fp32,
seeded random weights), one 180 s synthetic 16 kHz song -> 211 sequential windows (stride 13 094 samples), greedy decode,
`min_new_tokens = 64`, `max_length = P + 64` (random weights have no EOS behaviour, so the token budget is pinned:
211 x 64 = 13 504 event tokens per step), real look-back / look-ahead EOS sets and logits-processor chain, prompt =
16 conditioning ids + SOS + ctx_sos(MAP) (+ the last 32 generated ids of the previous window -> sequential dependency).

A "step" = one full song.
  value : tokens/s with the PCM windows already resident in HBM (engine path: one batched mel+encoder pass over all
          windows, cross-K/V resident, then the sequential prefill + token loop per window), CUDA-event timed.
  e2e   : the same song through the reference-facing call, `server.model_generate(model, tokenizer, model_kwargs,
          generate_kwargs)` once per window with HOST tensors (pinned PCM in, CPU LongTensor out) — H2D, per-call encoder
          re-run (as the reference does), D2H all inside the timed region.
  --impl reference : the CPU oracle port of the reference path (same per-window call pattern) on the host cores.
Multi-GPU (torchrun): one song per rank per step (weak scaling), NCCL gather of the token streams inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from mapperatorinator_b200 import TokenLayout, v29_model_config  # noqa: E402
from mapperatorinator_b200.pipeline import gather_token_streams, segment  # noqa: E402
from mapperatorinator_b200.weights import init_model_state_dict  # noqa: E402

SONG_SECONDS = 180.0
NEW_TOKENS = 64
COND_IDS = [3667, 3680, 3700, 3710, 3730, 3798, 3810, 3870, 3965, 3975, 3992, 4006, 4100, 3862, 3863, 3864]   # 16 input-only ids


def synth_song(seed: int, seconds: float = SONG_SECONDS, sr: int = 16000) -> np.ndarray:
    """8 log-spaced sinusoids 55 Hz-7 kHz + 120 BPM click train + N(0, 0.01) noise, peak-normalised."""
    rng = np.random.default_rng(seed)
    n = int(seconds * sr)
    t = np.arange(n) / sr
    x = sum(np.sin(2 * np.pi * f * t + rng.uniform(0, 2 * np.pi)) for f in np.geomspace(55, 7000, 8)) / 8
    clicks = np.zeros(n)
    clicks[(np.arange(0, seconds, 0.5) * sr).astype(int)] = 1.0
    x = x + np.convolve(clicks, np.hanning(64), mode="same") + rng.normal(0, 0.01, n)
    return (x / np.abs(x).max()).astype(np.float32)


def prompt_for(i: int, streams) -> list:
    base = COND_IDS + [1, 9]
    return base if i == 0 else base + streams[i - 1][-32:]


# one `ncu --set full` capture of decode_megakernel<1> (63 tokens): 25.731 GB read + 4.2 MB written (profiles/r1_megakernel_ncu.md)
MEGA_DRAM_BYTES_PER_LAUNCH = 25_730_578_000 + 4_189_184


def gen_kwargs(i: int, n_windows: int, P: int) -> dict:
    ms = 8184.0
    return dict(do_sample=False, num_beams=1, top_p=0.9, top_k=0, cfg_scale=1.0, timeshift_bias=0, types_first=True, temperature=0.9,
                timing_temperature=0.1, mania_column_temperature=0.5, taiko_hit_temperature=0.5, max_length=P + NEW_TOKENS,
                min_new_tokens=NEW_TOKENS, lookback_time=0.5 * ms if i > 0 else 0.0, lookahead_time=0.4 * ms if i < n_windows - 1 else 0.0,
                context_type="map")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.rows.append(l) for l in self.proc.stdout], daemon=True).start()
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()

    def summary(self) -> dict:
        sm, mx, reasons = [], 0, set()
        for l in self.rows:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args, rank: int, world: int) -> None:
    """CPU arm: the oracle port of the reference path, per-window `model_generate` calls exactly like Processor does."""
    if rank != 0:
        return
    from oracle import generate as gen_oracle
    cores = args.cpu_threads or min(os.cpu_count() or 1, 16)
    torch.set_num_threads(cores)
    cfg = v29_model_config()
    layout = TokenLayout.from_json(os.path.join(ROOT, "tests", "golden", "tokenizer_v29.json"))
    sd = init_model_state_dict(cfg, 0)
    windows, _, _ = segment(synth_song(0), cfg)
    n_windows = windows.shape[0]
    sample_windows = args.cpu_windows

    def step():
        streams, toks = [], 0
        for i in range(sample_windows):
            prompt = torch.tensor([prompt_for(i, streams)])
            ids, stats = gen_oracle.model_generate(sd, cfg, layout, dict(inputs=windows[i:i + 1], decoder_input_ids=prompt,
                                                                         decoder_attention_mask=prompt.ne(0)),
                                                   gen_kwargs(i, n_windows, prompt.shape[1]))
            streams.append(ids[0, prompt.shape[1]:].tolist()); toks += stats["generated_tokens"]
        return toks
    with torch.no_grad():
        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        toks = sum(step() for _ in range(args.steps))
        dt = time.perf_counter() - t0
    v = toks / dt
    sample = f"first {sample_windows} of {n_windows} sequential windows x {NEW_TOKENS} tokens per step, encoder re-run per call"
    print(json.dumps({
        "impl": "reference", "metric": "event tokens/sec end-to-end", "value": v, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(n_windows),
        "cpu_baseline": {"value": v, "unit": "tokens/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def workload_config(n_windows: int) -> dict:
    return {"workload": "osuT5 v29 full-song inference, 180 s synthetic 16 kHz audio, sequential sliding windows (configs[1], SURVEY 8d 2a)",
            "windows": n_windows, "new_tokens_per_window": NEW_TOKENS, "decode": "greedy, min_new_tokens=64", "batch": 1,
            "weights": "seeded random init, whisper-small dims (213M), fp32", "songs_per_gpu_per_step": 1,
            "l2": "inputs larger than L2: each token streams the 464 MB fp32 decoder (L2 = 126 MB)"}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-windows", type=int, default=4, help="windows per CPU step (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pdl", type=int, default=int(os.environ.get("MB200_PDL", "0")))
    ap.add_argument("--windows", type=int, default=0, help="debug: truncate the song to this many windows")
    ap.add_argument("--tc", type=int, default=int(os.environ.get("MB200_TC", "1")), help="1 = tcgen05 3xTF32 GEMMs where eligible, 0 = fp32 SIMT GEMM everywhere")
    ap.add_argument("--mega", type=int, default=1, help="1 = persistent token-loop megakernel (default), 0 = CUDA-graph replay per token")
    ap.add_argument("--cpu-threads", type=int, default=int(os.environ.get("MB200_CPU_THREADS", "0")),
                    help="torch threads of the CPU arm (0 = min(cores, 16): measured best on the GPU box; 32+ threads slow a batch-1 decoder down)")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from mapperatorinator_b200 import _lib
    from mapperatorinator_b200.modeling import B200Mapperatorinator
    from mapperatorinator_b200.pipeline import SongDecoder
    from mapperatorinator_b200.server import model_generate
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    cfg = v29_model_config()
    layout = TokenLayout.from_json(os.path.join(ROOT, "tests", "golden", "tokenizer_v29.json"))
    sd = init_model_state_dict(cfg, 0)                       # same weights on every rank
    windows, _, _ = segment(synth_song(rank), cfg)           # rank r decodes song r
    if args.windows:
        windows = windows[:args.windows]
    n_windows = windows.shape[0]
    model = B200Mapperatorinator(cfg, sd, max_windows=n_windows, max_batch=2, device=dev)
    del sd
    if args.pdl:
        model.engine.set_option("pdl", 1)
    model.engine.set_option("mega", args.mega)
    song = SongDecoder(model, layout)
    pinned = windows.pin_memory()
    resident = windows.to(dev)
    lib = _lib.load()
    lib.mb200_set_tensor_cores(int(args.tc))

    def step_resident():
        song.encode_song(resident)
        streams = song.decode_windows(n_windows, prompt_for, lambda i: gen_kwargs(i, n_windows, 18 if i == 0 else 50))
        if world > 1:
            gather_token_streams([sum(streams, [])], [rank])
        return sum(len(s) for s in streams), streams

    def step_e2e():
        streams, toks = [], 0
        for i in range(n_windows):
            prompt = torch.tensor([prompt_for(i, streams)])
            ids, stats = model_generate(model, layout, dict(inputs=pinned[i:i + 1], decoder_input_ids=prompt, decoder_attention_mask=prompt.ne(0)),
                                        gen_kwargs(i, n_windows, prompt.shape[1]))
            streams.append(ids[0, prompt.shape[1]:].tolist()); toks += stats["generated_tokens"]
        if world > 1:
            gather_token_streams([sum(streams, [])], [rank])
        return toks, streams

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.mb200_launch_count()
        mstats = np.zeros(3, dtype=np.float64)
        lib.mb200_model_mega_stats(model.engine.handle, mstats.ctypes.data, 1)      # reset the megakernel event counters
        e0.record()
        toks = 0
        for _ in range(steps):
            t, streams = fn()
            toks += t
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        launches = lib.mb200_launch_count() - l0
        lib.mb200_model_mega_stats(model.engine.handle, mstats.ctypes.data, 0)
        timed.mega = mstats.copy()
        if world > 1:
            tt = torch.tensor([ms, float(toks)], device=dev, dtype=torch.float64)
            mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = tt.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            ms, toks = float(mx[0]), float(sm[1])
        return ms, toks, launches, streams

    with ClockSampler(local) as clk:
        ms, toks, launches, streams = timed(step_resident, args.steps, args.warmup)
    clocks = clk.summary()
    mega_resident, ms_resident = timed.mega, ms
    ms_e2e, toks_e2e, _, streams2 = timed(step_e2e, max(1, args.steps // 2), 1)
    assert streams == streams2, "resident-encoder path and per-window drop-in path must emit identical tokens"

    # ---- roofline of the dominant kernel, timed live with CUDA events on the launching stream --------------------------------
    d, f, V, L = cfg.d_model, cfg.ffn_dim, cfg.vocab_size_out, cfg.decoder_layers
    w_bytes = 4 * (L * (3 * d * d + 2 * d * d + d * d + 2 * d * f) + V * d)                  # decoder weights streamed once per token
    ctx = 50 + NEW_TOKENS // 2
    kv_bytes = 4 * L * 2 * (cfg.max_source_positions + ctx) * d                               # cross + self K/V read per token
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    mega = mega_resident
    if args.mega and mega[0] > 0:
        # persistent token-loop kernel: one launch per window decodes NEW_TOKENS-1 tokens; events recorded around every launch
        tok_per_launch = mega[2] / mega[0]
        us_per_launch = 1000.0 * mega[1] / mega[0]
        bytes_per_launch = (w_bytes + kv_bytes) * tok_per_launch
        achieved = bytes_per_launch / (us_per_launch * 1e-6) / 1e9
        roofline = {"bound": "hbm", "kernel": "decode_megakernel<1> (persistent cooperative kernel: all layers of all tokens of one generate() call)",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": MEGA_DRAM_BYTES_PER_LAUNCH if abs(tok_per_launch - 63.0) < 1e-6 else None,
                    "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one 63-token launch (profiles/r1_megakernel_ncu.md)",
                    "peak_source": peak_src, "bytes_per_launch": bytes_per_launch, "us_per_launch": us_per_launch, "tokens_per_launch": tok_per_launch,
                    "us_per_token": us_per_launch / tok_per_launch, "bytes_per_token": w_bytes + kv_bytes,
                    "share_of_step": mega[1] / (ms_resident), "token_floor_us": (w_bytes + kv_bytes) / (peak * 1e3)}
    else:
        out_us = np.zeros(4, dtype=np.float32)
        _lib.check(lib.mb200_model_profile_step(model.engine.handle, 1, 1, 50 + NEW_TOKENS, 20, out_us.ctypes.data, torch.cuda.current_stream().cuda_stream))
        n_gemv = int(out_us[3]) // 1000000
        gemv_us = float(out_us[0])
        achieved = (w_bytes / n_gemv) / (gemv_us / n_gemv * 1e-6) / 1e9 if gemv_us > 0 else None
        roofline = {"bound": "hbm", "kernel": f"gemv_kernel<1> ({n_gemv} launches per token, CUDA-graph path; eager event timing includes launch gaps)",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if achieved else None, "traffic": None,
                    "peak_source": peak_src, "bytes_per_launch": w_bytes / n_gemv, "us_per_launch": gemv_us / n_gemv,
                    "per_token_us": {"gemv": gemv_us, "attention": float(out_us[1]), "sample": float(out_us[2])},
                    "token_floor_us": (w_bytes + kv_bytes) / (peak * 1e3)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        from oracle import generate as gen_oracle
        cores = args.cpu_threads or min(os.cpu_count() or 1, 16)
        torch.set_num_threads(cores)
        sd_cpu = init_model_state_dict(cfg, 0)
        cs, ctoks = [], 0
        t0 = time.perf_counter()
        with torch.no_grad():
            for i in range(args.cpu_windows):
                prompt = torch.tensor([prompt_for(i, cs)])
                ids, st = gen_oracle.model_generate(sd_cpu, cfg, layout, dict(inputs=windows[i:i + 1], decoder_input_ids=prompt,
                                                                              decoder_attention_mask=prompt.ne(0)),
                                                    gen_kwargs(i, n_windows, prompt.shape[1]))
                cs.append(ids[0, prompt.shape[1]:].tolist()); ctoks += st["generated_tokens"]
        cdt = time.perf_counter() - t0
        match = cs == streams[:args.cpu_windows]
        cpu = {"value": ctoks / cdt, "unit": "tokens/s", "cores": cores, "kind": "port",
               "sample": f"first {args.cpu_windows} of {n_windows} windows x {NEW_TOKENS} tokens, encoder re-run per call (reference call pattern)",
               "token_ids_match_gpu": bool(match)}
    h2d = n_windows * cfg.samples_per_window * 4
    d2h = n_windows * (50 + NEW_TOKENS) * 8
    print(json.dumps({
        "metric": "event tokens/sec end-to-end", "value": toks / (ms / 1000), "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(n_windows), "clocks": clocks,
        "e2e": {"value": toks_e2e / (ms_e2e / 1000), "unit": "tokens/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "server.model_generate per window (host tensors in, CPU LongTensor out)"},
        "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "pdl": bool(args.pdl), "tensor_cores": bool(args.tc),
        "token_stream_sha1": __import__("hashlib").sha1(json.dumps(streams).encode()).hexdigest()}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
