#!/usr/bin/env python3
"""Benchmark of the ZipVoice sampler hot path (BASELINE.json metric: generated mel frames/s of
16-step ZipVoice sampling; p50 RTF).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm: the UNMODIFIED reference (staged under
                                                              # baseline/_ref) on the host cores

Headline workload (config C3 of BASELINE.json, per GPU): 64 utterances, 3 s prompt (281 frames, 45 tokens)
+ ~10 s target (938 frames, 150 tokens), ZipVoice 123M with seeded synthetic weights, 16 Euler steps,
classifier-free guidance 1.0, t_shift 0.5.  `value` / `e2e` are WEAK scaling: every rank samples its own 64
utterances, no collective on the data path (the output mels are gathered at the end).  One "step" = one `sample`
call over the rank's batch.  Prints ONE JSON line (rank 0).  Extra keys of the same line:

  strong          one fixed ragged set of 512 utterances (targets U[600,938] frames) partitioned over the N ranks
                  by `sharding.choose_partition` (length-sorted runs per rank, cuts balancing the padded cost; the
                  valid-frame balance of `partition_utterances` with ZVB_STRONG_LPT=1), each shard sampled through the
                  length-sorted batcher (`batcher.sample_batched`, 64-frame buckets) with host inputs, then `sharding.gather_mels`
                  (NCCL) and the device->host copy of the gathered mels on rank 0: strong scaling, wall = slowest
                  rank + gather
  other_configs   C1 (one utterance), C2 (distill, 64 utterances, 4 steps), C4 (60 s dialog), C5 (stereo, 16
                  utterances) timed through `solver.sample` (N = 1 runs only); C1 also end to end through `model.sample`
                  with host inputs and outputs (`e2e`: p50 / p90 latency of 20 calls)
  torch_gpu_baseline  the unmodified reference, eager PyTorch on cuda:0 (fp32 and bf16 autocast), informational
  stages          the steps either side of the sampler at the same shape: prompt log-mel of 64 x 3 s
                  host waveforms (zvb_fbank), vocoder decode of the 64 x 938 generated frames (zvb_vocoder_decode), and the
                  whole chain waveform -> log-mel -> model.sample -> vocoder -> waveform on the host (N = 1 runs only)
"""
from __future__ import annotations

import argparse
import collections
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "generated mel frames/s, 16-step ZipVoice sampling (CFG)"
UNIT = "frames/s"
FRAME_SEC = 256.0 / 24000.0
PROMPT_FRAMES, TARGET_FRAMES, PROMPT_TOKENS, TOKENS = 281, 938, 45, 150
NUM_STEP, GUIDANCE, T_SHIFT = 16, 1.0, 0.5
STRONG_UTTS, STRONG_LO, STRONG_HI = 512, 600, 938
REF_DIRS = ("/root/reference", os.path.join(ROOT, "baseline", "_ref"))


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tflops=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm_gbs=6650.0, tflops=1400.0, src="fallback")   # B200_PROFILING.md fallback


def workload_config(batch_per_gpu, n_gpus, workspace_gb=None):
    return {"workload": f"C3: ZipVoice 123M 16-step CFG sampling, {batch_per_gpu} utterances/GPU "
                        f"({PROMPT_FRAMES}+{TARGET_FRAMES} frames), guidance {GUIDANCE}, t_shift {T_SHIFT}",
            "utterances_per_gpu": batch_per_gpu, "global_utterances": batch_per_gpu * n_gpus,
            "prompt_frames": PROMPT_FRAMES, "target_frames": TARGET_FRAMES, "num_step": NUM_STEP,
            "parallelism": f"utterance-sharded x{n_gpus}, no data-path collective",
            "l2": ("working set (%s workspace per rank) exceeds the 126 MB L2; no explicit flush"
                   % (f"{workspace_gb:.1f} GB" if workspace_gb else "multi-GB"))}


def reference_path():
    for d in REF_DIRS:
        if os.path.isdir(os.path.join(d, "zipvoice", "models")):
            return d
    return None


def import_reference(variant: str):
    """The reference model class from the staged, unmodified package (None if absent)."""
    ref = reference_path()
    if ref is None:
        return None
    if ref not in sys.path:
        sys.path.insert(0, ref)
    import logging
    logging.disable(logging.WARNING)            # the reference warns about the optional k2 package on import
    try:
        from zipvoice.models.zipvoice import ZipVoice
        from zipvoice.models.zipvoice_dialog import ZipVoiceDialog, ZipVoiceDialogStereo
        from zipvoice.models.zipvoice_distill import ZipVoiceDistill
    finally:
        logging.disable(logging.NOTSET)
    return dict(zipvoice=ZipVoice, zipvoice_distill=ZipVoiceDistill, zipvoice_dialog=ZipVoiceDialog,
                zipvoice_dialog_stereo=ZipVoiceDialogStereo)[variant]


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_samples(n_steps: int, n_warm: int):
    """Times full `model.sample` calls (16 CFG Euler steps, text encoder, split) of ONE utterance
    (281+938 frames) on all host cores: the unmodified reference when its package is staged
    (kind "reference"), else the oracle port.  Returns (frames/s, seconds per sample, cores, kind)."""
    import torch
    from zipvoice_b200.config import ZipVoiceConfig
    from zipvoice_b200.synth import synth_state_dict, synth_utterances
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = ZipVoiceConfig("zipvoice")
    sd = synth_state_dict(cfg, 0)
    u = synth_utterances(cfg, batch=1, prompt_frames=PROMPT_FRAMES, target_frames=TARGET_FRAMES,
                         prompt_tokens=PROMPT_TOKENS, tokens=TOKENS, seed=1)
    kw = dict(features_lens=u["target_lens"], duration="real", num_step=NUM_STEP, guidance_scale=GUIDANCE, t_shift=T_SHIFT)
    cls = import_reference("zipvoice")
    if cls is not None:
        kind = "reference"
        model = cls(**cfg.model_kwargs()).eval()
        model.load_state_dict(sd, strict=True)
        run = lambda: model.sample(u["tokens"], u["prompt_tokens"], u["prompt_features"], u["prompt_features_lens"], **kw)
    else:
        from oracle import zipvoice_oracle as orc
        kind = "port"
        model = orc.OracleModel(cfg, sd)
        run = lambda: model.sample(u["tokens"], u["prompt_tokens"], u["prompt_features"], u["prompt_features_lens"], **kw)
    times = []
    with torch.inference_mode():
        for i in range(n_warm + n_steps):
            t0 = time.perf_counter()
            out = run()
            dt = time.perf_counter() - t0
            assert int(out[1][0]) == TARGET_FRAMES
            if i >= n_warm:
                times.append(dt)
    per = statistics.mean(times)
    return TARGET_FRAMES / per, per, cores, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps, warm = max(1, args.steps), max(0, min(args.warmup, 1))
    value, per, cores, kind = cpu_reference_samples(steps, warm)
    sample = (f"each timed step = one full model.sample of 1 utterance ({PROMPT_FRAMES}+{TARGET_FRAMES} frames, 16 CFG "
              f"Euler steps); {steps} timed after {warm} warm-up; "
              + ("unmodified reference package (baseline/_ref), fp32, torch CPU" if kind == "reference"
                 else "oracle port (reference package not staged), fp32, torch CPU"))
    cfg = workload_config(args.batch_per_gpu, args.gpus)
    cfg["reference_sample"] = f"bounded sample of the workload: 1 of the {args.batch_per_gpu} utterances per step"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": per * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "rtf": per / (TARGET_FRAMES * FRAME_SEC), "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------- GPU arm
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons, power = [], [], set(), []
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def fwd_flops(T: int, in_dim: int = 300, out_dim: int = 100) -> float:
    """Algorithmic FLOPs of one fm_decoder forward per batch row."""
    W = 512 * 272 + 2 * (512 * 48 + 48 * 512) + 2 * 512 * (1152 + 1536 + 1920) + (512 * 1152 + 384 * 512) + 2 * (512 * 1024 + 512 * 512)
    f = 0.0
    for ds, nl, k in [(1, 2, 31), (2, 2, 15), (4, 4, 7), (2, 4, 15), (1, 4, 31)]:
        L = -(-T // ds)
        f += nl * (2.0 * W * L + 2.0 * L * L * 32 * 4 + 2.0 * L * (2 * L - 1) * 4 * 4 + 2 * (2.0 * L * L * 12 * 4) + 2.0 * L * L * 384
                   + 2 * (2.0 * L * 512 * k) + 2.0 * (2 * L - 1) * 48 * 16)
    return f + 2.0 * T * (in_dim * 512 + 512 * out_dim)


def time_other_configs(dev, peaks):
    """C1 / C2 / C4 / C5 of BASELINE.json through `solver.sample` (device-resident inputs, CUDA graph), 3 timed
    calls after 2 warm-ups each: ms per sample, generated frames/s, RTF, and the fraction of the tensor-peak
    ideal time (SURVEY.md §8d FLOP model / MEASURED_PEAKS sustained)."""
    import torch
    from zipvoice_b200.config import ZipVoiceConfig
    from zipvoice_b200.model import build_model
    from zipvoice_b200.synth import synth_state_dict
    specs = [  # name, variant, B, prompt, target, steps, guidance
        ("C1_single_utterance", "zipvoice", 1, 281, 937, 16, 1.0),
        ("C2_distill_64x4step", "zipvoice_distill", 64, 281, 938, 4, 3.0),
        ("C4_dialog_60s", "zipvoice_dialog", 1, 938, 5625, 16, 1.5),
        ("C5_stereo_16", "zipvoice_dialog_stereo", 16, 469, 1875, 16, 1.5),
    ]
    out = {}
    for name, variant, B, Pf, Tg, steps, g in specs:
        cfg = ZipVoiceConfig(variant, vocab_size=362 if "dialog" in variant else 360)
        model = build_model(cfg, synth_state_dict(cfg, 0), dev, use_cuda_graph=True)
        F = cfg.feat_dim * (2 if cfg.is_stereo else 1)
        T = Pf + Tg
        gen = torch.Generator().manual_seed(2)
        x0 = torch.randn(B, T, F, generator=gen).to(dev)
        text = (torch.randn(B, T, cfg.feat_dim, generator=gen) * 0.5).to(dev)
        speech = torch.zeros(B, T, F)
        speech[:, :Pf] = torch.randn(B, Pf, F, generator=gen) * 0.3 - 0.5
        speech = speech.to(dev)
        mask = torch.zeros(B, T, dtype=torch.bool, device=dev)
        kw = dict(num_step=steps, guidance_scale=g, t_shift=0.5)
        call = lambda: model.solver.sample(x=x0, text_condition=text, speech_condition=speech, padding_mask=mask, **kw)
        with torch.inference_mode():
            for _ in range(2):
                x1 = call()
            torch.cuda.synchronize()
            ms = []
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); x1 = call(); e1.record()
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
        med = statistics.median(ms)
        rows = B if cfg.is_distill else 2 * B
        in_dim = 2 * F + cfg.feat_dim
        ideal_ms = steps * rows * fwd_flops(T, in_dim, F) / (peaks["tflops"] * 1e12) * 1e3
        out[name] = {"model": variant, "utterances": B, "frames": f"{Pf}+{Tg}", "steps": steps, "guidance": g,
                     "ms_per_sample": med, "frames_per_s": B * Tg / (med * 1e-3), "rtf": med * 1e-3 / (B * Tg * FRAME_SEC),
                     "ideal_ms_tensor_peak": ideal_ms, "frac_of_ideal": ideal_ms / med, "finite": bool(torch.isfinite(x1).all())}
        if name == "C1_single_utterance":
            out[name]["e2e"] = time_single_utterance_e2e(model, cfg, dev, Pf, Tg, kw)
        del model, x0, text, speech, mask, x1
        torch.cuda.empty_cache()
    return out


def time_single_utterance_e2e(model, cfg, dev, Pf, Tg, kw, calls=20):
    """Latency of ONE utterance through the public API: host tokens + pinned prompt mel in, `model.sample` (text prelude and
    encoder, duration rule, CUDA-graph sampler), host mel out; wall clock per call with a stream synchronize, p50 / p90."""
    import torch
    from zipvoice_b200.synth import synth_utterances
    u = synth_utterances(cfg, batch=1, prompt_frames=Pf, target_frames=Tg, prompt_tokens=PROMPT_TOKENS, tokens=TOKENS, seed=5)
    pf_host, pfl_host, tl_host = u["prompt_features"].pin_memory(), u["prompt_features_lens"].pin_memory(), u["target_lens"].pin_memory()
    out_host = torch.empty(1, Tg, cfg.feat_dim).pin_memory()

    def call():
        mel, _, _, _ = model.sample(u["tokens"], u["prompt_tokens"], pf_host.to(dev, non_blocking=True),
                                    pfl_host.to(dev, non_blocking=True), features_lens=tl_host.to(dev, non_blocking=True),
                                    duration="real", **kw)
        out_host.copy_(mel, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    with torch.inference_mode():
        for _ in range(3):
            call()
        ms = []
        for _ in range(calls):
            t0 = time.perf_counter()
            call()
            ms.append((time.perf_counter() - t0) * 1e3)
    ms.sort()
    p50, p90 = ms[len(ms) // 2], ms[int(len(ms) * 0.9)]
    return {"calls": calls, "ms_p50": p50, "ms_p90": p90, "rtf_p50": p50 * 1e-3 / (Tg * FRAME_SEC),
            "h2d_bytes": pf_host.numel() * 4 + 16, "d2h_bytes": out_host.numel() * 4,
            "finite": bool(torch.isfinite(out_host).all())}


def time_torch_gpu_baseline(dev):
    """The unmodified reference, eager PyTorch on cuda:0 -- what a user of the reference runs on this box
    (SURVEY.md §2.1): 32 utterances of the C3 shape, 16 CFG steps, fp32 (TF32 off) and bf16 autocast."""
    import torch
    from zipvoice_b200.config import ZipVoiceConfig
    from zipvoice_b200.synth import synth_state_dict, synth_utterances
    cls = import_reference("zipvoice")
    if cls is None:
        return None
    cfg = ZipVoiceConfig("zipvoice")
    model = cls(**cfg.model_kwargs()).eval()
    model.load_state_dict(synth_state_dict(cfg, 0), strict=True)
    model = model.to(dev)
    B = 32
    u = synth_utterances(cfg, batch=B, prompt_frames=PROMPT_FRAMES, target_frames=TARGET_FRAMES,
                         prompt_tokens=PROMPT_TOKENS, tokens=TOKENS, seed=9)
    pf, pfl, tl = u["prompt_features"].to(dev), u["prompt_features_lens"].to(dev), u["target_lens"].to(dev)
    kw = dict(features_lens=tl, duration="real", num_step=NUM_STEP, guidance_scale=GUIDANCE, t_shift=T_SHIFT)
    res = {"utterances": B, "sample": f"{B} utterances ({PROMPT_FRAMES}+{TARGET_FRAMES} frames), 16 CFG steps, model.sample, "
                                      "1 timed call after 1 warm-up"}
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    for tag, ctx in (("fp32", None), ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
        try:
            with torch.inference_mode():
                for i in range(2):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    if ctx is None:
                        out = model.sample(u["tokens"], u["prompt_tokens"], pf, pfl, **kw)
                    else:
                        with ctx:
                            out = model.sample(u["tokens"], u["prompt_tokens"], pf, pfl, **kw)
                    torch.cuda.synchronize()
                    dt = time.perf_counter() - t0
            res[tag] = {"frames_per_s": B * TARGET_FRAMES / dt, "ms_per_sample": dt * 1e3,
                        "finite": bool(torch.isfinite(out[0]).all())}
        except Exception as e:                  # informational leg: never fails the benchmark
            res[tag] = {"error": f"{type(e).__name__}: {e}"[:200]}
    del model
    torch.cuda.empty_cache()
    return res


def time_audio_stages(dev):
    """Prompt waveform -> log-mel (f3), generated mel -> waveform (f1), and the chain around `model.sample` with host
    waveforms in and host waveforms out (pinned buffers, copies inside the timed region)."""
    import torch
    from zipvoice_b200.config import ZipVoiceConfig
    from zipvoice_b200.frontend import VocosFbank
    from zipvoice_b200.model import build_model
    from zipvoice_b200.synth import synth_state_dict, synth_utterances
    from zipvoice_b200.vocoder import Vocos, synth_vocos_state_dict
    B, S = 64, PROMPT_FRAMES * 256
    cfg = ZipVoiceConfig("zipvoice")
    model = build_model(cfg, synth_state_dict(cfg, 0), dev, use_cuda_graph=True)
    u = synth_utterances(cfg, batch=B, prompt_frames=PROMPT_FRAMES, target_frames=TARGET_FRAMES,
                         prompt_tokens=PROMPT_TOKENS, tokens=TOKENS, seed=31)
    gen = torch.Generator().manual_seed(32)
    wav_host = (torch.randn(B, S, generator=gen) * 0.05).pin_memory()
    wav_lens = torch.full((B,), S)
    tl_host = u["target_lens"].pin_memory()
    fe = VocosFbank(device=dev)
    voc = Vocos(frame_bucket=1).load_state_dict(synth_vocos_state_dict(0)).to(dev)
    out_host = torch.empty(B, 256 * (TARGET_FRAMES - 1)).pin_memory()
    kw = dict(duration="real", num_step=NUM_STEP, guidance_scale=GUIDANCE, t_shift=T_SHIFT)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    marks = {}

    def chain(record=False):
        e = [ev() for _ in range(4)] if record else None
        wav = wav_host.to(dev, non_blocking=True)
        if record: e[0].record()
        feats, frames = fe.extract_batch(wav, wav_lens, 0.1)
        if record: e[1].record()
        mel, lens, _, _ = model.sample(u["tokens"], u["prompt_tokens"], feats, frames, features_lens=tl_host.to(dev, non_blocking=True), **kw)
        if record: e[2].record()
        audio, alens = voc.decode_batch(mel, lens, scale=10.0, clamp=True)
        if record: e[3].record()
        out_host.copy_(audio, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        if record:
            marks.update(fbank_ms=e[0].elapsed_time(e[1]), sample_ms=e[1].elapsed_time(e[2]), vocoder_ms=e[2].elapsed_time(e[3]))
        return frames, lens, alens

    with torch.inference_mode():
        for _ in range(2):
            chain()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            frames, lens, alens = chain()
        sec = (time.perf_counter() - t0) / reps
        chain(record=True)
    gen_frames = int(lens.sum())
    audio_s = float(alens.sum()) / 24000.0
    res = {"utterances": B, "prompt_seconds": S / 24000.0, "prompt_frames": int(frames[0]), "generated_frames": gen_frames,
           "fbank_ms": marks["fbank_ms"], "sample_ms": marks["sample_ms"], "vocoder_ms": marks["vocoder_ms"],
           "fbank_gbs": (B * S * 4 + B * int(frames[0]) * 400) / marks["fbank_ms"] / 1e6,
           "vocoder_frames_per_s": gen_frames / marks["vocoder_ms"] * 1e3,
           "chain_ms": sec * 1e3, "chain_frames_per_s": gen_frames / sec, "chain_rtf": sec / audio_s,
           "h2d_bytes": wav_host.numel() * 4 + tl_host.numel() * 8, "d2h_bytes": out_host.numel() * 4,
           "finite": bool(torch.isfinite(out_host).all()),
           "note": "synthetic vocoder weights with the key set of vocos-mel-24khz (no checkpoint offline)"}
    del model, voc, fe
    torch.cuda.empty_cache()
    return res


def run_strong(args, dev, rank, world, barrier, max_over_ranks):
    """Strong scaling: ONE ragged set of 512 utterances for the whole job (module docstring)."""
    import torch
    from zipvoice_b200.batcher import padding_waste, plan_batches, sample_batched
    from zipvoice_b200.config import ZipVoiceConfig
    from zipvoice_b200.model import build_model
    from zipvoice_b200.sharding import choose_partition, gather_mels, partition_utterances
    from zipvoice_b200.synth import synth_state_dict, synth_utterances
    cfg = ZipVoiceConfig("zipvoice")
    gen = torch.Generator().manual_seed(4242)
    tgt = torch.randint(STRONG_LO, STRONG_HI + 1, (STRONG_UTTS,), generator=gen)
    u = synth_utterances(cfg, batch=STRONG_UTTS, prompt_frames=PROMPT_FRAMES, target_frames=tgt.tolist(),
                         prompt_tokens=PROMPT_TOKENS, tokens=TOKENS, seed=4243)
    total = u["features_lens"].tolist()
    bucket = 64
    if os.environ.get("ZVB_STRONG_LPT", "0") == "1":     # round-2a policy: valid frames balanced, full length range per rank
        shards, max_rows = partition_utterances(total, world), 64
    else:                                                # length-sorted runs per rank, cuts balance the padded cost
        shards, max_rows = choose_partition(total, world, frame_bucket=bucket)
    shard = shards[rank]
    per_rank = max(len(s) for s in shards)
    model = build_model(cfg, synth_state_dict(cfg, 0), dev, use_cuda_graph=True, frame_bucket=bucket)
    toks = [u["tokens"][i] for i in shard]
    ptoks = [u["prompt_tokens"][i] for i in shard]
    idx = torch.tensor(shard)
    pf_host = u["prompt_features"][idx].pin_memory()
    pfl_host = u["prompt_features_lens"][idx].pin_memory()
    tl_host = u["target_lens"][idx].pin_memory()
    max_frames = int(u["target_lens"].max())
    out_host = torch.empty(STRONG_UTTS, max_frames, cfg.feat_dim).pin_memory() if rank == 0 else None
    kw = dict(num_step=NUM_STEP, guidance_scale=GUIDANCE, t_shift=T_SHIFT)

    def step():
        pf = pf_host.to(dev, non_blocking=True)
        pfl = pfl_host.to(dev, non_blocking=True)
        tl = tl_host.to(dev, non_blocking=True)
        mel, lens, _, _ = sample_batched(model, toks, ptoks, pf, pfl, features_lens=tl, max_rows=max_rows, **kw)
        full, full_lens = gather_mels(mel, lens, shard, STRONG_UTTS, max_frames, per_rank=per_rank)
        if rank == 0:
            out_host.copy_(full, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return full_lens

    with torch.inference_mode():
        step()                                   # builds the plans and captures the graphs of every bucket
        barrier()
        reps = 2
        t0 = time.perf_counter()
        for _ in range(reps):
            lens = step()
        barrier()
        sec = max_over_ranks(time.perf_counter() - t0) / reps
    ok = bool(torch.equal(lens.cpu(), u["target_lens"]))
    my_batches = plan_batches([total[i] for i in shard], max_rows=max_rows, frame_bucket=bucket)
    waste = padding_waste([total[i] for i in shard], my_batches, bucket)
    rank_frames = [sum(total[i] for i in s) for s in shards]
    plans = model.solver.decoders[cfg.feat_dim].plans
    res = {"utterances": STRONG_UTTS, "target_frames": f"U[{STRONG_LO},{STRONG_HI}]", "generated_frames": int(u["target_lens"].sum()),
           "seconds": sec, "frames_per_s": int(u["target_lens"].sum()) / sec, "rtf": sec / (int(u["target_lens"].sum()) * FRAME_SEC),
           "batches_on_rank0": len(my_batches), "padding_waste_rank0": waste, "frame_bucket": bucket,
           "partition": "lpt" if os.environ.get("ZVB_STRONG_LPT", "0") == "1" else "sorted runs, padded-cost balanced",
           "max_rows": max_rows, "utterances_per_rank": [len(s) for s in shards],
           "plans_built_rank0": plans.created, "frames_per_rank_max_over_mean": max(rank_frames) / (sum(rank_frames) / world),
           "h2d_bytes": pf_host.numel() * 4 + pfl_host.numel() * 8 + tl_host.numel() * 8,
           "d2h_bytes_rank0": STRONG_UTTS * max_frames * cfg.feat_dim * 4, "lengths_ok": ok,
           "timing": f"wall clock of {reps} passes (after 1 warm-up pass), barrier + synchronize both sides, max over ranks"}
    del model
    torch.cuda.empty_cache()
    return res


def run_b200(args):
    import torch
    import torch.distributed as dist
    from zipvoice_b200 import _lib
    from zipvoice_b200.config import ZipVoiceConfig
    from zipvoice_b200.model import build_model
    from zipvoice_b200.synth import synth_state_dict, synth_utterances

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    B = args.batch_per_gpu
    cfg = ZipVoiceConfig("zipvoice")
    model = build_model(cfg, synth_state_dict(cfg, 0), dev, use_cuda_graph=True)
    u = synth_utterances(cfg, batch=B, prompt_frames=PROMPT_FRAMES, target_frames=TARGET_FRAMES,
                         prompt_tokens=PROMPT_TOKENS, tokens=TOKENS, seed=666 + rank)
    kw = dict(num_step=NUM_STEP, guidance_scale=GUIDANCE, t_shift=T_SHIFT)
    pf_host = u["prompt_features"].pin_memory()
    pfl_host = u["prompt_features_lens"].pin_memory()
    tgt_host = u["target_lens"].pin_memory()
    frames_per_step = int(u["target_lens"].sum())
    audio_sec = frames_per_step * FRAME_SEC

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident leg: conditions already in HBM, K calls of solver.sample
    with torch.inference_mode():
        tc, pm = model.forward_text_inference_gt_duration(
            tokens=u["tokens"], features_lens=u["target_lens"].to(dev), prompt_tokens=u["prompt_tokens"],
            prompt_features_lens=u["prompt_features_lens"].to(dev))
        T = tc.shape[1]
        sc = torch.zeros(B, T, cfg.feat_dim, device=dev)
        sc[:, :PROMPT_FRAMES] = u["prompt_features"].to(dev)
        x0 = u["x0"].to(dev)
        lc0 = _lib.launch_count()
        model.solver.sample(x=x0, text_condition=tc, speech_condition=sc, padding_mask=pm, **kw)   # captures the graph
        launches_per_sample = _lib.launch_count() - lc0
        for _ in range(max(args.warmup, 3) - 1):
            model.solver.sample(x=x0, text_condition=tc, speech_condition=sc, padding_mask=pm, **kw)
        barrier()
        clocks = ClockSampler(local)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        evs[0].record()
        for i in range(args.steps):
            x1 = model.solver.sample(x=x0, text_condition=tc, speech_condition=sc, padding_mask=pm, **kw)
            evs[i + 1].record()
        barrier()
        step_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
        total_ms = max_over_ranks(evs[0].elapsed_time(evs[-1]))
        clock_info = clocks.stop()
        finite = bool(torch.isfinite(x1).all())

        # ---- end-to-end leg: the public API with HOST inputs (tokens, pinned prompt mel), H2D and D2H inside
        out_host = torch.empty(B, TARGET_FRAMES, cfg.feat_dim).pin_memory()
        lens_host = torch.empty(B, dtype=torch.int64).pin_memory()

        def e2e_step():
            pf = pf_host.to(dev, non_blocking=True)
            pfl = pfl_host.to(dev, non_blocking=True)
            tl = tgt_host.to(dev, non_blocking=True)
            mel, lens, _, _ = model.sample(u["tokens"], u["prompt_tokens"], pf, pfl, features_lens=tl,
                                           duration="real", **kw)
            if world > 1:     # the only collective: gather the output mels (NCCL over NVLink)
                gathered = torch.empty(world, *mel.shape, device=dev, dtype=mel.dtype)
                dist.all_gather_into_tensor(gathered, mel.contiguous())
            out_host.copy_(mel, non_blocking=True)
            lens_host.copy_(lens, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
    h2d = pf_host.numel() * 4 + pfl_host.numel() * 8 + tgt_host.numel() * 8
    d2h = out_host.numel() * 4 + lens_host.numel() * 8

    # ---- roofline: CUDA events around every kernel of one decoder forward (on the launching stream), each
    # kernel against ITS OWN bound: max(algorithmic FLOPs / tensor peak, algorithmic HBM bytes / copy peak).
    plan = model.solver.decoders[cfg.feat_dim].plans.get(2 * B, T)
    plan.profile()
    peaks = load_peaks()
    agg = collections.OrderedDict()
    reps = 3
    for _ in range(reps):
        for cat, ms, work, nbytes in plan.profile(with_bytes=True):
            tensor = cat.startswith("gemm") or cat == "attn_weights"
            flops = work if tensor else 0.0
            t_t = flops / (peaks["tflops"] * 1e12) * 1e3
            t_h = nbytes / (peaks["hbm_gbs"] * 1e9) * 1e3
            d = agg.setdefault(cat, dict(n=0, ms=0.0, flops=0.0, bytes=0.0, roof=0.0, roof_t=0.0, roof_h=0.0))
            d["n"] += 1; d["ms"] += ms; d["flops"] += flops; d["bytes"] += nbytes
            d["roof"] += max(t_t, t_h)
            if t_t >= t_h:
                d["roof_t"] += t_t
            else:
                d["roof_h"] += t_h
    tot_ms = sum(d["ms"] for d in agg.values())
    kernels = []
    for cat, d in agg.items():
        bound = "tensor" if d["roof_t"] >= d["roof_h"] else "hbm"
        rate = (d["flops"] / 1e12 if bound == "tensor" else d["bytes"] / 1e9) / (d["ms"] * 1e-3) if d["ms"] > 0 else 0.0
        kernels.append({"kernel": cat, "launches_per_forward": d["n"] // reps, "share_of_forward": d["ms"] / tot_ms,
                        "bound": bound, "achieved": rate, "unit": "TFLOP/s" if bound == "tensor" else "GB/s",
                        "peak": peaks["tflops"] if bound == "tensor" else peaks["hbm_gbs"],
                        # time the kernel's own roofline allows / time measured (mixed bound, launch by launch)
                        "frac": d["roof"] / d["ms"] if d["ms"] > 0 else 0.0,
                        "tflops": d["flops"] / 1e12 / (d["ms"] * 1e-3) if d["ms"] > 0 else 0.0,
                        "gbs": d["bytes"] / 1e9 / (d["ms"] * 1e-3) if d["ms"] > 0 else 0.0})
    # dominant kernel = gemm_kernel (every epilogue variant: linear, gated, P.V share one kernel template)
    gemm = [d for cat, d in agg.items() if cat.startswith("gemm")]
    g_n = sum(d["n"] for d in gemm); g_ms = sum(d["ms"] for d in gemm); g_work = sum(d["flops"] for d in gemm)
    g_roof = sum(d["roof"] for d in gemm); g_bytes = sum(d["bytes"] for d in gemm)
    achieved = g_work / (g_ms * 1e-3) / 1e12
    traffic, traffic_src = None, None
    for name in ("traffic_r2.json", "traffic_r1.json"):   # avg dram bytes per gemm_kernel launch of an ncu --set full capture
        tpath = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("gemm_kernel_dram_bytes_per_launch")
            traffic_src = f"committed ncu capture profiles/{name} (not measured in this run)"
            break
    roofline = {"kernel": "gemm_kernel (tcgen05, all epilogue variants)", "bound": "tensor", "achieved": achieved,
                "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"], "traffic": traffic,
                "traffic_source": traffic_src,
                "algorithmic_flops_per_launch": g_work / g_n, "algorithmic_bytes_per_launch": g_bytes / g_n,
                "avg_launch_ms": g_ms / g_n, "launches_per_forward": g_n // reps,
                # launch by launch against max(tensor time, HBM time): the HBM-bound launches (residual-stream
                # GEMMs with K <= 512, SelfAttention P.V) cannot reach the tensor peak by construction
                "frac_of_own_roofline": g_roof / g_ms,
                "peak_source": f"{peaks['src']} (sustained figure: kernel timed inside a long step)",
                "share_of_forward": g_ms / tot_ms, "forward_ms": tot_ms / reps,
                "forward_frac_of_roofline": sum(d["roof"] for d in agg.values()) / tot_ms}
    workspace_gb = plan.workspace_bytes / 1e9
    del plan, model, tc, pm, sc, x0, x1
    torch.cuda.empty_cache()

    strong = None if args.no_strong else run_strong(args, dev, rank, world, barrier, max_over_ranks)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    other = torch_gpu = cpu_baseline = stages = None
    if world == 1 and not args.quick:
        def guarded(fn, *a):                       # informational legs never take the headline line down with them
            try:
                return fn(*a)
            except Exception as e:                  # noqa: BLE001
                torch.cuda.empty_cache()
                return {"error": f"{type(e).__name__}: {e}"[:300]}
        other = guarded(time_other_configs, dev, peaks)
        stages = guarded(time_audio_stages, dev)
        torch_gpu = guarded(time_torch_gpu_baseline, dev)
    if world == 1 and not args.no_cpu_baseline:
        v, per, cores, kind = cpu_reference_samples(1, 0)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                        "sample": f"one full model.sample of 1 utterance ({PROMPT_FRAMES}+{TARGET_FRAMES} frames, 16 CFG Euler "
                                  f"steps): {per:.1f} s, no warm-up"}
    value = frames_per_step * world * args.steps / (total_ms * 1e-3)
    e2e_value = frames_per_step * world * args.steps / e2e_s
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
            "config": workload_config(B, world, workspace_gb),
            "rtf_p50": statistics.median(step_ms) * 1e-3 / audio_sec,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s * 1e3 / args.steps, "rtf": e2e_s / args.steps / audio_sec},
            "gpu_launches": launches_per_sample * args.steps,
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu_baseline, "clocks": clock_info,
            "finite": finite, "strong": strong, "other_configs": other, "stages": stages, "torch_gpu_baseline": torch_gpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch-per-gpu", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the 512-utterance strong-scaling leg")
    ap.add_argument("--quick", action="store_true", help="skip the other-configs and torch-GPU legs (N = 1 extras)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
