#!/usr/bin/env python3
"""Quick timing + per-kernel-category breakdown of one sampler configuration (GPU box)."""
import argparse
import collections
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from zipvoice_b200.config import ZipVoiceConfig  # noqa: E402
from zipvoice_b200.model import build_model  # noqa: E402
from zipvoice_b200.synth import synth_state_dict  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", default="zipvoice")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--prompt", type=int, default=281)
    ap.add_argument("--target", type=int, default=938)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--guidance", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--detail", action="store_true")
    a = ap.parse_args()
    cfg = ZipVoiceConfig(a.variant, vocab_size=362 if "dialog" in a.variant else 360)
    t0 = time.time()
    model = build_model(cfg, synth_state_dict(cfg, 0), "cuda", use_cuda_graph=not a.no_graph)
    print(f"model built in {time.time() - t0:.1f}s", flush=True)
    B, T = a.batch, a.prompt + a.target
    F = cfg.feat_dim * (2 if cfg.is_stereo else 1)
    g = torch.Generator(device="cpu").manual_seed(1)
    x0 = torch.randn(B, T, F, generator=g).cuda()
    text = (torch.randn(B, T, cfg.feat_dim, generator=g) * 0.5).cuda()
    speech = torch.zeros(B, T, F)
    speech[:, : a.prompt] = torch.randn(B, a.prompt, F, generator=g) * 0.3 - 0.5
    speech = speech.cuda()
    mask = torch.zeros(B, T, dtype=torch.bool, device="cuda")
    kw = dict(num_step=a.steps, guidance_scale=a.guidance, t_shift=0.5)
    torch.cuda.synchronize()
    for rep in range(a.reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        x1 = model.solver.sample(x=x0, text_condition=text, speech_condition=speech, padding_mask=mask, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        frames = B * a.target
        print(f"rep {rep}: {ms:.1f} ms  -> {frames / ms * 1e3:,.0f} frames/s, RTF {ms / 1e3 / (frames * 256 / 24000):.5f}, "
              f"finite={bool(torch.isfinite(x1).all())}", flush=True)
    # per-kernel breakdown of one decoder forward
    N = 2 * B if (a.guidance != 0 and not cfg.is_distill) else B
    plan = model.solver.decoders[F].plans.get(N, T)
    plan.profile()
    prof = plan.profile()
    agg = collections.OrderedDict()
    for cat, ms, work in prof:
        d = agg.setdefault(cat, [0, 0.0, 0.0])
        d[0] += 1
        d[1] += ms
        d[2] += work
    tot = sum(d[1] for d in agg.values())
    print(f"one forward (N={N}, T={T}): {tot:.2f} ms over {len(prof)} kernels")
    for cat, (n, ms, work) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        rate = work / (ms * 1e-3) if ms > 0 else 0
        unit = "TFLOP/s" if cat.startswith("gemm") or cat == "attn_weights" else "GB/s"
        scale = 1e12 if unit == "TFLOP/s" else 1e9
        print(f"  {cat:16s} n={n:4d}  {ms:8.2f} ms  {100 * ms / tot:5.1f}%   {rate / scale:9.1f} {unit}")
    if a.detail:
        det = plan.profile(shapes=True, with_bytes=True)
        byk = collections.OrderedDict()
        for cat, ms, work, nbytes, shp in det:
            if cat == "dwconv_swooshr":
                d = byk.setdefault(shp[2], [0, 0.0, 0.0])
                d[0] += 1; d[1] += ms; d[2] += nbytes
        for k, (n, ms, nb) in byk.items():
            print(f"  dwconv K={k:2d}: n={n:3d} {ms:7.3f} ms  {nb / ms / 1e6:7.0f} GB/s")
        # first layer of the first stack: ops after the preamble up to the first biasnorm
        i0 = next(i for i, d in enumerate(det) if d[0] == "attn_weights") - 1
        i1 = next(i for i, d in enumerate(det) if d[0] == "biasnorm_bypass")
        print("first full-rate layer, kernel by kernel (roof = max(FLOPs / 1349.8 TF/s, bytes / 6532.9 GB/s)):")
        for cat, ms, work, nbytes, shp in det[i0:i1 + 1]:
            tensor = cat.startswith("gemm") or cat == "attn_weights"
            tf = work / (ms * 1e-3) / 1e12 if tensor and ms > 0 else 0.0
            gb = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
            roof = max((work / 1349.8e12 if tensor else 0.0), nbytes / 6532.9e9) * 1e3
            print(f"  {cat:16s} {str(shp):28s} {ms * 1e3:8.1f} us  {tf:7.1f} TFLOP/s {gb:7.0f} GB/s  roof {roof * 1e3:6.1f} us = {100 * roof / ms:4.0f}%")
    print(f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")


if __name__ == "__main__":
    main()
