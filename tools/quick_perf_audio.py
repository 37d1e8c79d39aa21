#!/usr/bin/env python3
"""Timing of the stages either side of the sampler at the C3 per-GPU shape (GPU box): prompt log-mel of 64 x 3 s
waveforms, vocoder decode of 64 x 938 generated frames, with the vocoder's per-kernel breakdown."""
import argparse
import collections
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from zipvoice_b200.frontend import VocosFbank  # noqa: E402
from zipvoice_b200.vocoder import Vocos, synth_vocos_state_dict  # noqa: E402


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--prompt-sec", type=float, default=3.0)
    ap.add_argument("--frames", type=int, default=938)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    g = torch.Generator().manual_seed(0)
    S = int(a.prompt_sec * 24000)
    wav = (torch.randn(a.batch, S, generator=g) * 0.1).cuda()
    lens = torch.full((a.batch,), S)
    fe = VocosFbank()
    ms_fb = timed(lambda: fe.extract_batch(wav, lens, 0.1), a.reps)
    fb_bytes = wav.numel() * 4 + a.batch * ((S + 128) // 256) * 100 * 4
    voc = Vocos(frame_bucket=1).load_state_dict(synth_vocos_state_dict(0)).to("cuda")
    mel = (torch.randn(a.batch, a.frames, 100, generator=g) * 0.2 - 0.4).cuda()
    ml = torch.full((a.batch,), a.frames).cuda()
    ms_voc = timed(lambda: voc.decode_batch(mel, ml, scale=10.0, clamp=True), a.reps)
    plan = voc._plan(a.batch, a.frames)
    plan.profile()
    prof = plan.profile()
    agg = collections.OrderedDict()
    for cat, ms, work in prof:
        d = agg.setdefault(cat, [0, 0.0, 0.0])
        d[0] += 1; d[1] += ms; d[2] += work
    audio_s = a.batch * (a.frames - 1) * 256 / 24000
    out = dict(fbank_ms=ms_fb, fbank_gbs=fb_bytes / ms_fb / 1e6, fbank_frames_per_s=a.batch * ((S + 128) // 256) / ms_fb * 1e3,
               vocoder_ms=ms_voc, vocoder_frames_per_s=a.batch * a.frames / ms_voc * 1e3, vocoder_rtf=ms_voc / 1e3 / audio_s,
               vocoder_kernels={k: dict(n=v[0], ms=round(v[1], 3),
                                        rate=round(v[2] / (v[1] * 1e-3) / (1e12 if k.startswith("gemm") else 1e9), 1) if v[1] > 0 else 0)
                                for k, v in agg.items()},
               vocoder_workspace_gib=plan.workspace_bytes / 2**30)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
