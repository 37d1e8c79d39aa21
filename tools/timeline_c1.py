#!/usr/bin/env python3
"""Critical path of the GEMM launches of one single-utterance sample (C1), from a -DZVB_TIMELINE build of the library
(debug build only: CTA 0 of every GEMM launch stamps clock64 at set-up done / dependency wait done / first operands landed /
MMAs issued / accumulator ready / epilogue done / all roles done).

  cd zipvoice_b200 && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC \
      -DZVB_TIMELINE '-DZVB_SOURCE_HASH="timeline"' -o libzvb_timeline.so csrc/engine.cu
  ZVB_LIB=$PWD/zipvoice_b200/libzvb_timeline.so python tools/timeline_c1.py [--batch 1] [--target 937]
"""
import argparse
import collections
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from zipvoice_b200 import _lib  # noqa: E402
from zipvoice_b200.config import ZipVoiceConfig  # noqa: E402
from zipvoice_b200.model import build_model  # noqa: E402
from zipvoice_b200.synth import synth_state_dict  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--prompt", type=int, default=281)
    ap.add_argument("--target", type=int, default=937)
    ap.add_argument("--steps", type=int, default=16)
    a = ap.parse_args()
    cfg = ZipVoiceConfig("zipvoice", vocab_size=360)
    model = build_model(cfg, synth_state_dict(cfg, 0), "cuda", use_cuda_graph=True)
    lib = C.CDLL(_lib.LIB_PATH)
    lib.zvb_debug_timeline.restype = C.c_int
    lib.zvb_debug_timeline.argtypes = [C.c_void_p, C.c_int]
    B, T, F = a.batch, a.prompt + a.target, cfg.feat_dim
    g = torch.Generator(device="cpu").manual_seed(1)
    x0 = torch.randn(B, T, F, generator=g).cuda()
    text = (torch.randn(B, T, F, generator=g) * 0.5).cuda()
    speech = torch.zeros(B, T, F)
    speech[:, : a.prompt] = torch.randn(B, a.prompt, F, generator=g) * 0.3 - 0.5
    speech = speech.cuda()
    mask = torch.zeros(B, T, dtype=torch.bool, device="cuda")
    kw = dict(num_step=a.steps, guidance_scale=1.0, t_shift=0.5)
    for _ in range(3):
        model.solver.sample(x=x0, text_condition=text, speech_condition=speech, padding_mask=mask, **kw)
    torch.cuda.synchronize()
    buf = np.zeros((1 << 15, 20), dtype=np.uint64)
    lib.zvb_debug_timeline(buf.ctypes.data, buf.shape[0])              # reset
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    model.solver.sample(x=x0, text_condition=text, speech_condition=speech, padding_mask=mask, **kw)
    e1.record()
    torch.cuda.synchronize()
    n = lib.zvb_debug_timeline(buf.ctypes.data, buf.shape[0])
    t = buf[:n].astype(np.int64)
    order = np.argsort(t[:, 8])
    t = t[order]
    ms = e0.elapsed_time(e1)
    print(f"sample: {ms:.2f} ms, {n} GEMM launches recorded ({n / a.steps:.0f} per step)")
    cyc = (t[:, 7] - t[:, 0]).astype(np.float64)
    ns = (t[:, 9] - t[:, 8]).astype(np.float64)
    ok = ns > 0
    ghz = float(cyc[ok].sum() / ns[ok].sum())
    print(f"SM clock from the stamps: {ghz:.3f} GHz; globaltimer step ~{np.min(np.diff(np.unique(t[:, 8])))} ns")
    names = ["set-up", "dependency wait", "first operands", "MMAs issued", "accumulator ready", "epilogue", "drain (stores, roles)"]
    groups = collections.OrderedDict()
    for r in t:
        M, n_out = int(r[10] >> 32), int(r[10] & 0xFFFFFFFF)
        kb, bn, cl, lean = int(r[11] >> 32), int((r[11] >> 8) & 0xFFF), int((r[11] >> 4) & 0xF), int(r[11] & 0xF)
        groups.setdefault((M, n_out, kb, bn, cl, lean), []).append(r)
    gap = (t[1:, 8] - t[:-1, 9]).astype(np.float64) / 1e3
    print(f"all GEMMs: in-kernel {cyc.mean() / ghz / 1e3:.2f} us mean; gap exit -> next GEMM entry {np.median(gap):.2f} us median "
          f"(other kernels sit in some gaps), sum in-kernel {cyc.sum() / ghz / 1e6:.2f} ms, sum gaps {gap.sum() / 1e3:.2f} ms")
    print("phases in us (mean over launches of CTA 0): " + " | ".join(names))
    for key, rows in sorted(groups.items(), key=lambda kv: -len(kv[1])):
        rr = np.array(rows, dtype=np.int64)
        d = np.diff(rr[:, :8].astype(np.float64), axis=1) / ghz / 1e3
        # stamps 3/4 come from the MMA thread, 5/6 from epilogue thread 0: keep them as differences along the chain
        line = " ".join(f"{v:6.2f}" for v in d.mean(axis=0))
        tot = (rr[:, 7] - rr[:, 0]).mean() / ghz / 1e3
        su = np.stack([rr[:, 16] - rr[:, 0], rr[:, 17] - rr[:, 0], rr[:, 18] - rr[:, 0], rr[:, 19] - rr[:, 0]], axis=1).astype(np.float64) / ghz / 1e3
        line += "   set-up (entry -> barriers initialised / TMEM allocated / block barrier / cluster barrier): " + " ".join(f"{v:5.2f}" for v in su.mean(axis=0))
        if key[5] == 0 and (rr[:, 12] > 0).all():      # generic linear epilogue: accumulator ready -> unit start -> TMEM read -> math -> stored
            g = np.stack([rr[:, 12] - rr[:, 5], rr[:, 13] - rr[:, 12], rr[:, 14] - rr[:, 13], rr[:, 15] - rr[:, 14],
                          rr[:, 6] - rr[:, 15]], axis=1).astype(np.float64) / ghz / 1e3
            line += "   generic epilogue: " + " ".join(f"{v:5.2f}" for v in g.mean(axis=0))
        print(f"  M={key[0]:6d} N={key[1]:5d} kb={key[2]:3d} bn={key[3]:3d} cl={key[4]} lean={key[5]} x{len(rows):5d}: {line}  total {tot:6.2f}")


if __name__ == "__main__":
    main()
