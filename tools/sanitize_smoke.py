#!/usr/bin/env python3
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): tiny ZipVoice sample without CUDA graphs,
a base-width decoder forward (lean GEMM epilogues, bias cache, CTA pairs), log-mel and a small vocoder.
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zipvoice_b200.config import ZipVoiceConfig, tiny_config  # noqa: E402
from zipvoice_b200.frontend import VocosFbank  # noqa: E402
from zipvoice_b200.model import build_model  # noqa: E402
from zipvoice_b200.synth import synth_state_dict, synth_utterances  # noqa: E402
from zipvoice_b200.vocoder import Vocos, synth_vocos_state_dict  # noqa: E402

cfg = tiny_config("zipvoice")
model = build_model(cfg, synth_state_dict(cfg, 0), "cuda", use_cuda_graph=False)
u = synth_utterances(cfg, batch=2, prompt_frames=20, target_frames=[60, 45], prompt_tokens=6, tokens=17)
out = model.sample(u["tokens"], u["prompt_tokens"], u["prompt_features"], u["prompt_features_lens"],
                   features_lens=u["target_lens"], duration="real", num_step=2, guidance_scale=1.0, t_shift=0.5)
torch.cuda.synchronize()
print("tiny sample", tuple(out[0].shape), bool(torch.isfinite(out[0]).all()))
if "--base" in sys.argv:
    cfg = ZipVoiceConfig("zipvoice")
    model = build_model(cfg, synth_state_dict(cfg, 0), "cuda", use_cuda_graph=False)
    x = torch.randn(2, 300, 300, device="cuda")
    y = model.fm_decoder(x=x, t=torch.tensor([0.3, 0.3], device="cuda"), padding_mask=torch.zeros(2, 300, dtype=torch.bool, device="cuda"))
    torch.cuda.synchronize()
    print("base forward", tuple(y.shape), bool(torch.isfinite(y).all()))
fe = VocosFbank()
f = fe.extract(torch.randn(1, 5000) * 0.1, 24000)
voc = Vocos().load_state_dict(synth_vocos_state_dict(0, dim=256, intermediate=512, n_layers=2)).to("cuda")
w = voc.decode(torch.randn(2, 100, 30).cuda())
torch.cuda.synchronize()
print("audio", tuple(f.shape), tuple(w.shape), bool(torch.isfinite(w).all()))
