#!/usr/bin/env python3
"""Generate the FULL-SIZE fixtures tests/golden/full_*.pt by running the UNMODIFIED reference on
the CPU (build container only; a few minutes on 8 cores).

Usage:  PYTHONDONTWRITEBYTECODE=1 python tools/make_golden_full.py [case ...]

These hold the reference's fp32 results at BASELINE.json's shapes and step counts (SURVEY.md §8
config table): C1/C3 (T = 281 + 937, 16 CFG steps, g = 1.0, t_shift = 0.5), a ragged three-utterance
C3 batch, C2 (distill, 4 steps, g = 3.0), C5 (stereo, T = 2344, both CFG branches), C4 (dialog,
T = 6563, one decoder forward + both CFG branches) and two runs on the reference's OWN
`torch.manual_seed(0)` initialisation (a tiny model whose state_dict is stored in the fixture, and
the 123 M model whose state_dict is rebuilt from the staged reference package, baseline/_ref).

To stay small a fixture stores: the text-encoder output before the frame gather (`text_embed`,
(B, S, 100)), the final state `x1` in full, and the CFG-blended velocities of a few steps at a
frame stride.  Everything else (tokens, prompt mels, x0) is regenerated from the seed by
`zipvoice_b200.synth.synth_utterances`.
"""
import hashlib
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference" if os.path.isdir("/root/reference") else os.path.join(ROOT, "baseline", "_ref")
sys.path.insert(0, REF)

from zipvoice_b200.synth import synth_state_dict, synth_utterances  # noqa: E402

from zipvoice.models.zipvoice import ZipVoice  # noqa: E402
from zipvoice.models.zipvoice_dialog import ZipVoiceDialog, ZipVoiceDialogStereo  # noqa: E402
from zipvoice.models.zipvoice_distill import ZipVoiceDistill  # noqa: E402
from zipvoice.utils.common import make_pad_mask  # noqa: E402

CLS = dict(zipvoice=ZipVoice, zipvoice_distill=ZipVoiceDistill, zipvoice_dialog=ZipVoiceDialog,
           zipvoice_dialog_stereo=ZipVoiceDialogStereo)

sys.path.insert(0, os.path.join(ROOT, "tests"))
from fullsize_cases import CASES  # noqa: E402


def sd_checksum(sd) -> str:
    """Order-independent digest of a state_dict (used to check a rebuilt reference init)."""
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().contiguous().float().numpy().tobytes())
    return h.hexdigest()


def build(cfg, weights):
    if weights == "synth":
        model = CLS[cfg.variant](**cfg.model_kwargs())
        model.load_state_dict(synth_state_dict(cfg, 0), strict=True)
    else:
        torch.manual_seed(0)                     # seed, then construct the reference class
        model = CLS[cfg.variant](**cfg.model_kwargs())
    return model.eval()


@torch.inference_mode()
def run_case(name, case):
    t_begin = time.time()
    cfg = case["cfg"]
    model = build(cfg, case["weights"])
    u = synth_utterances(cfg, **case["ukw"])
    cat_tokens = [p + t for p, t in zip(u["prompt_tokens"], u["tokens"])]
    embed, tokens_lens = model.forward_text_embed(cat_tokens)                     # zipvoice.py:187-212
    text_condition, padding_mask = model.forward_text_condition(embed, tokens_lens, u["features_lens"])
    T = text_condition.shape[1]
    pf = u["prompt_features"]
    speech = torch.nn.functional.pad(pf, (0, 0, 0, T - pf.size(1)))               # zipvoice.py:445-451
    speech = torch.where(make_pad_mask(u["prompt_features_lens"], T).unsqueeze(-1), torch.zeros_like(speech), speech)
    keep = set(case["vel_steps"])
    stride = case["vel_stride"]
    vel, count = [], [0]

    def hook(m, i, o):
        if count[0] in keep:
            vel.append(o[:, ::stride].clone())
        count[0] += 1

    h = model.solver.model.register_forward_hook(hook)
    x1 = model.solver.sample(x=u["x0"], text_condition=text_condition, speech_condition=speech,
                             padding_mask=padding_mask, **case["skw"])
    h.remove()
    out = dict(variant=cfg.variant, weights=case["weights"], ukw=case["ukw"], skw=case["skw"],
               text_embed=embed, tokens_lens=tokens_lens, features_lens=u["features_lens"],
               text_condition_sum=text_condition.double().sum(dim=(1, 2)), padding_mask_lens=(~padding_mask).sum(-1),
               vel_steps=case["vel_steps"], vel_stride=stride, velocities=torch.stack(vel), x1=x1)
    if case["fm"]:      # one decoder forward through seam 1 (zipvoice.py:180-184)
        N = u["x0"].shape[0]
        xin = torch.cat([u["x0"], text_condition, speech], dim=2)
        t = torch.linspace(0.1, 0.9, N) if N > 1 else torch.tensor([0.3])
        kw = dict(guidance_scale=torch.full((N,), 2.0)) if cfg.is_distill else {}
        out["fm_in_t"] = t
        out["fm_out"] = model.fm_decoder(x=xin, t=t, padding_mask=padding_mask, **kw)[:, ::stride]
    if case["weights"] != "synth":
        sd = model.state_dict()
        out["sd_sha256"] = sd_checksum(sd)
        if case["weights"] == "refinit_store":
            out["state_dict"] = {k: v.clone() for k, v in sd.items()}
    out = {k: (v.contiguous().clone() if torch.is_tensor(v) else v) for k, v in out.items()}
    path = os.path.join(ROOT, "tests", "golden", name + ".pt")
    torch.save(out, path)
    print(f"{name}: T={T} x1 rms {float(x1.pow(2).mean().sqrt()):.4f} v rms "
          f"{float(out['velocities'].pow(2).mean().sqrt()):.4f} {os.path.getsize(path) / 1e6:.2f} MB "
          f"{time.time() - t_begin:.0f} s", flush=True)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 8)
    only = sys.argv[1:]
    for name, case in CASES.items():
        if only and name not in only:
            continue
        run_case(name, case)
