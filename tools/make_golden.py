#!/usr/bin/env python3
"""Generate tests/golden/*.pt by running the UNMODIFIED reference (build container only).

Usage:  PYTHONPATH=/root/reference PYTHONDONTWRITEBYTECODE=1 python tools/make_golden.py

For each case the reference model class is constructed, the seeded synthetic weights from
`zipvoice_b200.synth` are loaded with `strict=True`, and the reference's own
`model.sample`-level entry points are driven (`forward_text_*`, `solver.sample`,
`sample_intermediate`) on seeded synthetic inputs.  Only inputs that cannot be regenerated
from the seed and the outputs are stored, so the fixtures stay small.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from zipvoice_b200.config import ZipVoiceConfig, tiny_config  # noqa: E402
from zipvoice_b200.synth import synth_state_dict, synth_utterances  # noqa: E402

from zipvoice.models.zipvoice import ZipVoice  # noqa: E402
from zipvoice.models.zipvoice_dialog import ZipVoiceDialog, ZipVoiceDialogStereo  # noqa: E402
from zipvoice.models.zipvoice_distill import ZipVoiceDistill  # noqa: E402

CLS = dict(zipvoice=ZipVoice, zipvoice_distill=ZipVoiceDistill, zipvoice_dialog=ZipVoiceDialog,
           zipvoice_dialog_stereo=ZipVoiceDialogStereo)

# name -> (config, utterance kwargs, sampler kwargs)
CASES = {
    "tiny_zipvoice_cfg": (tiny_config("zipvoice"),
                          dict(batch=3, prompt_frames=23, target_frames=58, prompt_tokens=7,
                               tokens=19, ragged=True),
                          dict(num_step=4, guidance_scale=1.0, t_shift=0.5)),
    "tiny_zipvoice_g0": (tiny_config("zipvoice"),
                         dict(batch=2, prompt_frames=16, target_frames=37, prompt_tokens=5,
                              tokens=11),
                         dict(num_step=2, guidance_scale=0.0, t_shift=1.0)),
    "tiny_distill": (tiny_config("zipvoice_distill"),
                     dict(batch=2, prompt_frames=20, target_frames=45, prompt_tokens=6, tokens=14,
                          ragged=True),
                     dict(num_step=3, guidance_scale=3.0, t_shift=0.5)),
    "tiny_dialog": (tiny_config("zipvoice_dialog"),
                    dict(batch=2, prompt_frames=30, target_frames=71, prompt_tokens=9, tokens=52),
                    dict(num_step=3, guidance_scale=1.5, t_shift=0.5)),
    "tiny_stereo": (tiny_config("zipvoice_dialog_stereo"),
                    dict(batch=2, prompt_frames=18, target_frames=50, prompt_tokens=8, tokens=30,
                         ragged=True),
                    dict(num_step=3, guidance_scale=1.5, t_shift=0.5)),
    "base_zipvoice_cfg": (ZipVoiceConfig("zipvoice"),
                          dict(batch=2, prompt_frames=40, target_frames=93, prompt_tokens=10,
                               tokens=25, ragged=True),
                          dict(num_step=3, guidance_scale=1.0, t_shift=0.5)),
}


def build(cfg, seed=0):
    model = CLS[cfg.variant](**cfg.model_kwargs())
    model.load_state_dict(synth_state_dict(cfg, seed), strict=True)
    return model.eval()


@torch.inference_mode()
def run_case(name, cfg, ukw, skw):
    model = build(cfg)
    u = synth_utterances(cfg, **ukw)
    # prelude exactly as ZipVoice.sample does with duration="real" (zipvoice.py:431-451)
    text_condition, padding_mask = model.forward_text_inference_gt_duration(
        tokens=u["tokens"], features_lens=u["target_lens"], prompt_tokens=u["prompt_tokens"],
        prompt_features_lens=u["prompt_features_lens"])
    T = text_condition.shape[1]
    pf = u["prompt_features"]
    speech = torch.nn.functional.pad(pf, (0, 0, 0, T - pf.size(1)))
    from zipvoice.utils.common import make_pad_mask
    speech = torch.where(make_pad_mask(u["prompt_features_lens"], T).unsqueeze(-1),
                         torch.zeros_like(speech), speech)
    vel = []
    h = model.solver.model.register_forward_hook(lambda m, i, o: vel.append(o.clone()))
    x1 = model.solver.sample(x=u["x0"], text_condition=text_condition, speech_condition=speech,
                             padding_mask=padding_mask, **skw)
    h.remove()
    # the ratio-duration rule (zipvoice.py:290-330) on the same tokens
    tc_pred, pm_pred = model.forward_text_inference_ratio_duration(
        tokens=u["tokens"], prompt_tokens=u["prompt_tokens"],
        prompt_features_lens=u["prompt_features_lens"], speed=1.0)
    out = dict(text_condition=text_condition, padding_mask=padding_mask, speech_condition=speech,
               velocities=torch.stack(vel), x1=x1, pred_lens=(~pm_pred).sum(-1),
               pred_text_condition_sum=tc_pred.sum(dim=(1, 2)), ukw=ukw, skw=skw)
    # a single decoder forward through seam 1 (fm_decoder keyword call, zipvoice.py:180-184)
    N = u["x0"].shape[0]
    xin = torch.cat([u["x0"], text_condition, speech], dim=2)
    t = torch.linspace(0.1, 0.9, N)
    kw = dict(guidance_scale=torch.full((N,), 2.0)) if cfg.is_distill else {}
    out["fm_in_t"] = t
    out["fm_out"] = model.fm_decoder(x=xin, t=t, padding_mask=padding_mask, **kw)
    if name == "tiny_zipvoice_cfg":
        # sample_intermediate (zipvoice.py:488-534): tensor guidance (B,1,1), partial interval
        B = len(u["tokens"])
        g = torch.tensor([0.0, 0.7, 2.0])[:B].reshape(B, 1, 1)
        feats = torch.randn(B, T, cfg.feat_dim, generator=torch.Generator().manual_seed(5)) * 0.4
        scm = torch.arange(T)[None, :] >= u["prompt_features_lens"][:, None]
        xi, li = model.sample_intermediate(
            tokens=[p + t_ for p, t_ in zip(u["prompt_tokens"], u["tokens"])], features=feats,
            features_lens=u["features_lens"], noise=u["x0"], speech_condition_mask=scm,
            t_start=0.2, t_end=0.8, num_step=2, guidance_scale=g)
        out.update(si_features=feats, si_guidance=g, si_x=xi, si_lens=li)
    out = {k: (v.contiguous().clone() if torch.is_tensor(v) else v) for k, v in out.items()}
    torch.save(out, os.path.join(ROOT, "tests", "golden", name + ".pt"))
    print(name, "T", T, "x1 rms", float(x1.pow(2).mean().sqrt()),
          "v rms", float(out["velocities"].pow(2).mean().sqrt()))


if __name__ == "__main__":
    torch.set_num_threads(8)
    only = sys.argv[1:]
    for name, (cfg, ukw, skw) in CASES.items():
        if only and name not in only:
            continue
        run_case(name, cfg, ukw, skw)
