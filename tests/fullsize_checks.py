"""Parity of the CUDA path against the reference's fp32 results at BASELINE.json's full shapes and step
counts (fixtures tests/golden/full_*.pt from tools/make_golden_full.py; cases in fullsize_cases.py).
Same 16-bit thresholds as model_checks.py:
  text encoder output                   rel-L2 <= 3e-3
  decoder velocity (seam 1)             rel-L2 <= 3e-3, max-abs <= 0.03
  CFG-blended velocity of a step        rel-L2 <= 6e-3, max-abs <= 0.06
  final state x(t_end)                  rel-L2 <= 4e-3, max-abs <= 0.04
Saturation: the fp16 stores of the path saturate (`cvt.rn.satfinite`); every run also asserts that no
element of the residual stream or of the decoder's module-internal fp16 tensors reached the largest finite
fp16 value (zvb_plan_count_saturated over the plan's workspace)."""
from __future__ import annotations

import hashlib
import os
import sys

import torch

from zipvoice_b200.model import build_model
from zipvoice_b200.synth import synth_state_dict, synth_utterances
from fullsize_cases import CASES
from util import load_golden, max_abs, rel_l2
from model_checks import TOL_FM_ABS, TOL_FM_REL, TOL_TEXT_REL, TOL_V_ABS, TOL_V_REL, TOL_X_ABS, TOL_X_REL

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIRS = ("/root/reference", os.path.join(ROOT, "baseline", "_ref"))


def reference_path():
    for d in REF_DIRS:
        if os.path.isdir(os.path.join(d, "zipvoice", "models")):
            return d
    return None


def sd_checksum(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().contiguous().float().cpu().numpy().tobytes())
    return h.hexdigest()


def reference_init_state_dict(cfg):
    """`torch.manual_seed(0)`, then construct the reference class: needs the reference
    package (staged, unmodified, under baseline/_ref by tools/stage_reference.sh)."""
    ref = reference_path()
    if ref is None:
        return None
    if ref not in sys.path:
        sys.path.insert(0, ref)
    from zipvoice.models.zipvoice import ZipVoice
    from zipvoice.models.zipvoice_dialog import ZipVoiceDialog, ZipVoiceDialogStereo
    from zipvoice.models.zipvoice_distill import ZipVoiceDistill
    cls = dict(zipvoice=ZipVoice, zipvoice_distill=ZipVoiceDistill, zipvoice_dialog=ZipVoiceDialog,
               zipvoice_dialog_stereo=ZipVoiceDialogStereo)[cfg.variant]
    torch.manual_seed(0)
    return {k: v.detach().clone() for k, v in cls(**cfg.model_kwargs()).state_dict().items()}


def state_dict_for(case, gold):
    cfg = case["cfg"]
    if case["weights"] == "synth":
        return synth_state_dict(cfg, 0)
    if case["weights"] == "refinit_store":
        return gold["state_dict"]
    sd = reference_init_state_dict(cfg)
    if sd is None or sd_checksum(sd) != gold["sd_sha256"]:
        return None
    return sd


def run_case(name: str, use_cuda_graph: bool = True):
    case = CASES[name]
    cfg = case["cfg"]
    gold = load_golden(name)
    sd = state_dict_for(case, gold)
    if sd is None:
        return None
    u = synth_utterances(cfg, **gold["ukw"])
    model = build_model(cfg, sd, "cuda", use_cuda_graph=use_cuda_graph)
    model.solver.check_saturation = True          # every decoder plan counts saturated fp16 outputs, kernel by kernel
    dev = model.device
    res = {}
    cat_tokens = [p + t for p, t in zip(u["prompt_tokens"], u["tokens"])]
    embed, tokens_lens = model.forward_text_embed(cat_tokens)
    res["text_rel"] = rel_l2(embed, gold["text_embed"])
    # conditions rebuilt from the reference's text-encoder output (host logic only)
    tc, pm = model.forward_text_condition(gold["text_embed"].to(dev), gold["tokens_lens"].to(dev),
                                          gold["features_lens"].to(dev))
    res["tc_sum_err"] = float((tc.double().sum(dim=(1, 2)).cpu() - gold["text_condition_sum"]).abs().max())
    res["mask_equal"] = bool(torch.equal((~pm).sum(-1).cpu(), gold["padding_mask_lens"]))
    T = tc.shape[1]
    pf = u["prompt_features"].to(dev)
    speech = torch.nn.functional.pad(pf, (0, 0, 0, T - pf.size(1)))
    x0 = u["x0"].to(dev)
    stride = gold["vel_stride"]
    if "fm_out" in gold:
        xin = torch.cat([x0, tc, speech], dim=2)
        g = torch.full((xin.shape[0],), 2.0, device=dev) if cfg.is_distill else None
        out = model.fm_decoder(x=xin, t=gold["fm_in_t"].to(dev), padding_mask=pm, guidance_scale=g)[:, ::stride]
        res["fm_rel"], res["fm_abs"] = rel_l2(out, gold["fm_out"]), max_abs(out, gold["fm_out"])
        del xin, out
    model.solver.record_velocities = True
    x1 = model.solver.sample(x=x0, text_condition=tc, speech_condition=speech, padding_mask=pm, **gold["skw"])
    v = model.solver.last_velocities
    res["v_steps"] = list(gold["vel_steps"])
    res["v_rel"] = [rel_l2(v[s][:, ::stride], gold["velocities"][i]) for i, s in enumerate(gold["vel_steps"])]
    res["v_abs"] = [max_abs(v[s][:, ::stride], gold["velocities"][i]) for i, s in enumerate(gold["vel_steps"])]
    res["x_rel"], res["x_abs"] = rel_l2(x1, gold["x1"]), max_abs(x1, gold["x1"])
    res["finite"] = bool(torch.isfinite(x1).all())
    res["saturated"] = model.solver.count_saturated()
    res["sat_plans"] = sum(1 for d in model.solver.decoders.values() for p in d.plans.plans() if p._sat_counter is not None)
    log = os.environ.get("ZVB_PARITY_LOG")
    if log:
        import json
        with open(log, "a") as f:
            f.write(json.dumps(dict(case=name, **res)) + "\n")
    return res


def assert_case(name, res):
    assert res["finite"], (name, res)
    assert res["mask_equal"] and res["tc_sum_err"] < 1e-2, (name, res)
    assert res["text_rel"] <= TOL_TEXT_REL, (name, res)
    if "fm_rel" in res:
        assert res["fm_rel"] <= TOL_FM_REL and res["fm_abs"] <= TOL_FM_ABS, (name, res)
    assert max(res["v_rel"]) <= TOL_V_REL and max(res["v_abs"]) <= TOL_V_ABS, (name, res)
    assert res["x_rel"] <= TOL_X_REL and res["x_abs"] <= TOL_X_ABS, (name, res)
    assert res["sat_plans"] >= 1 and res["saturated"] == 0, (name, res)
