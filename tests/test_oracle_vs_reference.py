"""CPU: the oracle (oracle/zipvoice_oracle.py) against the LIVE, unmodified reference whenever its package is
importable (/root/reference in the build container, or the copy staged by tools/stage_reference.sh under
baseline/_ref), on inputs and weights the committed fixtures do not cover: other seeds, the reference's own
`torch.manual_seed(0)` initialisation, the public `model.sample` with the ratio-duration rule.  Also pins the
oracle at BASELINE.json's full length on the first CFG step of the C1 fixture.  Tolerance 2e-5 rel-L2 (fp32
re-association noise; SURVEY.md §8c measured 4.3e-7 between 1 and 8 threads of the reference itself)."""
import sys

import pytest
import torch

from oracle import zipvoice_oracle as orc
from zipvoice_b200.config import tiny_config
from zipvoice_b200.model import config_from_reference
from zipvoice_b200.synth import synth_state_dict, synth_utterances
from fullsize_cases import CASES
from fullsize_checks import reference_path
from util import load_golden, rel_l2

TOL = 2e-5
REF = reference_path()
needs_ref = pytest.mark.skipif(REF is None, reason="reference package not present (/root/reference or baseline/_ref)")


def _ref_class(variant):
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from zipvoice.models.zipvoice import ZipVoice
    from zipvoice.models.zipvoice_dialog import ZipVoiceDialog, ZipVoiceDialogStereo
    from zipvoice.models.zipvoice_distill import ZipVoiceDistill
    return dict(zipvoice=ZipVoice, zipvoice_distill=ZipVoiceDistill, zipvoice_dialog=ZipVoiceDialog,
                zipvoice_dialog_stereo=ZipVoiceDialogStereo)[variant]


@needs_ref
@pytest.mark.parametrize("variant", ["zipvoice", "zipvoice_distill", "zipvoice_dialog", "zipvoice_dialog_stereo"])
def test_oracle_matches_live_reference_on_its_own_init(variant):
    cfg = tiny_config(variant)
    torch.manual_seed(0)
    ref = _ref_class(variant)(**cfg.model_kwargs()).eval()
    assert config_from_reference(ref) == cfg
    sd = ref.state_dict()
    oracle = orc.OracleModel(cfg, sd)
    u = synth_utterances(cfg, batch=2, prompt_frames=21, target_frames=47, prompt_tokens=6, tokens=30, seed=99, ragged=True)
    kw = dict(num_step=3, guidance_scale=1.3, t_shift=0.5)
    with torch.inference_mode():
        tc, pm = ref.forward_text_inference_gt_duration(tokens=u["tokens"], features_lens=u["target_lens"],
                                                        prompt_tokens=u["prompt_tokens"],
                                                        prompt_features_lens=u["prompt_features_lens"])
        otc, osc, opm = oracle.prelude(u["tokens"], u["prompt_tokens"], u["prompt_features"], u["prompt_features_lens"],
                                       features_lens=u["target_lens"], duration="real")
        assert torch.equal(pm, opm) and rel_l2(otc, tc) < TOL
        want = ref.solver.sample(x=u["x0"], text_condition=tc, speech_condition=osc, padding_mask=pm, **kw)
        got = oracle.solve(u["x0"], otc, osc, opm, **kw)
    assert rel_l2(got, want) < TOL


@needs_ref
def test_oracle_sample_matches_reference_sample_api():
    """`model.sample` end to end (duration='predict', prompt/generated split) with the RNG draw pinned."""
    cfg = tiny_config("zipvoice")
    sd = synth_state_dict(cfg, 3)
    ref = _ref_class("zipvoice")(**cfg.model_kwargs()).eval()
    ref.load_state_dict(sd, strict=True)
    oracle = orc.OracleModel(cfg, sd)
    u = synth_utterances(cfg, batch=3, prompt_frames=25, target_frames=40, prompt_tokens=7, tokens=15, seed=5, ragged=True)
    kw = dict(speed=1.1, t_shift=0.7, duration="predict", num_step=2, guidance_scale=0.8)
    _, _, pm = oracle.prelude(u["tokens"], u["prompt_tokens"], u["prompt_features"], u["prompt_features_lens"],
                              speed=1.1, duration="predict")
    x0 = torch.randn(pm.shape[0], pm.shape[1], cfg.feat_dim, generator=torch.Generator().manual_seed(1))
    real_randn = torch.randn
    try:
        torch.randn = lambda *a, **k: x0.clone()                 # zipvoice.py:453 draws x0 with the global RNG
        with torch.inference_mode():
            want = ref.sample(u["tokens"], u["prompt_tokens"], u["prompt_features"], u["prompt_features_lens"], **kw)
    finally:
        torch.randn = real_randn
    got = oracle.sample(u["tokens"], u["prompt_tokens"], u["prompt_features"], u["prompt_features_lens"], x0=x0, **kw)
    assert torch.equal(got[1], want[1]) and torch.equal(got[3], want[3])
    assert rel_l2(got[0], want[0]) < TOL and rel_l2(got[2], want[2]) < TOL


def test_oracle_matches_the_full_length_fixture_first_step():
    """T = 1218 (C1): the oracle's first CFG-blended velocity against the reference's (fixture)."""
    name = "full_c1_zipvoice_16step"
    case, gold = CASES[name], load_golden(name)
    cfg = case["cfg"]
    oracle = orc.OracleModel(cfg, synth_state_dict(cfg, 0))
    u = synth_utterances(cfg, **gold["ukw"])
    tc, pm = orc.forward_text_condition(gold["text_embed"], gold["tokens_lens"], gold["features_lens"])
    T = tc.shape[1]
    pf = u["prompt_features"]
    speech = torch.nn.functional.pad(pf, (0, 0, 0, T - pf.size(1)))
    ts = orc.get_time_steps(0.0, 1.0, gold["skw"]["num_step"], gold["skw"]["t_shift"])
    with torch.inference_mode():
        v = orc.cfg_velocity(oracle.sd, oracle.fc, ts[0], u["x0"], tc, speech, pm, gold["skw"]["guidance_scale"], False)
    assert gold["vel_steps"][0] == 0 and gold["vel_stride"] == 1
    assert rel_l2(v, gold["velocities"][0]) < TOL
