"""The oracle (oracle/zipvoice_oracle.py) against fixtures produced by the reference itself
(tools/make_golden.py).  Tolerance: 2e-5 rel-L2 -- fp32 re-association noise; the reference's
own 1-thread vs 8-thread runs differ by 4.3e-7."""
import pytest
import torch

from oracle import zipvoice_oracle as orc
from zipvoice_b200.synth import synth_state_dict, synth_utterances
from util import CASE_CFG, load_golden, rel_l2

TOL = 2e-5
CASES = list(CASE_CFG)


@pytest.fixture(scope="module", params=CASES)
def case(request):
    name = request.param
    cfg = CASE_CFG[name]()
    gold = load_golden(name)
    model = orc.OracleModel(cfg, synth_state_dict(cfg, 0))
    u = synth_utterances(cfg, **gold["ukw"])
    return name, cfg, gold, model, u


def test_prelude_matches_reference(case):
    name, cfg, gold, model, u = case
    tc, sc, pm = model.prelude(u["tokens"], u["prompt_tokens"], u["prompt_features"],
                               u["prompt_features_lens"], features_lens=u["target_lens"],
                               duration="real")
    assert torch.equal(pm, gold["padding_mask"])
    assert rel_l2(tc, gold["text_condition"]) < TOL
    assert torch.equal(sc, gold["speech_condition"])
    # ratio-duration rule
    tc2, _, pm2 = model.prelude(u["tokens"], u["prompt_tokens"], u["prompt_features"],
                                u["prompt_features_lens"], duration="predict")
    assert torch.equal((~pm2).sum(-1), gold["pred_lens"])
    assert rel_l2(tc2.sum(dim=(1, 2)), gold["pred_text_condition_sum"]) < 1e-4


def test_solver_velocities_and_final_state(case):
    name, cfg, gold, model, u = case
    rec = []
    x1 = model.solve(u["x0"], gold["text_condition"], gold["speech_condition"],
                     gold["padding_mask"], record=rec, **gold["skw"])
    v = torch.stack(rec)
    assert v.shape == gold["velocities"].shape
    for i in range(v.shape[0]):
        assert rel_l2(v[i], gold["velocities"][i]) < TOL, (name, i)
    assert rel_l2(x1, gold["x1"]) < TOL


def test_fm_decoder_seam(case):
    name, cfg, gold, model, u = case
    xin = torch.cat([u["x0"], gold["text_condition"], gold["speech_condition"]], dim=2)
    g = torch.full((xin.shape[0],), 2.0) if cfg.is_distill else None
    out = orc.tts_zipformer(model.sd, "fm_decoder.", model.fc, xin, gold["fm_in_t"],
                            gold["padding_mask"], g)
    assert rel_l2(out, gold["fm_out"]) < TOL


def test_sample_intermediate():
    name = "tiny_zipvoice_cfg"
    cfg = CASE_CFG[name]()
    gold = load_golden(name)
    model = orc.OracleModel(cfg, synth_state_dict(cfg, 0))
    u = synth_utterances(cfg, **gold["ukw"])
    T = gold["text_condition"].shape[1]
    scm = torch.arange(T)[None, :] >= u["prompt_features_lens"][:, None]
    x, lens = model.sample_intermediate(
        [p + t for p, t in zip(u["prompt_tokens"], u["tokens"])], gold["si_features"],
        u["features_lens"], u["x0"], scm, 0.2, 0.8, num_step=2, guidance_scale=gold["si_guidance"])
    assert torch.equal(lens, gold["si_lens"])
    assert rel_l2(x, gold["si_x"]) < TOL


def test_sample_split_round_trip():
    """sample() = prelude + solve + split; prompt part / generated part re-assemble to x1."""
    cfg = CASE_CFG["tiny_zipvoice_cfg"]()
    gold = load_golden("tiny_zipvoice_cfg")
    model = orc.OracleModel(cfg, synth_state_dict(cfg, 0))
    u = synth_utterances(cfg, **gold["ukw"])
    out, lens, pr, pl = model.sample(u["tokens"], u["prompt_tokens"], u["prompt_features"],
                                     u["prompt_features_lens"], features_lens=u["target_lens"],
                                     duration="real", x0=u["x0"], **gold["skw"])
    assert torch.equal(lens, u["target_lens"])
    for b in range(out.shape[0]):
        p, g = int(pl[b]), int(lens[b])
        assert rel_l2(pr[b, :p], gold["x1"][b, :p]) < TOL
        assert rel_l2(out[b, :g], gold["x1"][b, p:p + g]) < TOL
        assert float(out[b, g:].abs().max() if g < out.shape[1] else 0.0) == 0.0
