"""-m gpu: the public model API (`sample`, `sample_intermediate`, `forward_fm_decoder`) against the
fixtures produced by the reference and against the oracle, including the ratio-duration rule, the
prompt/generated split and tensor-valued guidance."""
import pytest
import torch

from oracle import zipvoice_oracle as orc
from zipvoice_b200.model import build_model
from zipvoice_b200.synth import synth_state_dict, synth_utterances
from util import CASE_CFG, load_golden, max_abs, rel_l2

pytestmark = pytest.mark.gpu
TOL_X_REL, TOL_X_ABS = 4e-3, 0.04


@pytest.fixture(scope="module")
def tiny():
    cfg = CASE_CFG["tiny_zipvoice_cfg"]()
    gold = load_golden("tiny_zipvoice_cfg")
    sd = synth_state_dict(cfg, 0)
    return cfg, gold, sd, build_model(cfg, sd, "cuda", use_cuda_graph=True), synth_utterances(cfg, **gold["ukw"])


def test_sample_intermediate_tensor_guidance(tiny):
    cfg, gold, sd, model, u = tiny
    T = gold["text_condition"].shape[1]
    scm = torch.arange(T)[None, :] >= u["prompt_features_lens"][:, None]
    x, lens = model.sample_intermediate(
        tokens=[p + t for p, t in zip(u["prompt_tokens"], u["tokens"])], features=gold["si_features"],
        features_lens=u["features_lens"], noise=u["x0"], speech_condition_mask=scm, t_start=0.2, t_end=0.8,
        num_step=2, guidance_scale=gold["si_guidance"])
    assert torch.equal(lens.cpu(), gold["si_lens"])
    assert rel_l2(x, gold["si_x"]) <= TOL_X_REL and max_abs(x, gold["si_x"]) <= TOL_X_ABS


def test_sample_predict_duration_and_split(tiny):
    cfg, gold, sd, model, u = tiny
    kw = dict(speed=1.0, t_shift=0.5, duration="predict", num_step=3, guidance_scale=1.0)
    oracle = orc.OracleModel(cfg, sd)
    # x0 must cover the predicted length: draw it for the oracle's predicted number of frames
    tc, _, pm = oracle.prelude(u["tokens"], u["prompt_tokens"], u["prompt_features"], u["prompt_features_lens"])
    x0 = torch.randn(pm.shape[0], pm.shape[1], cfg.feat_dim, generator=torch.Generator().manual_seed(11))
    want = oracle.sample(u["tokens"], u["prompt_tokens"], u["prompt_features"], u["prompt_features_lens"], x0=x0, **kw)
    got = model.sample(u["tokens"], u["prompt_tokens"], u["prompt_features"], u["prompt_features_lens"], x0=x0, **kw)
    assert torch.equal(got[1].cpu(), want[1]) and torch.equal(got[3].cpu(), want[3])
    assert torch.equal((~pm).sum(-1), gold["pred_lens"])
    for g, w in ((got[0], want[0]), (got[2], want[2])):
        assert g.shape == w.shape
        assert rel_l2(g, w) <= TOL_X_REL and max_abs(g, w) <= TOL_X_ABS
    # zero padding beyond each utterance's length, as the reference's zero-initialised outputs
    for b in range(got[0].shape[0]):
        assert float(got[0][b, int(got[1][b]):].abs().sum()) == 0.0


def test_forward_fm_decoder_scalar_and_batched_t(tiny):
    cfg, gold, sd, model, u = tiny
    dev = model.device
    x = u["x0"].to(dev)
    tc, sc, pm = gold["text_condition"].to(dev), gold["speech_condition"].to(dev), gold["padding_mask"].to(dev)
    t = torch.tensor(0.37, device=dev)
    a = model.forward_fm_decoder(t=t, xt=x, text_condition=tc, speech_condition=sc, padding_mask=pm)
    b = model.forward_fm_decoder(t=t.repeat(x.shape[0]).reshape(-1, 1, 1), xt=x, text_condition=tc,
                                 speech_condition=sc, padding_mask=pm)
    assert torch.equal(a, b)
    oracle = orc.OracleModel(cfg, sd)
    want = orc.forward_fm_decoder(oracle.sd, oracle.fc, torch.tensor(0.37), u["x0"], gold["text_condition"],
                                  gold["speech_condition"], gold["padding_mask"])
    assert rel_l2(a, want) <= 3e-3


def test_guidance_zero_takes_the_single_pass_path(tiny):
    cfg, gold, sd, model, u = tiny
    dev = model.device
    args = dict(x=u["x0"].to(dev), text_condition=gold["text_condition"].to(dev),
                speech_condition=gold["speech_condition"].to(dev), padding_mask=gold["padding_mask"].to(dev),
                num_step=2, t_shift=0.5)
    a = model.solver.sample(guidance_scale=0.0, **args)
    b = model.solver.sample(guidance_scale=torch.zeros(3, 1, 1), **args)
    assert torch.equal(a, b)
