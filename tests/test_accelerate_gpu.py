"""-m gpu: the reference-side binding itself.  An instance of the UNMODIFIED reference model classes (package
staged under baseline/_ref, or /root/reference in the build container) is patched in place by
`zipvoice_b200.accelerate()` -- the `load_trt`-style attribute swap (reference: zipvoice/utils/tensorrt.py:128-143)
-- and the reference's OWN `model.sample` / `model.sample_intermediate` code (zipvoice.py:388-534,
zipvoice_dialog.py:118-159) then runs on top of the CUDA path.  Checked against the oracle."""
import sys

import pytest
import torch

from oracle import zipvoice_oracle as orc
from zipvoice_b200.config import tiny_config
from zipvoice_b200.model import B200EulerSolver, B200Zipformer, accelerate
from zipvoice_b200.synth import synth_state_dict, synth_utterances
from fullsize_checks import reference_path
from util import max_abs, rel_l2

pytestmark = pytest.mark.gpu
REF = reference_path()
TOL_X_REL, TOL_X_ABS = 4e-3, 0.04


def _ref_model(variant, sd, cfg):
    if REF is None:
        pytest.skip("reference package not present (/root/reference or baseline/_ref)")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from zipvoice.models.zipvoice import ZipVoice
    from zipvoice.models.zipvoice_dialog import ZipVoiceDialog, ZipVoiceDialogStereo
    from zipvoice.models.zipvoice_distill import ZipVoiceDistill
    cls = dict(zipvoice=ZipVoice, zipvoice_distill=ZipVoiceDistill, zipvoice_dialog=ZipVoiceDialog,
               zipvoice_dialog_stereo=ZipVoiceDialogStereo)[variant]
    m = cls(**cfg.model_kwargs())
    m.load_state_dict(sd, strict=True)
    return m.eval().to("cuda")


@pytest.mark.parametrize("variant,guidance", [("zipvoice", 1.0), ("zipvoice_distill", 3.0), ("zipvoice_dialog_stereo", 1.5)])
def test_reference_sample_runs_on_the_cuda_path(variant, guidance, monkeypatch):
    cfg = tiny_config(variant)
    sd = synth_state_dict(cfg, 0)
    ref = accelerate(_ref_model(variant, sd, cfg))
    assert isinstance(ref.fm_decoder, (B200Zipformer,)) or type(ref.fm_decoder).__name__ == "_WidthDispatch"
    assert isinstance(ref.text_encoder, B200Zipformer) and isinstance(ref.solver, B200EulerSolver)
    u = synth_utterances(cfg, batch=3, prompt_frames=24, target_frames=61, prompt_tokens=7, tokens=26, seed=21, ragged=True)
    kw = dict(speed=1.0, t_shift=0.5, duration="predict", num_step=4, guidance_scale=guidance)
    oracle = orc.OracleModel(cfg, sd)
    _, _, pm = oracle.prelude(u["tokens"], u["prompt_tokens"], u["prompt_features"], u["prompt_features_lens"])
    F = cfg.feat_dim * (2 if cfg.is_stereo else 1)
    x0 = torch.randn(pm.shape[0], pm.shape[1], F, generator=torch.Generator().manual_seed(4))
    want = oracle.sample(u["tokens"], u["prompt_tokens"], u["prompt_features"], u["prompt_features_lens"], x0=x0, **kw)
    real_randn = torch.randn
    monkeypatch.setattr(torch, "randn", lambda *a, **k: x0.to("cuda") if k.get("device") is not None else real_randn(*a, **k))
    with torch.inference_mode():
        got = ref.sample(u["tokens"], u["prompt_tokens"], u["prompt_features"].cuda(), u["prompt_features_lens"].cuda(), **kw)
    monkeypatch.undo()
    assert torch.equal(got[1].cpu(), want[1]) and torch.equal(got[3].cpu(), want[3])
    for g, w in ((got[0], want[0]), (got[2], want[2])):
        assert g.shape == w.shape and rel_l2(g, w) <= TOL_X_REL and max_abs(g, w) <= TOL_X_ABS


def test_reference_sample_intermediate_runs_on_the_cuda_path():
    cfg = tiny_config("zipvoice")
    sd = synth_state_dict(cfg, 0)
    ref = accelerate(_ref_model("zipvoice", sd, cfg), use_cuda_graph=False)
    u = synth_utterances(cfg, batch=3, prompt_frames=24, target_frames=61, prompt_tokens=7, tokens=26, seed=22, ragged=True)
    T = int(u["features_lens"].max())
    g = torch.tensor([0.0, 0.7, 2.0]).reshape(3, 1, 1)
    feats = torch.randn(3, T, cfg.feat_dim, generator=torch.Generator().manual_seed(5)) * 0.4
    scm = torch.arange(T)[None, :] >= u["prompt_features_lens"][:, None]
    toks = [p + t for p, t in zip(u["prompt_tokens"], u["tokens"])]
    oracle = orc.OracleModel(cfg, sd)
    want, wl = oracle.sample_intermediate(toks, feats, u["features_lens"], u["x0"], scm, 0.2, 0.8, num_step=2, guidance_scale=g)
    with torch.inference_mode():
        got, gl = ref.sample_intermediate(tokens=toks, features=feats.cuda(), features_lens=u["features_lens"].cuda(),
                                          noise=u["x0"].cuda(), speech_condition_mask=scm.cuda(), t_start=0.2, t_end=0.8,
                                          num_step=2, guidance_scale=g.cuda())
    assert torch.equal(gl.cpu(), wl)
    assert rel_l2(got, want) <= TOL_X_REL and max_abs(got, want) <= TOL_X_ABS
