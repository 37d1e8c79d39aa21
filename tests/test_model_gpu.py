"""-m gpu: the whole CUDA path (text encoder, fm_decoder seam, Euler+CFG solver seam) against the
fixtures the reference produced; bf16 thresholds from BASELINE.md §5."""
import pytest

import model_checks as mc
from util import CASE_CFG

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(CASE_CFG))
def test_against_reference_fixture(name):
    mc.assert_case(name, mc.run_case(name))


def test_cuda_graph_replay_matches_eager():
    a = mc.run_case("tiny_zipvoice_cfg", use_cuda_graph=False)
    b = mc.run_case("tiny_zipvoice_cfg", use_cuda_graph=True)
    assert abs(a["x_rel"] - b["x_rel"]) < 1e-6, (a, b)


@pytest.mark.parametrize("switch", ["ZVB_NO_FUSED_PROLOGUE", "ZVB_NO_MERGE", "ZVB_NO_FAST_EPI", "ZVB_NO_PDL"])
def test_plan_switches_keep_parity(switch):
    """The A/B switches of csrc/engine.cu (`load_switches`) are read once per process: each legacy path is run in its own
    interpreter and must stay inside the same thresholds as the default plan."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, **{switch: "1"})
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "gpu_check.py"), "model:tiny_zipvoice_cfg",
                          "model:base_zipvoice_cfg"], env=env, capture_output=True, text=True, timeout=600)
    lines = [json.loads(ln) for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert out.returncode == 0 and len(lines) == 2 and all(r["ok"] for r in lines), (out.stdout[-2000:], out.stderr[-2000:])


@pytest.mark.parametrize("variant,B,T", [("zipvoice", 1, 1), ("zipvoice", 2, 3), ("zipvoice", 3, 17), ("zipvoice", 2, 37),
                                         ("zipvoice_distill", 2, 9), ("zipvoice_dialog_stereo", 2, 21)])
def test_extreme_short_shapes_match_the_oracle(variant, B, T):
    """Sequences shorter than a convolution kernel (31 taps), than the down-sampling factors (4) and than one tile:
    the reference's own edge handling (zipformer.py:899-901 pads the down-sampled tail, :1560-1590 convolves with zero
    padding) has to come out of the CUDA plan unchanged.  Ragged lengths where T allows."""
    import torch
    from oracle import zipvoice_oracle as orc
    from zipvoice_b200.config import tiny_config
    from zipvoice_b200.model import build_model
    from zipvoice_b200.synth import synth_state_dict
    cfg = tiny_config(variant)
    sd = synth_state_dict(cfg, 0)
    model = build_model(cfg, sd, "cuda", use_cuda_graph=False)
    oracle = orc.OracleModel(cfg, sd)
    g = torch.Generator().manual_seed(100 * B + T)
    F = cfg.feat_dim * (2 if cfg.is_stereo else 1)
    lens = torch.randint(max(1, T // 2), T + 1, (B,), generator=g)
    lens[0] = T
    mask = torch.arange(T)[None, :] >= lens[:, None]
    x0 = torch.randn(B, T, F, generator=g)
    text = torch.randn(B, T, cfg.feat_dim, generator=g) * 0.5
    speech = torch.zeros(B, T, F)
    speech[:, : max(1, T // 3)] = torch.randn(B, max(1, T // 3), F, generator=g) * 0.3 - 0.5
    kw = dict(num_step=2, guidance_scale=2.0 if cfg.is_distill else 1.0, t_shift=0.5)
    got = model.solver.sample(x=x0.cuda(), text_condition=text.cuda(), speech_condition=speech.cuda(),
                              padding_mask=mask.cuda(), **kw).cpu()
    want = oracle.solve(x0, text, speech, mask, **kw)
    assert torch.isfinite(got).all()
    for r in range(B):
        n = int(lens[r])
        rel = float((got[r, :n] - want[r, :n]).norm() / want[r, :n].norm())
        assert rel <= 4e-3, (variant, B, T, r, rel)
