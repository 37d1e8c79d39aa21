"""-m gpu: the shape-bucketed plan / CUDA-graph cache.  Ragged traffic stays within a bounded
number of plans; a bucketed call equals the exact-shape call on the valid frames up to the down-sampling edge
the reference itself shows when a short utterance is batched with a longer one (zipformer.py:899-901): identical
away from the last few frames of the longest utterance, within the published 16-bit tolerance there; and after
plan eviction results are still correct (graphs die with their plan)."""
import random

import pytest
import torch

from zipvoice_b200.config import tiny_config
from zipvoice_b200.model import build_model
from zipvoice_b200.synth import synth_state_dict

pytestmark = pytest.mark.gpu


def _inputs(cfg, B, T, seed):
    g = torch.Generator().manual_seed(seed)
    F = cfg.feat_dim
    lens = torch.randint(max(8, int(T * 0.6)), T + 1, (B,), generator=g)
    lens[0] = T
    mask = torch.arange(T)[None, :] >= lens[:, None]
    x0 = torch.randn(B, T, F, generator=g)
    text = torch.randn(B, T, F, generator=g) * 0.5
    speech = torch.zeros(B, T, F)
    speech[:, :10] = torch.randn(B, 10, F, generator=g) * 0.3 - 0.5
    return dict(x=x0.cuda(), text_condition=text.cuda(), speech_condition=speech.cuda(), padding_mask=mask.cuda()), lens


def _pad_like_the_bucket(args, Bp, Tp):
    """What a caller of the reference would pass to batch these utterances with longer ones: zero inputs and
    masked frames / rows beyond the real ones."""
    B, T, F = args["x"].shape
    out = {}
    for k in ("x", "text_condition", "speech_condition"):
        out[k] = torch.zeros(Bp, Tp, args[k].shape[2], device="cuda")
        out[k][:B, :T] = args[k]
    out["padding_mask"] = torch.ones(Bp, Tp, dtype=torch.bool, device="cuda")
    out["padding_mask"][:B, :T] = args["padding_mask"]
    return out


def test_ragged_calls_stay_within_the_bucket_grid_and_equal_the_padded_exact_call():
    cfg = tiny_config("zipvoice")
    sd = synth_state_dict(cfg, 0)
    exact = build_model(cfg, sd, "cuda", use_cuda_graph=True)
    exact.solver.decoders[cfg.feat_dim].plans.max_plans = 3          # forces evictions on the exact-shape side
    buck = build_model(cfg, sd, "cuda", use_cuda_graph=True, frame_bucket=64, row_bucket=4)
    kw = dict(num_step=3, guidance_scale=1.0, t_shift=0.5)
    rnd = random.Random(1)
    worst = 0.0
    for i in range(20):
        B, T = rnd.randint(1, 6), rnd.randint(70, 250)
        args, lens = _inputs(cfg, B, T, 100 + i)
        b = buck.solver.sample(**args, **kw)
        assert b.shape == (B, T, cfg.feat_dim) and torch.isfinite(b).all()
        Bp, Tp = (B + 3) // 4 * 4, (T + 63) // 64 * 64
        # (1) plumbing: bit-identical to the exact-shape path fed with explicitly padded, masked inputs
        a = exact.solver.sample(**_pad_like_the_bucket(args, Bp, Tp), **kw)[:B, :T]
        assert torch.equal(a, b), (B, T)
        # (2) against the unpadded exact call only the down-sampling edge of the longest utterance differs
        # (zipformer.py:899-901: padded frames leak into the last low-rate frame, which every query attends to);
        # the shorter rows were padded in both runs and agree closely
        u = exact.solver.sample(**args, **kw)
        for r in range(1, B):
            n = int(lens[r])
            if n <= T - 4:
                worst = max(worst, float((u[r, :n] - b[r, :n]).norm() / u[r, :n].norm()))
    plans = buck.solver.decoders[cfg.feat_dim].plans
    assert plans.created <= 2 * 3           # rows {4, 8} (doubled for CFG) x frames {128, 192, 256}
    assert len(exact.solver.decoders[cfg.feat_dim].plans) <= 3
    assert worst <= 4e-3, worst             # final-state tolerance of BASELINE.md section 5


def test_bucketed_call_matches_the_oracle_on_the_padded_batch():
    """The bucketed result is the reference's result for the same utterances batched with longer ones."""
    from oracle import zipvoice_oracle as orc
    cfg = tiny_config("zipvoice")
    sd = synth_state_dict(cfg, 0)
    buck = build_model(cfg, sd, "cuda", use_cuda_graph=True, frame_bucket=64, row_bucket=2)
    oracle = orc.OracleModel(cfg, sd)
    kw = dict(num_step=3, guidance_scale=1.0, t_shift=0.5)
    args, lens = _inputs(cfg, 3, 101, 5)
    got = buck.solver.sample(**args, **kw)
    pad = _pad_like_the_bucket(args, 4, 128)
    want = oracle.solve(pad["x"].cpu(), pad["text_condition"].cpu(), pad["speech_condition"].cpu(),
                        pad["padding_mask"].cpu(), **kw)[:3, :101]
    for r in range(3):
        n = int(lens[r])
        rel = float((got[r, :n].cpu() - want[r, :n]).norm() / want[r, :n].norm())
        assert rel <= 4e-3, (r, rel)


def test_results_survive_plan_eviction_and_replay():
    """Same inputs before and after their plan (and its graphs) were evicted and rebuilt: bit-identical."""
    cfg = tiny_config("zipvoice")
    model = build_model(cfg, synth_state_dict(cfg, 0), "cuda", use_cuda_graph=True)
    plans = model.solver.decoders[cfg.feat_dim].plans
    plans.max_plans = 2
    kw = dict(num_step=2, guidance_scale=1.0, t_shift=0.5)
    args, _ = _inputs(cfg, 2, 90, 7)
    first = model.solver.sample(**args, **kw)
    again = model.solver.sample(**args, **kw)                       # graph replay
    for t in (100, 110, 120):                                       # pushes (4, 90) out of the cache
        other, _ = _inputs(cfg, 2, t, t)
        model.solver.sample(**other, **kw)
    assert all((p.N, p.T) != (4, 90) for p in plans.plans())
    rebuilt = model.solver.sample(**args, **kw)
    assert torch.equal(first, again) and torch.equal(first, rebuilt)


def test_text_encoder_token_bucket_is_exact():
    """The text encoder buckets its token axis by 32: valid positions are bit-identical to the exact shape."""
    cfg = tiny_config("zipvoice")
    sd = synth_state_dict(cfg, 0)
    a = build_model(cfg, sd, "cuda")
    b = build_model(cfg, sd, "cuda")
    b.text_encoder.plans.frame_bucket = 0
    toks = [[5, 9, 200, 31, 7], [17] * 23, [3, 4]]
    ea, la = a.forward_text_embed(toks)
    eb, lb = b.forward_text_embed(toks)
    assert torch.equal(la, lb)
    for i, n in enumerate(la.tolist()):
        assert torch.equal(ea[i, : n + 1], eb[i, : n + 1])
