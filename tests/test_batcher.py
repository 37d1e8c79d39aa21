"""CPU: the length-sorted batcher (zipvoice_b200/batcher.py) -- `batchify_tokens` against the live reference
function when its package is importable, `plan_batches` through its invariants, `sample_batched` with a stand-in
model (order restored, zero padding, one `sample` call per batch)."""
import random
import sys

import pytest
import torch

from zipvoice_b200.batcher import batchify_tokens, padding_waste, plan_batches, sample_batched
from fullsize_checks import reference_path


def _tokens(n, seed):
    r = random.Random(seed)
    return [[r.randint(1, 300) for _ in range(r.randint(1, 120))] for _ in range(n)]


def test_batchify_tokens_budget_and_order():
    toks = _tokens(50, 0)
    batches, index = batchify_tokens(toks, max_duration=100.0, prompt_duration=3.0, token_duration=0.1)
    flat = [t for b in batches for t in b]
    assert flat == [toks[i] for i in index] and sorted(index) == list(range(50))
    assert [len(t) for t in flat] == sorted(len(t) for t in toks)
    for b in batches:
        n = sum(len(t) for t in b)
        assert len(b) == 1 or n * 0.1 + (len(b) - 1) * 3.0 <= 100.0 + 1e-9


def test_batchify_tokens_matches_the_reference_function():
    ref = reference_path()
    if ref is None:
        pytest.skip("reference package not present")
    sys.path.insert(0, ref)
    try:
        from zipvoice.utils.infer import batchify_tokens as ref_fn
    except Exception as e:           # utils/infer.py imports pydub at module level (absent offline)
        import importlib.util, types
        for name in ("pydub", "pydub.silence"):
            sys.modules.setdefault(name, types.ModuleType(name))
        sys.modules["pydub"].AudioSegment = object
        sys.modules["pydub.silence"].detect_leading_silence = None
        sys.modules["pydub.silence"].split_on_silence = None
        try:
            from zipvoice.utils.infer import batchify_tokens as ref_fn
        except Exception as e2:
            pytest.skip(f"reference utils/infer.py not importable offline: {e2}")
    for seed, (md, pd, td) in enumerate([(100.0, 3.0, 0.1), (30.0, 5.5, 0.07), (1.0, 2.0, 0.5)]):
        toks = _tokens(80, seed)
        assert batchify_tokens(toks, md, pd, td) == ref_fn(toks, md, pd, td)


def test_plan_batches_invariants():
    r = random.Random(3)
    total = [r.randint(881, 1219) for _ in range(512)]
    batches = plan_batches(total, max_rows=64, frame_bucket=64)
    assert sorted(i for b in batches for i in b) == list(range(512))
    assert all(len(b) <= 64 for b in batches) and len(batches) == 8
    firsts = [total[b[0]] for b in batches]
    assert firsts == sorted(firsts, reverse=True)                 # longest first
    assert all(total[b[0]] == max(total[i] for i in b) for b in batches)
    assert padding_waste(total, batches, 64) < 0.08
    assert padding_waste(total, [list(range(512))], 0) > 0.1      # one batch of everything pads far more
    assert [len(b) for b in plan_batches(total[:70], max_rows=64)] == [35, 35]        # no 6-row tail batch
    assert [len(b) for b in plan_batches(total[:130], max_rows=64)] == [44, 43, 43] and plan_batches([], max_rows=4) == []
    capped = plan_batches(total, max_rows=64, frame_bucket=64, max_batch_frames=20000)
    assert all(len(b) * ((max(total[i] for i in b) + 63) // 64 * 64) <= 20000 or len(b) == 1 for b in capped)


class _FakeModel:
    frame_bucket = 0

    def __init__(self):
        self.calls = []

    def sample(self, tokens, prompt_tokens, prompt_features, prompt_features_lens, features_lens=None, speed=1.0,
               duration="predict", **kw):
        self.calls.append(len(tokens))
        B = len(tokens)
        gl = features_lens
        mel = torch.zeros(B, int(gl.max()), 4)
        pm = torch.zeros(B, int(prompt_features_lens.max()), 4)
        for k in range(B):
            mel[k, : int(gl[k])] = float(tokens[k][0])            # tags every row with its utterance
            pm[k, : int(prompt_features_lens[k])] = float(tokens[k][0]) + 0.5
        return mel, gl.clone(), pm, prompt_features_lens.clone()


def test_sample_batched_restores_order_and_padding():
    U = 10
    tokens = [[100 + i] for i in range(U)]
    ptoks = [[1]] * U
    fl = torch.tensor([30, 12, 50, 44, 9, 28, 31, 50, 7, 19])
    pfl = torch.tensor([5, 6, 7, 5, 6, 7, 5, 6, 7, 5])
    pf = torch.zeros(U, 7, 4)
    m = _FakeModel()
    x1, l1, xp, lp = sample_batched(m, tokens, ptoks, pf, pfl, features_lens=fl, max_rows=4, num_step=2)
    assert m.calls == [4, 3, 3]                          # equal row counts, not 4 + 4 + 2
    assert torch.equal(l1, fl) and torch.equal(lp, pfl) and x1.shape == (U, 50, 4) and xp.shape == (U, 7, 4)
    for i in range(U):
        assert float(x1[i, : fl[i]].min()) == 100 + i and float(x1[i, fl[i]:].abs().sum()) == 0
        assert float(xp[i, : pfl[i]].min()) == 100.5 + i and float(xp[i, pfl[i]:].abs().sum()) == 0
