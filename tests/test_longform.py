"""CPU: the long-form helpers (zipvoice_b200/longform.py) against the reference's own functions (utils/infer.py) when the
package is importable, and through their invariants otherwise."""
import random
import sys
import types

import pytest
import torch

from fullsize_checks import reference_path
from zipvoice_b200 import longform as lf


def _ref_module():
    ref = reference_path()
    if ref is None:
        return None
    sys.path.insert(0, ref)
    for name in ("pydub", "pydub.silence"):            # imported at module level by utils/infer.py, absent offline
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["pydub"].AudioSegment = getattr(sys.modules["pydub"], "AudioSegment", object)
    sys.modules["pydub.silence"].detect_leading_silence = None
    sys.modules["pydub.silence"].split_on_silence = None
    try:
        import zipvoice.utils.infer as m
        return m
    except Exception:
        return None


def _text(seed, n):
    r = random.Random(seed)
    alphabet = list("abcdefg ") + [",", ".", "!", "?", "，", "。"]
    return [r.choice(alphabet) for _ in range(n)]


def test_chunkers_match_reference_or_invariants():
    ref = _ref_module()
    for seed in range(5):
        toks = _text(seed, 400)
        got = lf.chunk_tokens_punctuation(toks, max_tokens=60)
        assert [t for c in got for t in c] == toks
        if ref is not None:
            assert got == ref.chunk_tokens_punctuation(toks, max_tokens=60)
        r = random.Random(seed)
        dlg = []
        for _ in range(12):
            dlg += ["[S1]"] + [r.choice("abc") for _ in range(r.randint(1, 30))] + ["[S2]"] + [r.choice("xyz") for _ in range(r.randint(1, 30))]
        gd = lf.chunk_tokens_dialog(dlg, max_tokens=50)
        assert [t for c in gd for t in c] == dlg and all(c[0] == "[S1]" for c in gd)
        if ref is not None:
            assert gd == ref.chunk_tokens_dialog(dlg, max_tokens=50)


def _loop_cross_fade(chunks, fade_duration, sample_rate):
    """the reference's algorithm, restated (utils/infer.py:173-229) for boxes without the reference package"""
    final = chunks[0]
    fade_samples = int(fade_duration * sample_rate)
    for nxt in chunks[1:]:
        k = min(fade_samples, final.shape[-1], nxt.shape[-1])
        if k <= 0:
            final = torch.cat([final, nxt], dim=-1)
            continue
        fade = torch.linspace(1, 0, k)[None]
        final = torch.cat([final[..., :-k], final[..., -k:] * fade + nxt[..., :k] * (1 - fade), nxt[..., k:]], dim=-1)
    return final


@pytest.mark.parametrize("lens", [[5000, 7000, 300, 9000], [100, 50, 4000], [2400, 2400], [10, 20, 30, 40, 5000]])
def test_cross_fade_concat(lens):
    ref = _ref_module()
    g = torch.Generator().manual_seed(sum(lens))
    chunks = [torch.randn(1, n, generator=g) for n in lens]
    got = lf.cross_fade_concat(chunks, fade_duration=0.1, sample_rate=24000)
    want = _loop_cross_fade(chunks, 0.1, 24000)
    assert got.shape == want.shape and torch.allclose(got, want, atol=1e-6)
    if ref is not None:
        assert torch.allclose(got, ref.cross_fade_concat(chunks, fade_duration=0.1, sample_rate=24000), atol=1e-6)
    assert torch.equal(lf.cross_fade_concat(chunks, fade_duration=0.0), torch.cat(chunks, dim=-1))
    assert lf.cross_fade_concat(chunks[:1]) is chunks[0]


def test_generate_long_order_and_volume():
    class FakeModel:
        def sample(self, tokens, prompt_tokens, prompt_features, prompt_features_lens, duration, **kw):
            B = len(tokens)
            lens = torch.tensor([2 + len(t) for t in tokens])
            mel = torch.zeros(B, int(lens.max()), 100)
            for i, t in enumerate(tokens):
                mel[i, : lens[i]] = float(t[0])             # the chunk's first token marks its mel
            return mel, lens, prompt_features, prompt_features_lens

    class FakeVocoder:
        def decode_batch(self, mel, lens, scale=1.0, clamp=False):
            wav = mel[:, :-1, :1].repeat_interleave(4, dim=1).squeeze(-1) * scale * 0.01
            return wav, (lens - 1) * 4

    chunks = [[3, 1, 1, 1, 1], [1, 9], [2, 7, 7]]
    out = lf.generate_long(FakeModel(), FakeVocoder(), chunks, [5, 5], torch.zeros(7, 100), prompt_duration=1.0, token_duration=0.1,
                           prompt_rms=0.05, target_rms=0.1, fade_duration=0.0, feat_scale=0.1)
    # chunks come back in the caller's order (batchify sorts by length), each (len + 1) * 4 samples, scaled by 0.05 / 0.1
    marks = [3.0] * 24 + [1.0] * 12 + [2.0] * 16
    assert out.shape == (1, 52)
    assert torch.allclose(out[0], torch.tensor(marks) * 10.0 * 0.01 * 0.5)
