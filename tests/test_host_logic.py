"""CPU: host-side bookkeeping of the product (zipvoice_b200.model) against the oracle's restatement
of the reference's Python loops, the schedule, synthetic weights vs the reference's key set."""
import os

import pytest
import torch

from oracle import zipvoice_oracle as orc
from zipvoice_b200 import model as M
from zipvoice_b200.config import VARIANTS, ZipVoiceConfig, tiny_config
from zipvoice_b200.sharding import partition_utterances
from zipvoice_b200.synth import synth_state_dict, synth_utterances


def test_tokens_index_matches_reference_loops():
    g = torch.Generator().manual_seed(0)
    for _ in range(20):
        B = int(torch.randint(1, 6, (1,), generator=g))
        tl = torch.randint(1, 40, (B,), generator=g)
        fl = torch.randint(1, 300, (B,), generator=g)
        T = int(fl.max())
        assert torch.equal(M.tokens_index(fl, tl, T), orc.tokens_index(fl, tl, T))
    # more tokens than frames: every frame points at the appended pad token
    fl, tl = torch.tensor([3]), torch.tensor([7])
    assert torch.equal(M.tokens_index(fl, tl, 3), orc.tokens_index(fl, tl, 3))


def test_pad_labels_and_masks():
    y = [[5, 6, 7], [], [9]]
    assert torch.equal(M.pad_labels(y, 0, "cpu"), orc.pad_labels(y, 0))
    lens = torch.tensor([3, 1, 5])
    assert torch.equal(M.make_pad_mask(lens, 4), orc.make_pad_mask(lens, 4))
    assert torch.equal(M.make_pad_mask(lens, 7), orc.make_pad_mask(lens, 7))


@pytest.mark.parametrize("args", [(0.0, 1.0, 16, 0.5), (0.2, 0.8, 2, 1.0), (0.0, 1.0, 4, 0.3)])
def test_time_grid(args):
    a, b = M.get_time_steps(*args), orc.get_time_steps(*args)
    assert torch.equal(a, b)
    if args == (0.0, 1.0, 16, 0.5):
        # reference: steps 0-10 have t <= 0.5, steps 11-15 t > 0.5
        assert [bool(v > 0.5) for v in a[:-1]] == [False] * 11 + [True] * 5


def test_duration_rule():
    pfl, pl, tl = torch.tensor([281, 100]), torch.tensor([45, 13]), torch.tensor([150, 7])
    assert orc.predict_features_lens(pfl, pl, tl, 1.0).tolist() == [281 + 937, 100 + 54]


@pytest.mark.parametrize("variant", VARIANTS)
def test_synth_state_dict_is_deterministic_and_complete(variant):
    cfg = tiny_config(variant)
    a, b = synth_state_dict(cfg, 3), synth_state_dict(cfg, 3)
    assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)
    assert ("spk_embed.weight" in a) == cfg.is_dialog
    assert ("fm_decoder.guidance_scale_embed.weight" in a) == cfg.is_distill
    assert ("fm_decoder.in_proj.1.weight" in a) == cfg.is_stereo


@pytest.mark.skipif(not os.path.isdir("/root/reference/zipvoice"), reason="reference not mounted")
@pytest.mark.parametrize("variant", VARIANTS)
def test_synth_keys_and_shapes_equal_the_reference(variant):
    import logging
    import sys
    logging.disable(logging.WARNING)
    sys.path.insert(0, "/root/reference")
    sys.dont_write_bytecode = True
    from zipvoice.models.zipvoice import ZipVoice
    from zipvoice.models.zipvoice_dialog import ZipVoiceDialog, ZipVoiceDialogStereo
    from zipvoice.models.zipvoice_distill import ZipVoiceDistill
    cls = dict(zipvoice=ZipVoice, zipvoice_distill=ZipVoiceDistill, zipvoice_dialog=ZipVoiceDialog,
               zipvoice_dialog_stereo=ZipVoiceDialogStereo)[variant]
    cfg = ZipVoiceConfig(variant, vocab_size=362 if "dialog" in variant else 360)
    ref = cls(**cfg.model_kwargs()).state_dict()
    mine = synth_state_dict(cfg)
    assert set(ref) == set(mine)
    for k in ref:
        assert tuple(ref[k].shape) == tuple(mine[k].shape), k


def test_partition_balances_frames():
    lens = [1219, 900, 1200, 640, 700, 1100, 1000, 650, 800]
    for world in (1, 2, 4, 8):
        shards = partition_utterances(lens, world)
        assert sorted(i for s in shards for i in s) == list(range(len(lens)))
        loads = [sum(lens[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(lens)
    assert partition_utterances(lens, 2) == partition_utterances(list(lens), 2)     # deterministic


def test_ragged_synthetic_batch_is_sorted_and_consistent():
    cfg = tiny_config()
    u = synth_utterances(cfg, batch=5, prompt_frames=20, target_frames=50, ragged=True)
    assert u["target_lens"].tolist() == sorted(u["target_lens"].tolist(), reverse=True)
    assert int(u["features_lens"].max()) == u["x0"].shape[1]
    assert float(u["prompt_features"][1, int(u["prompt_features_lens"][1]):].abs().sum()) == 0.0


def test_pos_table_pair_layout():
    """weights.pack_pos_table: fp16 column pairs {log2e*E[r][d], log2e*E[r+1][d]} at entry r + POS_PAD, zeros
    outside the 2L-1 offsets, followed by max_r |E[h][r]|_2 per head (the layout include/zipvoice_b200.h
    documents for zvb_layer.pos_table)."""
    from zipvoice_b200.weights import POS_PAD, pack_pos_table
    H, L = 3, 37
    R = 2 * L - 1
    g = torch.Generator().manual_seed(3)
    e = torch.randn(H, R, 4, generator=g)
    raw = pack_pos_table(e)
    n_entries = R + 2 * POS_PAD
    assert raw.dtype == torch.uint8 and raw.numel() == H * n_entries * 16 + H * 4
    pairs = raw[: H * n_entries * 16].view(torch.float16).reshape(H, n_entries, 4, 2).float()
    emax = raw[H * n_entries * 16:].view(torch.float32)
    want = torch.zeros(H, n_entries + 1, 4)
    want[:, POS_PAD:POS_PAD + R] = e * 1.4426950408889634
    assert torch.allclose(pairs[..., 0], want[:, :-1], rtol=1e-3, atol=1e-4)       # fp16 rounding
    assert torch.allclose(pairs[..., 1], want[:, 1:], rtol=1e-3, atol=1e-4)
    assert float(pairs[:, :POS_PAD - 1].abs().max()) == 0.0 and float(pairs[:, POS_PAD + R:].abs().max()) == 0.0
    assert torch.allclose(emax, e.norm(dim=2).amax(dim=1))
    # every 255-entry window a score tile can ask for lies inside the table
    for i0 in range(0, L, 128):
        for j0 in range(0, L, 128):
            start = (j0 - i0) - 127 + (L - 1) + POS_PAD
            assert 0 <= start and start + 255 <= n_entries
