"""CPU: the audio oracle (oracle/audio_oracle.py) against the committed torchaudio-generated vectors, against torchaudio
itself when importable, and torch.istft against the independent numpy overlap-add; host tables of the product's frontend
(mel filterbank, frame rule, speaker-cache bookkeeping) against the same sources."""
import numpy as np
import pytest
import torch

from oracle import audio_oracle as ao
from util import load_golden

FBANK_CASES = ["fbank_mono", "fbank_stereo", "fbank_short"]
# log-mel tolerance: torchaudio computes in fp32 (its own rounding on a bin ~60 dB below the frame energy is ~1e-4 in the
# log); frames clamped at log(1e-7) are exact
LOGMEL_ATOL = 2e-3


@pytest.mark.parametrize("name", FBANK_CASES)
def test_fbank_oracle_matches_reference_vectors(name):
    g = load_golden(name)
    wav, want = g["wav"], g["logmel"]
    got = np.concatenate([ao.vocos_fbank(wav[c].numpy()) for c in range(wav.shape[0])], axis=1)
    assert got.shape == tuple(want.shape)
    err = np.abs(got - want.numpy())
    assert err.max() < LOGMEL_ATOL, err.max()


def test_fbank_oracle_matches_live_torchaudio():
    ta = pytest.importorskip("torchaudio")
    g = torch.Generator().manual_seed(5)
    wav = (torch.randn(1, 7000, generator=g) * 0.1).float()
    fb = ta.transforms.MelSpectrogram(sample_rate=24000, n_fft=1024, hop_length=256, n_mels=100, center=True, power=1)
    want = fb(wav).clamp(min=1e-7).log()[0].t()[: ao.num_frames_for(7000)]
    got = ao.vocos_fbank(wav[0].numpy())
    assert np.abs(got - want.numpy()).max() < LOGMEL_ATOL
    assert np.abs(ao.mel_filterbank() - ta.functional.melscale_fbanks(513, 0.0, 12000.0, 100, 24000).numpy()).max() < 2e-5   # fp64 here, fp32 there


def test_frame_rule():
    # lhotse compute_num_frames: round-half-up of samples / hop in integer arithmetic
    assert [ao.num_frames_for(s) for s in (1, 127, 128, 255, 256, 383, 384, 72000)] == [0, 0, 1, 1, 1, 1, 2, 281]


def test_product_tables_match_oracle():
    from zipvoice_b200 import frontend as fe
    assert np.abs(fe.mel_filterbank().numpy() - ao.mel_filterbank()).max() < 2e-5
    assert all(fe.num_frames_for(s) == ao.num_frames_for(s) for s in range(1, 3000, 37))
    w = torch.randn(1000) * 0.01
    a, ra = fe.rms_norm(w, 0.1)
    b, rb = ao.rms_norm(w, 0.1)
    assert torch.equal(a, b) and torch.equal(ra, rb)


def test_istft_restatement_matches_torch():
    g = torch.Generator().manual_seed(3)
    T = 23
    spec = torch.complex(torch.randn(513, T, generator=g), torch.randn(513, T, generator=g))
    win = torch.hann_window(1024)
    want = torch.istft(spec.unsqueeze(0), 1024, 256, 1024, win, center=True)[0]
    got = ao.istft_numpy(spec.numpy().astype(np.complex128), win.numpy().astype(np.float64))
    assert got.shape[0] == 256 * (T - 1) == want.shape[0]
    assert np.abs(got - want.numpy()).max() < 1e-4 * float(want.abs().max())


def test_vocos_oracle_shapes_and_locality():
    from zipvoice_b200.vocoder import synth_vocos_state_dict
    sd = synth_vocos_state_dict(0, dim=256, intermediate=512, n_layers=2)
    mel = torch.randn(2, 100, 40)
    wav = ao.vocos_decode(sd, mel)
    assert wav.shape == (2, 256 * 39) and torch.isfinite(wav).all()
    # decoding is per utterance: row 0 does not depend on row 1
    assert torch.allclose(ao.vocos_decode(sd, mel[:1]), wav[:1], atol=1e-5)
    # checkpoint key set of vocos-mel-24khz (backbone.* / head.*)
    full = synth_vocos_state_dict(0)
    assert full["backbone.embed.weight"].shape == (512, 100, 7) and full["head.out.weight"].shape == (1026, 512)
    assert sum(k.endswith(".gamma") for k in full) == 8
