"""-m gpu: the stages either side of the sampler through the C ABI:
prompt log-mel (zvb_fbank) against the torchaudio-generated vectors and live torchaudio, the speaker cache, and the
Vocos vocoder (zvb_vocoder_decode) against the CPU oracle (oracle/audio_oracle.py), ragged batches included."""
import numpy as np
import pytest
import torch

from oracle import audio_oracle as ao
from util import load_golden, rel_l2
from zipvoice_b200.frontend import SpeakerCache, VocosFbank, num_frames_for, rms_norm
from zipvoice_b200.vocoder import Vocos, synth_vocos_state_dict

pytestmark = pytest.mark.gpu

# fp32 FFT + fp32 filterbank on both sides: the log of a mel bin ~60 dB under the frame energy moves by ~1e-4
LOGMEL_ATOL = 2e-3
# 16-bit backbone (fp16 operands, fp32 accumulate), fp32 inverse STFT: thresholds on the waveform
VOC_REL, VOC_MAXABS = 6e-3, 1e-2          # measured on B200: 1.9e-3 / 0.9e-2 of the peak -> see BASELINE.md


@pytest.mark.parametrize("name", ["fbank_mono", "fbank_stereo", "fbank_short"])
def test_fbank_matches_reference_vectors(name):
    g = load_golden(name)
    wav, want = g["wav"], g["logmel"]
    fe = VocosFbank(num_channels=wav.shape[0])
    got = fe.extract(wav, 24000)
    assert got.shape == want.shape and got.device.type == "cuda"
    err = (got.cpu() - want).abs()
    assert float(err.max()) < LOGMEL_ATOL, float(err.max())
    # the numpy path of the reference's extractor returns numpy
    assert isinstance(fe.extract(wav.numpy(), 24000), np.ndarray)


def test_fbank_batch_ragged_vs_torchaudio_and_oracle():
    ta = pytest.importorskip("torchaudio")
    g = torch.Generator().manual_seed(7)
    lens = torch.tensor([72000, 30017, 5000, 128])          # 3 s prompt, ragged, one-frame utterance
    wavs = torch.zeros(4, int(lens.max()))
    for i, n in enumerate(lens.tolist()):
        wavs[i, :n] = torch.randn(n, generator=g) * (0.02 + 0.1 * i)
    fe = VocosFbank()
    feats, frames = fe.extract_batch(wavs, lens, scale=0.1)
    assert frames.tolist() == [num_frames_for(n) for n in lens.tolist()] == [281, 117, 20, 1]
    fb = ta.transforms.MelSpectrogram(sample_rate=24000, n_fft=1024, hop_length=256, n_mels=100, center=True, power=1)
    for i, n in enumerate(lens.tolist()):
        T = int(frames[i])
        if n > 512:                                          # torch's reflect pad needs more samples than the pad
            want = fb(wavs[i:i + 1, :n]).clamp(min=1e-7).log()[0].t()[:T] * 0.1
            assert float((feats[i, :T].cpu() - want).abs().max()) < LOGMEL_ATOL * 0.1
            assert np.abs(feats[i, :T].cpu().numpy() - 0.1 * ao.vocos_fbank(wavs[i, :n].numpy())).max() < LOGMEL_ATOL * 0.1
        assert float(feats[i, T:].abs().max() if T < feats.shape[1] else 0.0) == 0.0


def test_speaker_cache():
    g = torch.Generator().manual_seed(9)
    fe = VocosFbank()
    cache = SpeakerCache(fe, max_speakers=2)
    wav_a = torch.randn(24000, generator=g) * 0.01           # quieter than target_rms: scaled up
    wav_b = torch.randn(2, 12000, generator=g) * 0.3         # stereo prompt: averaged
    ta, fa, ra = cache.get("a", wav_a, [1, 2, 3])
    want, _ = rms_norm(wav_a, 0.1)
    ref = torch.from_numpy(ao.vocos_fbank(want.numpy())).float() * 0.1
    assert fa.shape == ref.shape and float((fa.cpu() - ref).abs().max()) < LOGMEL_ATOL * 0.1
    assert abs(ra - float(wav_a.square().mean().sqrt())) < 1e-7 and ta == [1, 2, 3]
    cache.get("b", wav_b, [4])
    assert cache.get("a")[1] is fa and cache.hits == 1       # served from the cache, no re-extraction
    toks, feats, lens, rms = cache.batch(["a", "b"])
    assert feats.shape[0] == 2 and lens.tolist() == [num_frames_for(24000), num_frames_for(12000)] and toks == [[1, 2, 3], [4]]
    cache.get("a")                                           # "b" is now the least recently used entry
    cache.get("c", wav_a, [5])
    assert "b" not in cache and "a" in cache and len(cache) == 2     # least recently used entry dropped
    with pytest.raises(KeyError):
        cache.get("zzz")


def _mel(B, T, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 100, T, generator=g) * 2.0 - 4.0    # log-mel range of speech


def test_vocoder_decode_matches_oracle():
    sd = synth_vocos_state_dict(0)
    voc = Vocos().load_state_dict(sd).to("cuda").eval()
    mel = _mel(2, 150, 1)
    got = voc.decode(mel.cuda())
    want = ao.vocos_decode(sd, mel)
    assert got.shape == want.shape == (2, 256 * 149)
    r, m = rel_l2(got, want), float((got.cpu() - want).abs().max())
    print(f"vocoder decode: rel-L2 {r:.2e}, max-abs {m:.2e} (rms {float(want.square().mean().sqrt()):.3f})")
    assert r < VOC_REL and m < VOC_MAXABS * float(want.abs().max())


def test_vocoder_ragged_batch_equals_single_utterance_decodes():
    sd = synth_vocos_state_dict(1)
    voc = Vocos(frame_bucket=64).load_state_dict(sd).to("cuda").eval()
    lens = torch.tensor([333, 128, 47, 2])
    melT = torch.zeros(4, 333, 100)
    full = _mel(4, 333, 2).permute(0, 2, 1)
    for i, n in enumerate(lens.tolist()):
        melT[i, :n] = full[i, :n] * 0.1                       # feat-scaled, as model.sample returns it
    wav, wl = voc.decode_batch(melT.cuda(), lens.cuda(), scale=10.0, clamp=True)
    assert wl.tolist() == [(n - 1) * 256 for n in lens.tolist()]
    for i, n in enumerate(lens.tolist()):
        want = ao.vocos_decode(sd, (melT[i:i + 1, :n] * 10.0).permute(0, 2, 1))[0].clamp(-1, 1)
        got = wav[i, : want.numel()].cpu()
        assert rel_l2(got, want) < VOC_REL, (i, rel_l2(got, want))
        assert float(wav[i, want.numel():].abs().max() if want.numel() < wav.shape[1] else 0.0) == 0.0
    # the same utterance alone (another plan shape) gives the same samples as inside the batch
    alone, _ = voc.decode_batch(melT[1:2, :128].cuda(), lens[1:2].cuda(), scale=10.0, clamp=True)
    assert float((alone[0] - wav[1, : alone.shape[1]]).abs().max()) < 2e-3


def test_vocoder_full_pipeline_shapes():
    """wav -> fbank -> (identity sampler stand-in) -> vocoder: the glue a caller writes (infer_zipvoice.py:372-409)."""
    fe = VocosFbank()
    voc = Vocos().load_state_dict(synth_vocos_state_dict(2)).to("cuda")
    g = torch.Generator().manual_seed(4)
    wav = torch.randn(1, 24000, generator=g) * 0.05
    feats = fe.extract(wav, 24000).unsqueeze(0) * 0.1          # (1, T, 100) feat-scaled
    out = voc.decode((feats / 0.1).permute(0, 2, 1)).clamp(-1, 1)
    assert out.shape == (1, 256 * (feats.shape[1] - 1)) and torch.isfinite(out).all()


def test_generate_long_end_to_end():
    """Long-form driver over the real stages: batchify -> model.sample -> vocoder.decode_batch -> cross-fade equals the
    same chunks sampled and decoded one by one (same noise through a fixed generator seed)."""
    from zipvoice_b200.config import tiny_config
    from zipvoice_b200.longform import cross_fade_concat, generate_long
    from zipvoice_b200.model import build_model
    from zipvoice_b200.synth import synth_state_dict
    cfg = tiny_config("zipvoice")
    model = build_model(cfg, synth_state_dict(cfg, 0), "cuda", use_cuda_graph=False)
    voc = Vocos().load_state_dict(synth_vocos_state_dict(3, dim=256, intermediate=512, n_layers=2)).to("cuda")
    g = torch.Generator().manual_seed(11)
    prompt_feats = (torch.randn(24, 100, generator=g) * 0.3 - 0.5).cuda()
    prompt_tokens = [3, 4, 5, 6, 7, 8]
    chunks = [[9, 10, 11, 12, 13, 14, 15, 16, 17], [20, 21, 22], [30, 31, 32, 33, 34]]
    kw = dict(num_step=2, guidance_scale=1.0, t_shift=0.5)
    torch.manual_seed(5)
    wav = generate_long(model, voc, chunks, prompt_tokens, prompt_feats, prompt_duration=0.26, token_duration=0.04,
                        prompt_rms=0.05, target_rms=0.1, max_duration=100.0, **kw)
    assert wav.dim() == 2 and wav.shape[0] == 1 and torch.isfinite(wav).all() and float(wav.abs().max()) <= 0.5 + 1e-6
    # expected number of samples: every chunk gives 256 * (frames - 1) samples, joined with 0.1 s cross-fades
    lens = [int(torch.ceil(torch.tensor(24 / 6 * len(c))).item()) for c in chunks]
    parts = [torch.zeros(1, 256 * (n - 1)) for n in lens]
    assert wav.shape[1] == cross_fade_concat(parts, 0.1, 24000).shape[1]


def test_fbank_properties():
    """Size-independent properties at the bench shape: silence clamps to log(1e-7) exactly, scaling the waveform by c adds
    log(c) to every unclamped log-mel, and an utterance's features do not depend on its batch neighbours."""
    import math
    fe = VocosFbank()
    g = torch.Generator().manual_seed(21)
    S = 281 * 256
    wav = torch.randn(8, S, generator=g) * 0.05
    lens = torch.full((8,), S)
    f1, fr = fe.extract_batch(wav, lens)
    assert fr.tolist() == [281] * 8
    sil, _ = fe.extract_batch(torch.zeros(2, S), torch.full((2,), S), scale=0.1)
    assert torch.equal(sil, torch.full_like(sil, 0.1 * math.log(1e-7)))
    f2, _ = fe.extract_batch(wav * 4.0, lens)
    assert float((f2 - f1 - math.log(4.0)).abs().max()) < 1e-4
    perm = torch.tensor([3, 0, 7, 1, 6, 2, 5, 4])
    fp, _ = fe.extract_batch(wav[perm], lens)
    assert torch.equal(fp, f1[perm.to(f1.device)])


def test_vocoder_properties_fullsize():
    """64 x 938 frames (the bench shape): finite, deterministic, and every utterance independent of its neighbours
    (a permuted batch gives the permuted waveforms bit for bit); frames past an utterance's length change nothing."""
    voc = Vocos(frame_bucket=1).load_state_dict(synth_vocos_state_dict(4)).to("cuda")
    g = torch.Generator().manual_seed(22)
    B, T = 64, 938
    mel = (torch.randn(B, T, 100, generator=g) * 0.2 - 0.4).cuda()
    lens = torch.randint(600, T + 1, (B,), generator=g).cuda()
    w1, l1 = voc.decode_batch(mel, lens, scale=10.0, clamp=True)
    w2, _ = voc.decode_batch(mel, lens, scale=10.0, clamp=True)
    assert torch.isfinite(w1).all() and torch.equal(w1, w2)
    perm = torch.randperm(B, generator=g).cuda()
    wp, lp = voc.decode_batch(mel[perm], lens[perm], scale=10.0, clamp=True)
    assert torch.equal(lp, l1[perm]) and torch.equal(wp, w1[perm])
    junk = mel.clone()
    for i in range(B):
        junk[i, int(lens[i]):] = 7.0                          # garbage past the end of every utterance
    wj, _ = voc.decode_batch(junk, lens, scale=10.0, clamp=True)
    assert torch.equal(wj, w1)
    for i in (0, 17, 63):
        assert float(w1[i, int(l1[i]):].abs().max() if int(l1[i]) < w1.shape[1] else 0.0) == 0.0
