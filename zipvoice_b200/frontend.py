"""The step right before the sampler: prompt waveform -> log-mel features on the GPU, plus the
speaker cache the reference's serving path keeps.

Host-side mirror of
  * `VocosFbank` (reference: zipvoice/utils/feature.py:27-120): `extract(samples, sampling_rate)` -> (T, 100) or
    (T, 200) log-mel = torchaudio MelSpectrogram(24 kHz, n_fft 1024, hop 256, 100 mels, center=True, power=1),
    `.clamp(min=1e-7).log()`, trimmed to lhotse's `compute_num_frames` (lhotse 1.32.1, uv.lock:428-430:
    `int((num_samples + hop // 2) // hop)`);
  * `rms_norm` (reference: zipvoice/utils/infer.py:262-281);
  * the speaker cache (reference: runtime/nvidia_triton/pytriton_server.py:86-107 `speaker_info_dict`).
The arithmetic (reflect framing, Hann window, 1024-point FFT, |.|, mel filterbank, log) runs in ONE sm_100a kernel
behind the C ABI (`zvb_fbank`, csrc/audio.cuh); this file only builds the constant tables and does the bookkeeping.
There is no CPU path: waveforms are moved to the model's CUDA device."""
from __future__ import annotations

import collections
import math
from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch

from . import _lib

SAMPLING_RATE = 24000
N_FFT = 1024
HOP_LENGTH = 256
N_MELS = 100


def mel_filterbank(n_freqs: int = N_FFT // 2 + 1, n_mels: int = N_MELS, sample_rate: int = SAMPLING_RATE,
                   f_min: float = 0.0, f_max: Optional[float] = None) -> torch.Tensor:
    """(n_freqs, n_mels) triangular filters on the HTK mel scale, no area normalisation: the table torchaudio's
    MelScale builds with its defaults (mel_scale="htk", norm=None), in the same fp32 operation order."""
    f_max = float(sample_rate // 2) if f_max is None else f_max
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))


def num_frames_for(num_samples: int, hop: int = HOP_LENGTH) -> int:
    """lhotse.utils.compute_num_frames(duration, frame_shift, sampling_rate) in integer arithmetic."""
    return int((int(num_samples) + hop // 2) // hop)


def rms_norm(prompt_wav: torch.Tensor, target_rms: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """Scale the prompt up to `target_rms` when it is quieter (reference: zipvoice/utils/infer.py:262-281)."""
    prompt_rms = torch.sqrt(torch.mean(torch.square(prompt_wav)))
    if prompt_rms < target_rms:
        prompt_wav = prompt_wav * target_rms / prompt_rms
    return prompt_wav, prompt_rms


class VocosFbank:
    """GPU log-mel extractor with the reference extractor's surface (`extract`, `feature_dim`, `frame_shift`)."""

    name = "VocosFbank"

    def __init__(self, num_channels: int = 1, device: Union[str, torch.device] = "cuda"):
        assert num_channels in (1, 2)
        self.num_channels = num_channels
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.ZvbError("zipvoice_b200.frontend runs on a B200 only (no CPU path)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.lib = _lib.load()
        fb = mel_filterbank().t().contiguous()                      # (n_mels, n_freqs)
        nz = fb > 0
        lo = torch.where(nz.any(1), nz.float().argmax(1), torch.zeros(N_MELS, dtype=torch.long))
        hi = torch.where(nz.any(1), fb.shape[1] - nz.flip(1).float().argmax(1), torch.zeros(N_MELS, dtype=torch.long))
        self.fb = fb.to(self.device)
        self.fb_range = torch.stack([lo, hi], 1).to(torch.int32).contiguous().to(self.device)
        self.window = torch.hann_window(N_FFT, periodic=True).to(self.device)

    @property
    def frame_shift(self) -> float:
        return HOP_LENGTH / SAMPLING_RATE

    def feature_dim(self, sampling_rate: int = SAMPLING_RATE) -> int:
        return N_MELS

    def extract_batch(self, wavs: torch.Tensor, lens: torch.Tensor, scale: float = 1.0) -> Tuple[torch.Tensor, torch.Tensor]:
        """wavs (B, S) fp32 zero padded, lens (B,) samples -> (features (B, T, 100) with zeros past each utterance's
        frames, frame counts (B,)); `scale` = the caller's feat_scale (infer_zipvoice.py:372)."""
        assert wavs.dim() == 2
        wavs = wavs.to(self.device, torch.float32).contiguous()
        lens_h = [int(x) for x in lens.tolist()]
        assert all(0 < n <= wavs.shape[1] for n in lens_h), "wav_lens out of range"
        frames = torch.tensor([num_frames_for(n) for n in lens_h], dtype=torch.int64)
        T = max(1, int(frames.max()))
        lens_d = torch.tensor(lens_h, dtype=torch.int32, device=self.device)
        out = torch.empty(wavs.shape[0], T, N_MELS, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.zvb_fbank(wavs.data_ptr(), lens_d.data_ptr(), wavs.shape[0], wavs.shape[1],
                                          self.window.data_ptr(), self.fb.data_ptr(), self.fb_range.data_ptr(), N_MELS,
                                          HOP_LENGTH, float(scale), out.data_ptr(), T,
                                          torch.cuda.current_stream().cuda_stream))
        return out, frames.to(self.device)

    def extract(self, samples, sampling_rate: int):
        """One utterance, the reference's signature: samples (S,), (1, S) or (2, S) -> (T, n_mels * num_channels)."""
        assert sampling_rate == SAMPLING_RATE, f"Mismatched sampling rate: extractor expects {SAMPLING_RATE}, got {sampling_rate}"
        is_numpy = not isinstance(samples, torch.Tensor)
        if is_numpy:
            samples = torch.from_numpy(samples)
        if samples.dim() == 1:
            samples = samples.unsqueeze(0)
        assert samples.dim() == 2, samples.shape
        if self.num_channels == 1:
            if samples.shape[0] == 2:
                samples = samples.mean(dim=0, keepdim=True)
        else:
            assert samples.shape[0] == 2, samples.shape
        C, S = samples.shape
        feats, _ = self.extract_batch(samples, torch.full((C,), S))
        mel = feats.permute(1, 0, 2).reshape(feats.shape[1], C * N_MELS)       # (T, [ch0 | ch1])
        return mel.cpu().numpy() if is_numpy else mel


class SpeakerCache:
    """speaker id -> (prompt_tokens, prompt_features on the device (already * feat_scale), prompt_rms), least recently
    used entries dropped beyond `max_speakers` (reference: pytriton_server.py:86-107 keeps one unbounded dict)."""

    def __init__(self, extractor: VocosFbank, target_rms: float = 0.1, feat_scale: float = 0.1, max_speakers: int = 1024):
        self.extractor = extractor
        self.target_rms, self.feat_scale = target_rms, feat_scale
        self.max_speakers = max_speakers
        self._d: "collections.OrderedDict[str, tuple]" = collections.OrderedDict()
        self.hits = self.misses = 0

    def __len__(self) -> int:
        return len(self._d)

    def __contains__(self, key: str) -> bool:
        return key in self._d

    def put(self, key: str, prompt_wav: torch.Tensor, prompt_tokens: Sequence[int]):
        """prompt_wav (S,) or (C, S) at 24 kHz (stereo is averaged, as the reference does before caching)."""
        if prompt_wav.dim() == 2:
            prompt_wav = prompt_wav.mean(dim=0) if prompt_wav.shape[0] > 1 else prompt_wav[0]
        wav, rms = rms_norm(prompt_wav.float(), self.target_rms)
        feats, frames = self.extractor.extract_batch(wav.unsqueeze(0), torch.tensor([wav.numel()]), self.feat_scale)
        entry = (list(prompt_tokens), feats[0, : int(frames[0])].contiguous(), float(rms))
        self._d[key] = entry
        self._d.move_to_end(key)
        while len(self._d) > self.max_speakers:
            self._d.popitem(last=False)
        return entry

    def get(self, key: str, prompt_wav: Optional[torch.Tensor] = None, prompt_tokens: Optional[Sequence[int]] = None):
        e = self._d.get(key)
        if e is not None:
            self.hits += 1
            self._d.move_to_end(key)
            return e
        self.misses += 1
        if prompt_wav is None or prompt_tokens is None:
            raise KeyError(f"speaker {key!r} is not cached and no prompt was given")
        return self.put(key, prompt_wav, prompt_tokens)

    def batch(self, keys: Sequence[str]):
        """Padded (prompt_tokens list, prompt_features (B, Tmax, 100), prompt_features_lens (B,), rms list) of cached
        speakers: the arguments `model.sample` takes."""
        es = [self.get(k) for k in keys]
        lens = torch.tensor([e[1].shape[0] for e in es], dtype=torch.int64, device=self.extractor.device)
        feats = torch.nn.utils.rnn.pad_sequence([e[1] for e in es], batch_first=True)
        return [e[0] for e in es], feats, lens, [e[2] for e in es]
