#!/usr/bin/env python3
"""Golden vectors for the prompt feature extractor, generated with the reference's own dependency:
the exact torchaudio calls of zipvoice/utils/feature.py:47-59 (`MelSpectrogram(...)(samples).clamp(min=1e-7).log()`)
and its trimming to lhotse's frame count (:100-112).  Run in the build container:  python tools/make_golden_audio.py
Writes tests/golden/fbank_{mono,stereo,short}.pt (waveform + log-mel)."""
import math
import os
import sys

import torch
import torchaudio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")


def synth_wave(seed: int, samples: int, channels: int = 1) -> torch.Tensor:
    """Speech-like test signal: harmonic stacks with a moving pitch, an amplitude envelope spanning 50 dB, a silent
    stretch and a little noise (quiet high-frequency bins exercise the log)."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(samples) / 24000.0
    out = []
    for c in range(channels):
        f0 = 110.0 + 40.0 * torch.sin(2 * math.pi * (0.7 + 0.2 * c) * t)
        phase = 2 * math.pi * torch.cumsum(f0, 0) / 24000.0
        x = sum(torch.sin(k * phase) / k ** 1.5 for k in range(1, 30))
        env = (0.003 + torch.sin(2 * math.pi * 1.3 * t + c).clamp_min(0) ** 2)
        x = x * env * 0.2 + 1e-3 * torch.randn(samples, generator=g)
        x[int(0.4 * samples): int(0.45 * samples)] = 0.0
        out.append(x)
    return torch.stack(out).float()


def reference_logmel(samples: torch.Tensor) -> torch.Tensor:
    """(C, S) -> (T, C * 100), the lines of VocosFbank.extract (feature.py:47-59, 93-112)."""
    fbank = torchaudio.transforms.MelSpectrogram(sample_rate=24000, n_fft=1024, hop_length=256, n_mels=100, center=True,
                                                 power=1)
    mel = fbank(samples)
    logmel = mel.clamp(min=1e-7).log()
    mel = logmel.reshape(-1, logmel.shape[-1]).t()
    num_frames = int((samples.shape[1] + 256 // 2) // 256)         # lhotse.utils.compute_num_frames (1.32.1)
    if mel.shape[0] > num_frames:
        mel = mel[:num_frames]
    elif mel.shape[0] < num_frames:
        mel = torch.nn.functional.pad(mel.unsqueeze(0), (0, 0, 0, num_frames - mel.shape[0]), mode="replicate").squeeze(0)
    return mel


CASES = {"fbank_mono": (11, 55333, 1), "fbank_stereo": (12, 30001, 2), "fbank_short": (13, 1400, 1)}

if __name__ == "__main__":
    for name, (seed, n, ch) in CASES.items():
        wav = synth_wave(seed, n, ch)
        mel = reference_logmel(wav)
        torch.save({"wav": wav, "logmel": mel, "torchaudio": torchaudio.__version__}, os.path.join(OUT, name + ".pt"))
        print(name, tuple(wav.shape), tuple(mel.shape), float(mel.min()), float(mel.max()))
