"""Long-form driver around the sampler and the vocoder: split a long token sequence into chunks,
sample the chunks in length-sorted batches against one cached prompt, decode them, and join the waveforms with a
cross-fade -- the flow of the reference's `generate_sentence` (reference: zipvoice/bin/infer_zipvoice.py:424-640) on
top of its helpers `chunk_tokens_punctuation` / `chunk_tokens_dialog` / `batchify_tokens` / `cross_fade_concat`
(reference: zipvoice/utils/infer.py:12-229), restated here with the same signatures and results.  Tokenisers, silence
removal (pydub) and file I/O stay outside (out of scope); inputs are token lists."""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from .batcher import batchify_tokens

punctuation = {";", ":", ",", ".", "!", "?", "；", "：", "，", "。", "！", "？"}


def _pack(units: List[list], max_tokens: int) -> List[list]:
    """Greedy packing of whole units into chunks of at most `max_tokens` tokens (a longer unit stands alone)."""
    chunks, cur = [], []
    for u in units:
        if len(cur) + len(u) <= max_tokens:
            cur.extend(u)
        else:
            if cur:
                chunks.append(cur)
            cur = u
    if cur:
        chunks.append(cur)
    return chunks


def chunk_tokens_punctuation(tokens_list: List[str], max_tokens: int = 100) -> List[List[str]]:
    """Sentences end at a punctuation token; punctuation / blanks that open a sentence stick to the previous one
    (reference: utils/infer.py:15-58)."""
    sentences, cur = [], []
    for tok in tokens_list:
        if not cur and sentences and (tok in punctuation or tok == " "):
            sentences[-1].append(tok)
            continue
        cur.append(tok)
        if tok in punctuation:
            sentences.append(cur)
            cur = []
    if cur:
        sentences.append(cur)
    return _pack(sentences, max_tokens)


def chunk_tokens_dialog(tokens_list: List[str], max_tokens: int = 100) -> List[List[str]]:
    """A unit starts at every speaker-A turn symbol `[S1]` (reference: utils/infer.py:61-105)."""
    dialogs, cur = [], []
    for tok in tokens_list:
        if tok == "[S1]":
            if cur:
                dialogs.append(cur)
            cur = []
        cur.append(tok)
    if cur:
        dialogs.append(cur)
    return _pack(dialogs, max_tokens)


def cross_fade_concat(chunks: List[torch.Tensor], fade_duration: float = 0.1, sample_rate: int = 24000) -> torch.Tensor:
    """Join (C, T_i) waveforms, blending the last / first k = min(fade, len(previous result), len(next)) samples with
    a linear ramp (reference: utils/infer.py:173-229).  The result is assembled in ONE output buffer: every chunk is
    written once, scaled by its fade-in / fade-out ramps (the reference re-concatenates the growing result per chunk)."""
    if len(chunks) <= 1:
        return chunks[0] if chunks else torch.tensor([])
    fade = int(fade_duration * sample_rate)
    if fade <= 0:
        return torch.cat(chunks, dim=-1)
    # k_i: overlap between the running result and chunk i (the running length only depends on lengths)
    ks, run = [], chunks[0].shape[-1]
    for c in chunks[1:]:
        k = max(0, min(fade, run, c.shape[-1]))
        ks.append(k)
        run = run + c.shape[-1] - k
    out = torch.zeros(*chunks[0].shape[:-1], run, dtype=chunks[0].dtype, device=chunks[0].device)
    pos = 0
    for i, c in enumerate(chunks):
        k_in = ks[i - 1] if i > 0 else 0
        seg = c
        if k_in > 0:          # fade-in of this chunk over the previous result's tail: weight (1 - linspace(1, 0, k))
            w = 1 - torch.linspace(1, 0, k_in, device=c.device)
            seg = torch.cat([c[..., :k_in] * w, c[..., k_in:]], dim=-1)
        out[..., pos: pos + c.shape[-1]] += seg
        pos += c.shape[-1]
        if i < len(ks) and ks[i] > 0:      # fade-out of the RESULT's last k samples (may reach into earlier chunks)
            k = ks[i]
            out[..., pos - k: pos] *= torch.linspace(1, 0, k, device=c.device)
            pos -= k
    return out


@torch.inference_mode()
def generate_long(model, vocoder, chunked_tokens: List[List[int]], prompt_tokens: List[int], prompt_features: torch.Tensor,
                  prompt_duration: float, token_duration: float, prompt_rms: float = 1.0, target_rms: float = 0.1,
                  max_duration: float = 100.0, feat_scale: float = 0.1, fade_duration: float = 0.1, sampling_rate: int = 24000,
                  **sample_kwargs) -> torch.Tensor:
    """Chunks -> one waveform (1, T): the body of the reference's `generate_sentence` after tokenisation
    (infer_zipvoice.py:545-628).  `prompt_features`: (Tp, 100) feat-scaled (e.g. a SpeakerCache entry)."""
    batches, index = batchify_tokens(chunked_tokens, max_duration, prompt_duration, token_duration)
    wavs: List[torch.Tensor] = []
    for batch in batches:
        B = len(batch)
        pf = prompt_features.unsqueeze(0).expand(B, -1, -1).contiguous()
        pfl = torch.full((B,), prompt_features.shape[0], dtype=torch.int64, device=prompt_features.device)
        mel, lens, _, _ = model.sample(tokens=batch, prompt_tokens=[list(prompt_tokens)] * B, prompt_features=pf,
                                       prompt_features_lens=pfl, duration="predict", **sample_kwargs)
        audio, alens = vocoder.decode_batch(mel, lens, scale=1.0 / feat_scale, clamp=True)
        for i in range(B):
            w = audio[i: i + 1, : int(alens[i])]
            if prompt_rms < target_rms:
                w = w * prompt_rms / target_rms
            wavs.append(w)
    ordered = [w for _, w in sorted(zip(index, wavs), key=lambda p: p[0])]
    return cross_fade_concat(ordered, fade_duration=fade_duration, sample_rate=sampling_rate)
