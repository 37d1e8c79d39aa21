"""Length-sorted batching in front of the sampler.

The reference sorts sentences by token count and packs them greedily under a duration budget before it calls
`model.sample` once per batch (reference: zipvoice/utils/infer.py:108-170 `batchify_tokens`, used by
`generate_sentence`, zipvoice/bin/infer_zipvoice.py:355-420); its Triton backends leave the same job to the
server's dynamic batcher (runtime/nvidia_triton/model_repo/zipvoice/config.pbtxt:17-20).  Here the same policy
is stated in frames, which is what the CUDA plans are shaped by:

* `batchify_tokens` keeps the reference's signature and result (batches of token lists + the original index of
  every sorted sentence);
* `plan_batches` packs utterances, longest first, into batches of at most `max_rows` utterances and
  `max_batch_frames` padded frames, every batch's frame count rounded up to `frame_bucket` so that the
  (rows, frames) plans and CUDA graphs behind `model.sample` are reused across batches (engine.PlanCache);
* `sample_batched` runs `model.sample` batch by batch and restores the callers' order.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from .engine import round_up


def batchify_tokens(tokens_list: List[List[int]], max_duration: float, prompt_duration: float,
                    token_duration: float) -> Tuple[List[List[List[int]]], List[int]]:
    """Same contract as the reference's `batchify_tokens` (utils/infer.py:108-170): sentences sorted by
    token count (stable), packed while `tokens*token_duration + sentences*prompt_duration` stays within
    `max_duration`; returns the batches and, for every sorted position, the sentence's original index."""
    order = sorted(range(len(tokens_list)), key=lambda i: len(tokens_list[i]))
    batches: List[List[List[int]]] = []
    cur: List[List[int]] = []
    cur_tokens = 0
    for i in order:
        t = tokens_list[i]
        if cur and (cur_tokens + len(t)) * token_duration + len(cur) * prompt_duration > max_duration:
            batches.append(cur)
            cur, cur_tokens = [], 0
        cur.append(t)
        cur_tokens += len(t)
    if cur:
        batches.append(cur)
    return batches, order


def batch_cost(rows: int, frames: int) -> float:
    """Relative time of one sampler batch of `rows` utterances padded to `frames`, fitted to B200 measurements
    (profiles/batch_sweep_r2.txt: 59.6 / 71.5 / 77.3 / 81 / 82.2 k frames/s at 8 / 16 / 32 / 64 / 96 rows of 1219 frames =
    85.4 k x rows / (rows + 3.5)): every batch costs 3.5 rows of fixed work, and the attention share (26 % of a
    1219-frame forward: bench.py's per-kernel table) is quadratic in the frames, i.e. T / 3500 of the linear term."""
    return (rows + 3.5) * frames * (1.0 + frames / 3500.0)


def plan_batches(total_frames: Sequence[int], max_rows: int = 64, frame_bucket: int = 64,
                 max_batch_frames: Optional[int] = None) -> List[List[int]]:
    """Utterance indices per batch.  Longest first, so a batch's padded length is its first utterance's length
    rounded up to `frame_bucket`.  Without a frame budget the sorted list is cut into ceil(n / max_rows) batches of
    (nearly) equal row counts -- 70 utterances become 35 + 35, not 64 + 6: a small tail batch runs the GPU at a fraction
    of its rate.  With `max_batch_frames` a batch closes when it holds `max_rows` utterances or one more row would push
    rows x padded frames over the budget.  Deterministic (ties broken by index)."""
    order = sorted(range(len(total_frames)), key=lambda i: (-int(total_frames[i]), i))
    if max_batch_frames is None:
        n = len(order)
        if n == 0:
            return []
        nb = -(-n // max_rows)
        base, extra = divmod(n, nb)
        batches, a = [], 0
        for k in range(nb):
            b = a + base + (1 if k < extra else 0)
            batches.append(order[a:b])
            a = b
        return batches
    batches: List[List[int]] = []
    cur: List[int] = []
    cur_T = 0
    for i in order:
        if cur and (len(cur) >= max_rows or (len(cur) + 1) * cur_T > max_batch_frames):
            batches.append(cur)
            cur = []
        if not cur:
            cur_T = round_up(int(total_frames[i]), frame_bucket)
        cur.append(i)
    if cur:
        batches.append(cur)
    return batches


def batches_cost(total_frames: Sequence[int], batches: Sequence[Sequence[int]], frame_bucket: int = 0) -> float:
    """Sum of `batch_cost` over the batches (each padded to its longest utterance, rounded up to the bucket)."""
    return sum(batch_cost(len(b), round_up(max(int(total_frames[i]) for i in b), frame_bucket)) for b in batches if b)


def padding_waste(total_frames: Sequence[int], batches: Sequence[Sequence[int]], frame_bucket: int = 0) -> float:
    """Share of the computed frames that is padding: 1 - valid / (rows x padded frames)."""
    valid = sum(int(total_frames[i]) for b in batches for i in b)
    computed = sum(len(b) * round_up(max(int(total_frames[i]) for i in b), frame_bucket) for b in batches)
    return 1.0 - valid / max(1, computed)


@torch.inference_mode()
def sample_batched(model, tokens: List[List[int]], prompt_tokens: List[List[int]], prompt_features: torch.Tensor,
                   prompt_features_lens: torch.Tensor, features_lens: Optional[torch.Tensor] = None, speed: float = 1.0,
                   max_rows: int = 64, max_batch_frames: Optional[int] = None, **sample_kwargs):
    """`model.sample` over length-sorted batches.  Inputs as `ZipVoice.sample` for the whole utterance set
    (`prompt_features` zero padded to the longest prompt); returns the same 4-tuple in the callers' order, the
    generated mels zero padded to the longest one."""
    duration = "real" if features_lens is not None else "predict"
    pfl = prompt_features_lens.cpu()
    if features_lens is not None:
        total = (pfl + features_lens.cpu()).tolist()
    else:      # the ratio-duration rule (reference: zipvoice.py:316-322), host side, only to sort and pack
        pl = torch.tensor([len(t) for t in prompt_tokens], dtype=torch.int64)
        tl = torch.tensor([len(t) for t in tokens], dtype=torch.int64)
        total = (pfl + torch.ceil(pfl / pl * tl / speed).to(torch.int64)).tolist()
    bucket = getattr(model, "frame_bucket", 0)
    batches = plan_batches(total, max_rows=max_rows, frame_bucket=bucket, max_batch_frames=max_batch_frames)
    U = len(tokens)
    parts = []
    for b in batches:
        idx = torch.tensor(b, dtype=torch.int64)
        pf = prompt_features[idx.to(prompt_features.device)]
        pf = pf[:, : int(pfl[idx].max())]
        mel, lens, pmel, plens = model.sample(
            [tokens[i] for i in b], [prompt_tokens[i] for i in b], pf, prompt_features_lens[idx.to(prompt_features_lens.device)],
            features_lens=None if features_lens is None else features_lens[idx.to(features_lens.device)],
            speed=speed, duration=duration, **sample_kwargs)
        parts.append((idx, mel, lens, pmel, plens))
    dev, F = parts[0][1].device, parts[0][1].shape[-1]
    x1 = torch.zeros(U, max(p[1].shape[1] for p in parts), F, device=dev)
    xp = torch.zeros(U, max(p[3].shape[1] for p in parts), F, device=dev)
    x1_lens = torch.zeros(U, dtype=torch.int64, device=dev)
    xp_lens = torch.zeros(U, dtype=torch.int64, device=dev)
    for idx, mel, lens, pmel, plens in parts:      # one indexed copy per batch (outputs are already zero padded)
        idx = idx.to(dev)
        x1[idx, : mel.shape[1]] = mel
        xp[idx, : pmel.shape[1]] = pmel
        x1_lens[idx] = lens.to(dev)
        xp_lens[idx] = plens.to(dev)
    return x1, x1_lens, xp, xp_lens
