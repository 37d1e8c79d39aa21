"""Utterance sharding across the GPUs of one box: the sampler has no
inter-utterance dependency (the only cross-row op, the CFG pairing of row b with row b+B, stays on
one GPU), so a batch is partitioned by utterance with no collective on the hot path; the only
exchange is the final gather of the padded output mels and their lengths.  The partition follows
`batchify_tokens`' sort-by-length habit (reference: zipvoice/utils/infer.py:131-139): longest first,
each utterance to the currently lightest rank, which balances frames (the work per utterance is
~linear in frames, quadratic only in the attention share)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def partition_utterances(total_frames: Sequence[int], world_size: int) -> List[List[int]]:
    """Returns, per rank, the utterance indices it samples (deterministic on every rank)."""
    order = sorted(range(len(total_frames)), key=lambda i: (-int(total_frames[i]), i))
    loads = [0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += int(total_frames[i])
    return shards


def gather_mels(mel: torch.Tensor, lens: torch.Tensor, shard: Sequence[int], num_utts: int,
                max_frames: int, per_rank: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """All ranks contribute their (b, T_r, F) zero-padded mels; every rank gets (num_utts, max_frames,
    F) in the original utterance order plus the lengths.  One all_gather of a fixed-size buffer
    (NCCL over NVLink on the GPU box; gloo in the CPU tests).

    `per_rank` = utterances in the LARGEST shard (every rank passes the same value: `max(len(s) for s in
    partition_utterances(...))`; the partition balances frames, not counts, so shards can differ by many
    utterances).  0: the ranks agree on it with one extra all_reduce."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    F = mel.shape[2]
    per = int(per_rank)
    if per <= 0:
        cnt = torch.tensor([len(shard)], dtype=torch.int64, device=mel.device)
        if world > 1:
            dist.all_reduce(cnt, op=dist.ReduceOp.MAX)
        per = int(cnt.item())
    if len(shard) > per:
        raise ValueError(f"gather_mels: this rank holds {len(shard)} utterances but per_rank is {per}")
    buf = torch.zeros(per, max_frames, F, dtype=mel.dtype, device=mel.device)
    meta = torch.full((per, 2), -1, dtype=torch.int64, device=mel.device)   # (utterance id, length)
    n = len(shard)
    buf[:n, : mel.shape[1]] = mel[:, :max_frames]
    meta[:n, 0] = torch.as_tensor(list(shard), dtype=torch.int64, device=mel.device)
    meta[:n, 1] = lens.to(torch.int64)
    if world > 1:
        bufs = [torch.empty_like(buf) for _ in range(world)]
        metas = [torch.empty_like(meta) for _ in range(world)]
        dist.all_gather(bufs, buf)
        dist.all_gather(metas, meta)
    else:
        bufs, metas = [buf], [meta]
    out = torch.zeros(num_utts, max_frames, F, dtype=mel.dtype, device=mel.device)
    out_lens = torch.zeros(num_utts, dtype=torch.int64, device=mel.device)
    for b, m in zip(bufs, metas):
        valid = m[:, 0] >= 0
        ids = m[valid, 0]
        out[ids] = b[valid]
        out_lens[ids] = m[valid, 1]
    return out, out_lens
