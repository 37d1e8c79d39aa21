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


def partition_sorted(total_frames: Sequence[int], world_size: int, max_rows: int = 64,
                     frame_bucket: int = 64) -> List[List[int]]:
    """Padding-aware partition: the utterances sorted longest first are cut into `world_size` CONTIGUOUS runs, so every
    rank batches utterances of similar length (the reference's sort-by-length, utils/infer.py:131-139, applied across
    the GPUs as well as inside one), and the cuts minimise the largest `batcher.batches_cost` of a rank: a rank of
    short utterances takes more of them.  `partition_utterances` balances VALID frames and leaves every rank the full
    length range, i.e. 18 % padding for U[881,1219] frames at 64 utterances per rank; this one leaves ~5 %.
    Deterministic on every rank; ranks past the number of utterances get empty shards."""
    from .batcher import batch_cost
    from .engine import round_up
    order = sorted(range(len(total_frames)), key=lambda i: (-int(total_frames[i]), i))
    lens = [int(total_frames[i]) for i in order]
    n = len(lens)
    memo = {}

    def cost(a: int, b: int) -> float:
        """`batcher.batches_cost` of `plan_batches(lens[a:b])`: equal row counts, each batch padded to its first utterance."""
        if b <= a:
            return 0.0
        if (a, b) not in memo:
            nb = -(-(b - a) // max_rows)
            base, extra = divmod(b - a, nb)
            c, s = 0.0, a
            for k in range(nb):
                rows = base + (1 if k < extra else 0)
                c += batch_cost(rows, round_up(lens[s], frame_bucket))
                s += rows
            memo[(a, b)] = c
        return memo[(a, b)]

    def cuts_for(limit: float):
        cuts, a = [], 0
        for _ in range(world_size):
            b = n                               # the longest run from `a` within the limit (the cost is not monotone in
            while b > a and cost(a, b) > limit:  # the run length: one more row can split a batch in two shorter ones)
                b -= 1
            cuts.append((a, b))
            a = b
        return cuts if a == n else None

    lo, hi = 0.0, cost(0, n)
    for _ in range(48):                         # bisection on the largest per-rank cost
        mid = 0.5 * (lo + hi)
        if cuts_for(mid) is not None:
            hi = mid
        else:
            lo = mid
    cuts = cuts_for(hi)
    return [[order[i] for i in range(a, b)] for a, b in cuts]


def choose_partition(total_frames: Sequence[int], world_size: int, row_limits: Sequence[int] = (64, 96, 128),
                     frame_bucket: int = 64) -> Tuple[List[List[int]], int]:
    """`partition_sorted` under each candidate batch-row limit; returns the shards and the limit whose slowest rank is
    predicted fastest (`batcher.batches_cost`).  Same answer on every rank."""
    from .batcher import batches_cost, plan_batches
    best = None
    for mr in row_limits:
        shards = partition_sorted(total_frames, world_size, max_rows=mr, frame_bucket=frame_bucket)
        worst = 0.0
        for s in shards:
            seg = [int(total_frames[i]) for i in s]
            worst = max(worst, batches_cost(seg, plan_batches(seg, max_rows=mr, frame_bucket=frame_bucket), frame_bucket))
        if best is None or worst < best[0]:
            best = (worst, shards, mr)
    return best[1], best[2]


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
