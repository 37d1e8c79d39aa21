"""CPU, world_size 2 over gloo: the multi-GPU path's host logic -- deterministic utterance
partition, per-rank work on its own shard, one all_gather of mels+lengths, original order restored."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from zipvoice_b200.sharding import choose_partition, gather_mels, partition_sorted, partition_utterances


def _fake_mel(uid: int, length: int, F: int) -> torch.Tensor:
    return (torch.arange(length * F, dtype=torch.float32).reshape(length, F) % 17) + 100.0 * uid


def _worker(rank: int, world: int, port: int, lens, out_q, policy="lpt"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    F, U, maxf = 8, len(lens), max(lens)
    shard = (partition_utterances(lens, world) if policy == "lpt" else partition_sorted(lens, world, max_rows=4, frame_bucket=8))[rank]
    Tr = max(lens[i] for i in shard)
    mel = torch.zeros(len(shard), Tr, F)
    for k, uid in enumerate(shard):
        mel[k, : lens[uid]] = _fake_mel(uid, lens[uid], F)
    out, out_lens = gather_mels(mel, torch.tensor([lens[i] for i in shard]), shard, U, maxf)
    ok = out_lens.tolist() == list(lens)
    for uid in range(U):
        ok &= bool(torch.equal(out[uid, : lens[uid]], _fake_mel(uid, lens[uid], F)))
        ok &= float(out[uid, lens[uid]:].abs().sum()) == 0.0
    out_q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


def _run_world2(lens, policy="lpt"):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lens, q, policy)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == {0: True, 1: True}


def test_partition_and_gather_world2():
    _run_world2([37, 12, 50, 44, 9, 28, 31])


def test_skewed_lengths_world2():
    """One long utterance against ten short ones: the frame-balanced shards hold 1 and 10 utterances."""
    lens = [1000] + [100] * 10
    shards = partition_utterances(lens, 2)
    assert sorted(len(s) for s in shards) == [1, 10]
    _run_world2(lens)


def test_sorted_partition_world2():
    """The padding-aware partition (contiguous runs of the length-sorted list) through the same gather."""
    _run_world2([37, 12, 50, 44, 9, 28, 31, 48, 15, 40, 22], policy="sorted")


def test_partition_sorted_cuts_padding():
    import random
    from zipvoice_b200.batcher import batches_cost, plan_batches
    r = random.Random(11)
    lens = [281 + r.randint(600, 938) for _ in range(512)]

    def worst(shards, mr):
        return max(batches_cost([lens[i] for i in s], plan_batches([lens[i] for i in s], mr, 64), 64) for s in shards)

    for world in (2, 4, 8):
        shards = partition_sorted(lens, world, max_rows=64, frame_bucket=64)
        assert sorted(i for s in shards for i in s) == list(range(512))
        assert shards == partition_sorted(list(lens), world, max_rows=64, frame_bucket=64)       # deterministic
        tops = [max(lens[i] for i in s) for s in shards]
        lows = [min(lens[i] for i in s) for s in shards]
        assert all(lows[k] >= tops[k + 1] for k in range(world - 1))                             # contiguous length runs
        assert [len(s) for s in shards] == sorted(len(s) for s in shards)                        # shorter utterances, more rows
        assert worst(shards, 64) < worst(partition_utterances(lens, world), 64)
    assert worst(partition_sorted(lens, 8, 64, 64), 64) < 0.87 * worst(partition_utterances(lens, 8), 64)
    shards, mr = choose_partition(lens, 8)
    assert mr in (64, 96, 128) and worst(shards, mr) <= worst(partition_sorted(lens, 8, 64, 64), 64)
    assert partition_sorted([1000] + [100] * 10, 2) == [[0], list(range(1, 11))]
    assert partition_sorted([5, 3], 4) == [[0], [1], [], []] and partition_sorted([], 2) == [[], []]


def test_gather_single_process():
    lens = [5, 3]
    mel = torch.zeros(2, 5, 4)
    mel[0, :5] = 1.0
    mel[1, :3] = 2.0
    out, ol = gather_mels(mel, torch.tensor(lens), [0, 1], 2, 5)
    assert ol.tolist() == lens and torch.equal(out, mel)
