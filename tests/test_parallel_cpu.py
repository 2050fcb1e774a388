"""World-size-2 gloo test of the multi-GPU host logic (sharding + final transcript gather) on CPU."""
import os

import pytest
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from asr_transformer_b200.parallel import balanced_assignment, gather_tokens, shard_range


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 64, 65, 257):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_balanced_assignment():
    lens = [400, 1000, 800, 650, 990, 410, 700, 520]
    parts = balanced_assignment(lens, 2)
    assert sorted(sum(parts, [])) == list(range(8))
    loads = [sum(lens[i] for i in p) for p in parts]
    assert abs(loads[0] - loads[1]) < 0.1 * sum(lens)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, L = 5, 6
        lo, hi = shard_range(B, rank, world)
        counts = [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]
        full = torch.arange(B * (L + 1), dtype=torch.int32).view(B, L + 1)
        n_full = torch.arange(B, dtype=torch.int32) + 2
        tok, n = gather_tokens(full[lo:hi].clone(), n_full[lo:hi].clone(), counts)
        q.put((rank, torch.equal(tok, full) and torch.equal(n, n_full)))
    finally:
        dist.destroy_process_group()


def test_gather_tokens_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(results) == [(0, True), (1, True)]


def test_bucket_by_length():
    from asr_transformer_b200.parallel import bucket_by_length
    lens = [500, 1000, 400, 990, 410, 700, 1000]
    b = bucket_by_length(lens, 3)
    assert sorted(i for g in b for i in g) == list(range(len(lens)))
    assert b == [[1, 6, 3], [5, 0, 4], [2]]
    assert max(lens[i] for i in b[1]) - min(lens[i] for i in b[1]) < max(lens) - min(lens)
    assert bucket_by_length([], 4) == []
    with pytest.raises(ValueError):
        bucket_by_length(lens, 0)
