"""CPU, world_size 2 over gloo: the N > 1 host logic (shard ranges, metric reduction, max-over-
ranks timing).  The data path itself has no collective."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ns_gym_b200 import distributed as D


def test_shard_ranges_partition_the_batch():
    for n in (1, 7, 16, 1000003):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0
            assert sum(c for _, c in spans) == n
            for (o0, c0), (o1, _c1) in zip(spans, spans[1:]):
                assert o0 + c0 == o1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        D.shard_range(8, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, _local, w = D.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    off, cnt = D.shard_range(10, rank, world)
    stats = D.EpisodeStats(torch.device("cpu"), cnt)
    # deterministic synthetic step outputs keyed by GLOBAL env id: env g ends every (g + 2) steps
    gids = torch.arange(off, off + cnt)
    for k in range(1, 13):
        reward = torch.ones(cnt)
        flags = ((k % (gids + 2)) == 0).to(torch.uint8)            # terminated bit
        stats.update(reward, flags)
    red = stats.reduce()
    slow = D.max_over_ranks(1.0 + rank)
    if rank == 0:
        out.put((red, slow))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_metric_reduction_matches_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    red, slow = q.get()
    # single-process truth over all 10 envs
    stats = D.EpisodeStats(torch.device("cpu"), 10)
    gids = torch.arange(10)
    for k in range(1, 13):
        stats.update(torch.ones(10), ((k % (gids + 2)) == 0).to(torch.uint8))
    want = stats.reduce()
    assert red == want
    assert red["steps"] == 120 and red["episodes"] == sum(12 // (g + 2) for g in range(10))
    assert slow == 2.0
