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


def _synthetic_totals(off, cnt):
    """The totals vector (native.STAT_KEYS) a shard's device accumulator (nsgym_episode_stats) would
    hold after 12 steps of deterministic synthetic outputs keyed by GLOBAL env id: reward 1 per
    step, env g terminates every (g + 2) steps."""
    from ns_gym_b200.native import STAT_KEYS

    tot = torch.zeros(len(STAT_KEYS), dtype=torch.float64)
    run = torch.zeros(cnt, dtype=torch.float64)
    gids = torch.arange(off, off + cnt)
    for k in range(1, 13):
        ended = (k % (gids + 2)) == 0
        run += 1.0
        tot[0] += cnt
        tot[1] += ended.sum()
        tot[2] += run[ended].sum()
        tot[3] += run[ended].sum()
        tot[4] += ended.sum()
        run[ended] = 0
    return tot


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, _local, w = D.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    off, cnt = D.shard_range(10, rank, world)
    red = D.reduce_totals(_synthetic_totals(off, cnt))
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
    want = D.reduce_totals(_synthetic_totals(0, 10))
    assert red == want
    assert red["steps"] == 120 and red["episodes"] == sum(12 // (g + 2) for g in range(10))
    assert slow == 2.0
