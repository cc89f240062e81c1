"""Multi-GPU plumbing: one process per GPU, envs sharded by contiguous global index.

The step path has no exchange step: each wrapper instance of the reference owns
its env, theta and update-function state (``ns_gym/base.py:263-265``), and the reference itself
only parallelises whole episodes (``ns_gym/evaluate/run_experiment.py:220-239``).  So a shard is
just an ``NSVectorEnv`` created with ``env_id_offset = first global id``: Philox counters use
global ids, hence any (rank, world) layout reproduces the single-GPU results bit for bit.

The only collectives are a max over ranks of a timing and a sum of a <= 16-element fp64 metric
vector (NCCL over NVLink on GPUs, gloo in the CPU tests) -- latency-bound, off the step stream.
"""
from __future__ import annotations

import os
import torch
import torch.distributed as dist

def shard_range(n_global: int, rank: int, world: int):
    """(first global env id, count) of rank's contiguous slice; remainders go to the low ranks."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(int(n_global), int(world))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def init_from_env(backend: str | None = None):
    """Join the process group torchrun set up (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, local, world


def all_reduce_sum(vec: torch.Tensor) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM)
    return vec


def max_over_ranks(value: float, device=None) -> float:
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_totals(totals: torch.Tensor) -> dict:
    """Whole-job totals (sum over ranks of a ``native.STAT_KEYS`` vector) plus derived means."""
    from .native import STAT_KEYS

    tot = all_reduce_sum(totals.clone())
    out = {k: float(v) for k, v in zip(STAT_KEYS, tot.tolist())}
    ep = max(out["episodes"], 1.0)
    out["mean_return"] = out["return_sum"] / ep
    out["mean_length"] = out["length_sum"] / ep
    return out


class EpisodeStats:
    """On-device episode accumulators of one ``NSVectorEnv`` shard, reduced over ranks on demand.

    ``update()`` is ONE kernel launch (``nsgym_episode_stats``) over the step's reward / flag
    outputs: per-env running return / length, finished episodes folded into eight fp64 totals
    (``native.STAT_KEYS``; they include the constraint-rejection and bad-distribution counts, the
    batch counterpart of the reference's ``ConstraintViolationWarning`` / ``ValueError``).  No host
    synchronisation until ``reduce()`` reads the totals."""

    def __init__(self, env):
        from . import native as nv

        self.env = env
        self._nv = nv
        dev, n = env.device, env.num_envs
        self.running_return = torch.zeros(n, dtype=torch.float64, device=dev)
        self.running_length = torch.zeros(n, dtype=torch.int32, device=dev)
        self.totals = torch.zeros(len(nv.STAT_KEYS), dtype=torch.float64, device=dev)

    def update(self):
        """Fold the outputs of the env's last step into the accumulators."""
        import ctypes as C

        env, nv = self.env, self._nv
        with torch.cuda.device(env.device):
            nv.check(env.lib.nsgym_episode_stats(
                env._h, C.c_void_p(self.running_return.data_ptr()), C.c_void_p(self.running_length.data_ptr()),
                C.c_void_p(self.totals.data_ptr()), env._stream()), "nsgym_episode_stats")

    def reduce(self) -> dict:
        return reduce_totals(self.totals)
