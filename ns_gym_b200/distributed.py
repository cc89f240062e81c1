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
from dataclasses import dataclass

import torch
import torch.distributed as dist

METRIC_KEYS = ("steps", "episodes", "return_sum", "length_sum", "terminated", "truncated")


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


@dataclass
class EpisodeStats:
    """On-device episode accumulators, reduced over ranks on demand.

    ``update`` consumes the step kernel's outputs (reward float[N], flags uint8[N]); it keeps
    running per-env return / length and folds finished episodes into six scalars."""
    device: torch.device
    n_envs: int

    def __post_init__(self):
        self.running_return = torch.zeros(self.n_envs, dtype=torch.float64, device=self.device)
        self.running_length = torch.zeros(self.n_envs, dtype=torch.int64, device=self.device)
        self.totals = torch.zeros(len(METRIC_KEYS), dtype=torch.float64, device=self.device)

    def update(self, reward: torch.Tensor, flags: torch.Tensor):
        stepped = (flags & 4) == 0                      # NSGYM_FLAG_RESET calls are not env steps
        term = (flags & 1) != 0
        trunc = (flags & 2) != 0
        ended = term | trunc
        self.running_return += torch.where(stepped, reward.double(), torch.zeros_like(self.running_return))
        self.running_length += stepped.long()
        self.totals[0] += stepped.sum()
        self.totals[1] += ended.sum()
        self.totals[2] += self.running_return[ended].sum()
        self.totals[3] += self.running_length[ended].sum()
        self.totals[4] += term.sum()
        self.totals[5] += trunc.sum()
        self.running_return[ended] = 0
        self.running_length[ended] = 0

    def reduce(self) -> dict:
        """Whole-job totals (sum over ranks) plus derived means."""
        tot = all_reduce_sum(self.totals.clone())
        out = {k: float(v) for k, v in zip(METRIC_KEYS, tot.tolist())}
        ep = max(out["episodes"], 1.0)
        out["mean_return"] = out["return_sum"] / ep
        out["mean_length"] = out["length_sum"] / ep
        return out
