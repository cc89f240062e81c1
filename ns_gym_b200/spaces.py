"""Minimal action / observation space descriptors for the batched envs.

The reference's callers use ``env.action_space.sample()`` and ``env.action_space.n``
(``tests/test_step_reset.py``, ``benchmark_algorithms/MCTS.py:186-200``); gymnasium is not a
dependency of this package, so the two space kinds the supported envs need are described here.
``sample()`` draws one action PER ENV of the batch, on the batch's device."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Optional

import torch


@dataclass
class Discrete:
    n: int
    num_envs: int = 1
    device: Any = "cpu"

    def sample(self, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        return torch.randint(0, self.n, (self.num_envs,), generator=generator, device=self.device, dtype=torch.int32)

    def contains(self, x) -> bool:
        x = torch.as_tensor(x)
        return bool(((x >= 0) & (x < self.n)).all())


@dataclass
class Box:
    low: float
    high: float
    shape: tuple = (1,)
    num_envs: int = 1
    device: Any = "cpu"
    dtype: torch.dtype = torch.float32

    def sample(self, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        u = torch.rand(self.num_envs, generator=generator, device=self.device, dtype=self.dtype)
        return self.low + (self.high - self.low) * u

    def contains(self, x) -> bool:
        x = torch.as_tensor(x)
        return bool(((x >= self.low) & (x <= self.high)).all())
