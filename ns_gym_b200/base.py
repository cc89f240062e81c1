"""Core plugin types of the batched NS simulator.

Mirrors the reference's plugin surface for the step path (``ns_gym/base.py:33-203``):
``Reward``, ``Scheduler``, ``UpdateFn``, ``UpdateDistributionFn`` and ``TUNABLE_PARAMS``.

In the reference these objects *execute* the update on the host, one env at a time
(``UpdateFn.__call__`` -> ``Scheduler.__call__`` -> ``_update``).  Here they are pure
**descriptions**: ``ns_gym_b200.compile`` lowers a ``tunable_params`` dict to the
opcode-and-coefficient table interpreted by the CUDA step kernel.  They deliberately
have no host-side ``__call__`` -- there is no CPU execution path in this package.
``ns_gym_b200.evaluate_update`` runs the fire-test + advance stages alone on the GPU
for known-answer checks.

The compiler is duck-typed on class *names* and attribute names, so dictionaries built
from the reference's own classes compile as well.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Union

import numpy as np


@dataclass(frozen=True)
class Reward:
    """Batched counterpart of ``ns_gym.base.Reward`` (``base.py:33-47``).

    ``reward`` is a tensor ``[N]``; ``env_change`` / ``delta_change`` map parameter name to
    tensors ``[N]``; ``relative_time`` is an int tensor ``[N]``.
    """

    reward: Any
    env_change: dict
    delta_change: Union[dict, None]
    relative_time: Any


class Scheduler:
    """When a parameter changes.  ``start <= t <= end`` (both inclusive) gates every
    subclass rule (``base.py:56-81``)."""

    def __init__(self, start=0, end=np.inf) -> None:
        self.start = start
        self.end = end

    def __call__(self, t):
        raise RuntimeError(
            "ns_gym_b200 schedulers are descriptions compiled for the CUDA step kernel; "
            "there is no host execution path (use ns_gym_b200.evaluate_update on a GPU)."
        )


class UpdateFn:
    """How a scalar parameter changes when its scheduler fires (``base.py:98-182``)."""

    def __init__(self, scheduler: Scheduler) -> None:
        # duck-typed: accept the reference's Scheduler instances too
        assert hasattr(scheduler, "start") and hasattr(scheduler, "end"), (
            f"Expected scheduler to be a subclass of Scheduler, got {type(scheduler)}"
        )
        self.scheduler = scheduler
        self.prev_param = None
        self.prev_time = -1

    def __call__(self, param, t):
        raise RuntimeError(
            "ns_gym_b200 update functions are descriptions compiled for the CUDA step "
            "kernel; there is no host execution path (use ns_gym_b200.evaluate_update)."
        )


class UpdateDistributionFn(UpdateFn):
    """How a slip distribution (list of 3 or 4 probabilities) changes (``base.py:185-203``).
    The reported change is the 1-Wasserstein distance on indices."""


# Default physical parameters (``base.py:605-684`` reads them from live gymnasium envs;
# ``base.py:1161-1165`` hard-codes Bridge).  Values: gymnasium 1.2.1 defaults, see
# docs/source/env_pages/classic_control/*.md in the reference.
TUNABLE_PARAMS = {
    "CartPoleEnv": {
        "gravity": 9.8, "masscart": 1.0, "masspole": 0.1, "force_mag": 10.0, "tau": 0.02,
        "length": 0.5,
    },
    "AcrobotEnv": {
        "dt": 0.2, "LINK_LENGTH_1": 1.0, "LINK_LENGTH_2": 1.0, "LINK_MASS_1": 1.0,
        "LINK_MASS_2": 1.0, "LINK_COM_POS_1": 0.5, "LINK_COM_POS_2": 0.5, "LINK_MOI": 1.0,
    },
    "MountainCarEnv": {"gravity": 0.0025, "force": 0.001},
    "Continuous_MountainCarEnv": {"power": 0.0015},
    "PendulumEnv": {"m": 1.0, "l": 1.0, "dt": 0.05, "g": 10.0},
    "FrozenLakeEnv": {"P": [1.0, 0.0, 0.0]},
    "CliffWalkingEnv": {"P": [1.0, 0.0, 0.0, 0.0]},
    "Bridge": {"P": [1.0, 0.0, 0.0], "P_left": [1.0, 0.0, 0.0], "P_right": [1.0, 0.0, 0.0]},
}

SUPPORTED_CLASSIC_CONTROL_ENV_IDS = [
    "CartPole-v1", "Acrobot-v1", "MountainCar-v0", "MountainCarContinuous-v0", "Pendulum-v1",
]
SUPPORTED_GRID_WORLD_ENV_IDS = ["CliffWalking-v1", "FrozenLake-v1", "ns_gym/Bridge-v0"]
