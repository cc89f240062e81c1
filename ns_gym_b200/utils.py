"""Helpers of the reference's plugin surface that callers of the step path use
(``ns_gym/utils.py``)."""
from __future__ import annotations

from . import base


def type_mismatch_checker(observation=None, reward=None):
    """``ns_gym.utils.type_mismatch_checker`` (``utils.py:122-152``): strip the NS packaging --
    the observation dict down to its ``"state"`` entry, a ``Reward`` down to its ``reward`` -- so that
    agents written for plain gymnasium (``evaluate/run_experiment.py:91-148``, ``MCTS.py:177``)
    consume the batched results; the values stay tensors with one entry per env."""
    obs = None
    if observation is not None:
        obs = observation["state"] if isinstance(observation, dict) and "state" in observation else observation
    rew = None
    if reward is not None:
        rew = reward.reward if isinstance(reward, base.Reward) else reward
    assert not isinstance(obs, dict), "Observation is still a dict after type checking."
    return obs, rew
