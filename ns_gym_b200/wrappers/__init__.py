"""Batched counterparts of the reference's NS wrappers (``ns_gym/wrappers/__init__.py:2-15``).

Usage mirrors the reference, with ``ns_gym_b200.make`` standing in for ``gym.make``::

    env = ns_gym_b200.make("CartPole-v1", num_envs=1 << 20)
    env = NSClassicControlWrapper(env, {"masspole": IncrementUpdate(ContinuousScheduler(), k=0.1)},
                                  change_notification=True)
    obs, info = env.reset(seed=0)
    obs, reward, terminated, truncated, info = env.step(actions)     # tensors of shape [N, ...]
"""
from __future__ import annotations

from dataclasses import dataclass, field

from ..compile import ENV_TABLE
from ..vector_env import NSVectorEnv


class ConstraintViolationWarning(Warning):
    """Kept for API parity (``classic_control.py:9-12``).  The device path rejects a violating
    update silently (flag and delta zeroed, old value kept), exactly as the reference does
    apart from the Python warning."""


@dataclass
class BatchedEnvSpec:
    """What ``ns_gym_b200.make`` returns: the id, batch size and base-env options of a batch
    that does not exist on the device until an NS wrapper is applied."""
    env_id: str
    num_envs: int
    kwargs: dict = field(default_factory=dict)


def make(env_id: str, num_envs: int = 1, **kwargs) -> BatchedEnvSpec:
    if env_id not in ENV_TABLE:
        raise KeyError(f"unknown environment id {env_id!r}; supported: {sorted(ENV_TABLE)}")
    return BatchedEnvSpec(env_id, int(num_envs), dict(kwargs))


def _split(env, num_envs):
    if isinstance(env, BatchedEnvSpec):
        return env.env_id, env.num_envs, dict(env.kwargs)
    if isinstance(env, str):
        if num_envs is None:
            raise TypeError("num_envs is required when the env is given by id")
        return env, int(num_envs), {}
    raise TypeError("expected ns_gym_b200.make(...) or an environment id; a live gymnasium env cannot "
                    "be batched onto the device")


class _KindChecked(NSVectorEnv):
    _KINDS: tuple = ()

    def __init__(self, env, tunable_params, change_notification=False, delta_change_notification=False,
                 in_sim_change=False, *, num_envs=None, **kwargs):
        env_id, n, make_kw = _split(env, num_envs)
        cls = ENV_TABLE[env_id][1]
        assert cls in self._KINDS, f"{cls} is not a supported environment"      # classic_control.py:36-38
        make_kw.update(kwargs)
        make_kw.pop("is_slippery", None)   # the NS wrapper owns the slip distribution
        make_kw.pop("render_mode", None)
        super().__init__(env_id, tunable_params, n, change_notification=change_notification,
                         delta_change_notification=delta_change_notification,
                         in_sim_change=in_sim_change, **make_kw)


class NSClassicControlWrapper(_KindChecked):
    """``ns_gym/wrappers/classic_control.py:15-458`` for a batch."""
    _KINDS = ("CartPoleEnv", "AcrobotEnv", "MountainCarEnv", "Continuous_MountainCarEnv", "PendulumEnv")


class NSFrozenLakeWrapper(_KindChecked):
    """``ns_gym/wrappers/toy_text.py:265-519`` for a batch (``initial_prob_dist``,
    ``modified_rewards``, ``map_name`` / ``desc`` keywords)."""
    _KINDS = ("FrozenLakeEnv",)


class NSCliffWalkingWrapper(_KindChecked):
    """``ns_gym/wrappers/toy_text.py:14-262`` for a batch (``terminal_cliff`` keyword)."""
    _KINDS = ("CliffWalkingEnv",)


class NSBridgeWrapper(_KindChecked):
    """``ns_gym/wrappers/toy_text.py:524-715`` for a batch (uniform ``P`` or split
    ``P_left`` / ``P_right``)."""
    _KINDS = ("Bridge",)


__all__ = ["NSClassicControlWrapper", "NSFrozenLakeWrapper", "NSCliffWalkingWrapper", "NSBridgeWrapper",
           "ConstraintViolationWarning", "BatchedEnvSpec", "make"]
