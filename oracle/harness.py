"""Build matched batches of envs -- the REAL reference vs the port -- under stream injection
(TEST INFRASTRUCTURE -- see oracle/__init__.py).

A *case* is ``dict(env_id=..., params=lambda S, U: {...}, wrapper={...}, make={...})`` where
``S``/``U`` are a schedulers / update-functions namespace: the reference's own modules, or
``ns_gym_b200``'s descriptions -- the same case text drives every implementation.
"""
from __future__ import annotations

import numpy as np

from . import ref_loader
from . import streams as S_
from .ns_port import NSEnvPort


def params_of(case: dict, S, U, i: int) -> dict:
    """tunable_params of env ``i``: heterogeneous cases (BASELINE config C4) build one dict per env."""
    return case["params_of"](S, U, i) if "params_of" in case else case["params"](S, U)


def _inject(tunable_params: dict, st: S_.EnvStreams):
    for slot, fn in enumerate(tunable_params.values()):
        if hasattr(fn, "rng"):
            fn.rng = S_.SlotRng(st, slot)
        if hasattr(fn.scheduler, "rng"):
            fn.scheduler.rng = S_.SlotRng(st, slot)
        inner = getattr(fn, "update_fn", None)          # LCBoundedDistrubutionUpdate's inner RandomCategorical
        if inner is not None and hasattr(inner, "rng"):
            inner.rng = S_.SlotRng(st, slot)


def make_streams(seed, n_envs, n_rows, n_slots):
    """(clock, [EnvStreams per env], uniforms[K,L,N], normals[K,P,N])."""
    u, z = S_.draw_tables(seed, n_envs, n_rows, n_slots)
    clock = S_.Clock()
    per_env = [S_.EnvStreams(u[:, :, i], z[:, :, i], clock) for i in range(n_envs)]
    return clock, per_env, u, z


def reference_envs(case: dict, n_envs: int, env_streams=None):
    """``n_envs`` instances of the reference's wrapper stack for ``case``."""
    ns = ref_loader.load()
    gym = ref_loader.gym()
    import ns_gym.schedulers as RS
    import ns_gym.update_functions as RU
    from ns_gym.wrappers import (NSBridgeWrapper, NSClassicControlWrapper,
                                 NSCliffWalkingWrapper, NSFrozenLakeWrapper)

    env_id = case["env_id"]
    make_kw = dict(case.get("make", {}))
    if "FrozenLake" in env_id:
        wrapper_cls = NSFrozenLakeWrapper
        make_kw.setdefault("is_slippery", False)
    elif "CliffWalking" in env_id:
        wrapper_cls = NSCliffWalkingWrapper
    elif "Bridge" in env_id:
        wrapper_cls = NSBridgeWrapper
    else:
        wrapper_cls = NSClassicControlWrapper
    envs = []
    for i in range(n_envs):
        tp = params_of(case, RS, RU, i)
        if env_streams is not None:
            _inject(tp, env_streams[i])          # before the wrapper clones its template
        env = gym.make(env_id, **make_kw)
        if env_streams is not None:
            env.unwrapped.np_random = S_.EnvNpRandom(env_streams[i])
        envs.append(wrapper_cls(env, tp, **case.get("wrapper", {})))
    return envs


def port_envs(case: dict, n_envs: int, env_streams=None, namespaces=None):
    """``n_envs`` ``NSEnvPort`` instances for ``case``.  ``namespaces`` = (S, U) modules used
    to build the parameter objects (default: the ns_gym_b200 descriptions)."""
    if namespaces is None:
        import ns_gym_b200.schedulers as PS
        import ns_gym_b200.update_functions as PU
        namespaces = (PS, PU)
    envs = []
    for i in range(n_envs):
        tp = params_of(case, *namespaces, i)
        envs.append(NSEnvPort(case["env_id"], tp, streams=None if env_streams is None else env_streams[i],
                              **case.get("wrapper", {}), **case.get("make", {})))
    return envs


def planning_envs(envs, env_streams=None):
    """``get_planning_env()`` of every env (reference wrappers or ports).  The reference reseeds
    the copy's generators with fresh entropy (``_reseed_planning_env_rngs``) and builds a new base
    env; under injection both are pointed back at the env's pre-drawn tables."""
    plans = []
    for i, e in enumerate(envs):
        p = e.get_planning_env()
        if env_streams is not None and not isinstance(p, NSEnvPort):
            st = env_streams[i]
            for slot, fn in enumerate(p.tunable_params.values()):
                if hasattr(fn, "rng"):
                    fn.rng = S_.SlotRng(st, slot)
                inner = getattr(fn, "update_fn", None)
                if inner is not None and hasattr(inner, "rng"):
                    inner.rng = S_.SlotRng(st, slot)
            p.unwrapped.np_random = S_.EnvNpRandom(st)
        plans.append(p)
    return plans


def draw_actions(case: dict, seed: int, n_steps: int, n_envs: int):
    """Actions [K, N] (int64) or [K, N, 1] float64 for Box action spaces."""
    rng = np.random.default_rng(seed)
    env_id = case["env_id"]
    if "Pendulum" in env_id:
        return rng.uniform(-2.0, 2.0, size=(n_steps, n_envs, 1))
    if "MountainCarContinuous" in env_id:
        return rng.uniform(-1.0, 1.0, size=(n_steps, n_envs, 1))
    n_act = 2 if "CartPole" in env_id else 3 if ("Acrobot" in env_id or "MountainCar" in env_id) else 4
    return rng.integers(0, n_act, size=(n_steps, n_envs))
