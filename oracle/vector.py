"""Vector-env loops over per-env objects (TEST INFRASTRUCTURE -- see oracle/__init__.py).

``SyncVector`` restates ``gymnasium.vector.SyncVectorEnv`` with the 1.x default
``AutoresetMode.NEXT_STEP``: a sub-env that ended at call k is *reset*
on call k+1 -- its action is ignored and it returns the reset observation with reward 0 and
both flags False.  It drives either the reference's wrappers or ``ns_port.NSEnvPort``.

``trace()`` records everything the parity tests compare, in arrays shaped [K, N, ...].
``run_parallel()`` is the multi-process variant used only as the timed CPU baseline.
"""
from __future__ import annotations

import multiprocessing as mp
import time
import warnings
from contextlib import contextmanager

import numpy as np

from . import streams as S


@contextmanager
def patched_global_choice(current_streams):
    """Route Bridge's ``np.random.choice`` (envs/Bridge.py:95-97, global legacy RNG) to the
    injected uniform of whichever env is stepping.  ``current_streams`` is a 1-element list
    holding that env's ``EnvStreams``."""
    from .ns_port import legacy_choice_index

    real = np.random.choice

    def choice(a, size=None, replace=True, p=None):
        st = current_streams[0]
        if st is None or p is None or size is not None:
            return real(a, size=size, replace=replace, p=p)
        return a[legacy_choice_index(p, st.uniform(S.LANE_DYN))]

    np.random.choice = choice
    try:
        yield
    finally:
        np.random.choice = real


def _theta_of(env):
    """Ground-truth parameter values of one env (reference wrapper or port)."""
    if hasattr(env, "theta"):
        return env.theta()
    base = env.unwrapped
    out = {}
    for k in env.tunable_params:
        if k == "P" and hasattr(env, "transition_prob"):
            out[k] = list(env.transition_prob)
        else:
            v = getattr(base, k)
            out[k] = list(v) if isinstance(v, (list, tuple)) else v
    return out


def _raw_state_of(env):
    base = env.base if hasattr(env, "base") else env.unwrapped
    if hasattr(base, "state") and base.state is not None:
        return np.asarray(base.state, dtype=np.float64)
    return np.asarray(int(base.s))


class SyncVector:
    def __init__(self, envs, env_streams=None, clock=None, autoreset=True):
        self.autoreset = autoreset          # False: single-env semantics, stepping past an end allowed
        self.envs = list(envs)
        self.n = len(self.envs)
        self.env_streams = env_streams
        self.clock = clock
        self.needs_reset = np.zeros(self.n, dtype=bool)
        self._cur = [None]
        e0 = self.envs[0]
        self.keys = list(e0.keys) if hasattr(e0, "keys") and not callable(e0.keys) else list(e0.tunable_params)

    def _set_cur(self, i):
        if self.env_streams is not None:
            self._cur[0] = self.env_streams[i]

    def reset(self, k=0):
        """Reset every env.  Under injection the reset draws use table row ``k``."""
        if self.clock is not None:
            self.clock.k = k
        outs = []
        with patched_global_choice(self._cur):
            for i, e in enumerate(self.envs):
                self._set_cur(i)
                outs.append(e.reset())
        self.needs_reset[:] = False
        return outs

    def step(self, actions, k=None):
        if self.clock is not None and k is not None:
            self.clock.k = k
        outs = []
        with patched_global_choice(self._cur), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for i, e in enumerate(self.envs):
                self._set_cur(i)
                if self.needs_reset[i]:
                    obs, info = e.reset()
                    outs.append((obs, 0.0, False, False, info, True))
                else:
                    a = actions[i]
                    if isinstance(a, np.ndarray) and a.ndim == 0:
                        a = a.item()
                    obs, r, term, trunc, info = e.step(a)
                    outs.append((obs, r, term, trunc, info, False))
        self.needs_reset = np.array([(o[2] or o[3]) and self.autoreset for o in outs])
        return outs


def trace(vec: SyncVector, actions, first_row=1, do_reset=True):
    """Run ``len(actions)`` vector steps after an initial reset (table row 0) and record
    arrays [K, N, ...].  ``actions`` is [K, N] (or [K, N, A] for Box actions).
    ``do_reset=False`` continues from the envs' current state (planning copies)."""
    K, N = len(actions), vec.n
    keys = vec.keys
    rec = {}
    if do_reset:
        r0 = vec.reset(k=0)
        rec["obs0"] = np.stack([np.asarray(o[0]["state"]) for o in r0])
        rec["raw0"] = np.stack([_raw_state_of(e) for e in vec.envs])
    rec.update({
        "reward": np.zeros((K, N)),
        "terminated": np.zeros((K, N), dtype=bool),
        "truncated": np.zeros((K, N), dtype=bool),
        "was_reset": np.zeros((K, N), dtype=bool),
        "relative_time": np.zeros((K, N), dtype=np.int64),
        "env_change": np.zeros((K, N, len(keys)), dtype=np.int64),
        "delta_change": np.zeros((K, N, len(keys))),
        "gt_change": np.zeros((K, N, len(keys)), dtype=np.int64),
        "gt_delta": np.zeros((K, N, len(keys))),
    })
    obs_l, raw_l, theta_l = [], [], []
    for k in range(K):
        outs = vec.step(actions[k], k=first_row + k)
        obs_l.append(np.stack([np.asarray(o[0]["state"]) for o in outs]))
        raw_l.append(np.stack([_raw_state_of(e) for e in vec.envs]))
        th = []
        for e in vec.envs:
            t = _theta_of(e)
            th.append([np.atleast_1d(np.asarray(t[key], dtype=np.float64)) for key in keys])
        theta_l.append(np.stack([np.concatenate(row) for row in th]))
        for i, (obs, r, term, trunc, info, was_reset) in enumerate(outs):
            rec["reward"][k, i] = r if not isinstance(r, dict) else r["reward"]
            rec["terminated"][k, i] = term
            rec["truncated"][k, i] = trunc
            rec["was_reset"][k, i] = was_reset
            rec["relative_time"][k, i] = obs["relative_time"]
            for j, key in enumerate(keys):
                rec["env_change"][k, i, j] = obs["env_change"][key]
                rec["delta_change"][k, i, j] = obs["delta_change"][key]
                rec["gt_change"][k, i, j] = info["Ground Truth Env Change"][key]
                rec["gt_delta"][k, i, j] = info["Ground Truth Delta Change"][key]
    rec["obs"] = np.stack(obs_l)
    rec["raw"] = np.stack(raw_l)
    rec["theta"] = np.stack(theta_l)
    return rec


# --------------------------------------------------------------------------------------
# timed CPU baseline
# --------------------------------------------------------------------------------------


def _time_sync(make_envs, n_envs, n_steps, action_fn, seed):
    envs = make_envs(n_envs)
    vec = SyncVector(envs)
    rng = np.random.default_rng(seed)
    vec.reset()
    t0 = time.perf_counter()
    for _ in range(n_steps):
        vec.step(action_fn(rng, n_envs))
    return time.perf_counter() - t0


def _worker(args):
    make_envs, n_envs, n_steps, action_fn, seed = args
    return _time_sync(make_envs, n_envs, n_steps, action_fn, seed)


def run_sync(make_envs, n_envs, n_steps, action_fn, seed=0):
    """env-steps/s of the single-core Sync loop."""
    dt = _time_sync(make_envs, n_envs, n_steps, action_fn, seed)
    return n_envs * n_steps / dt


def run_parallel(make_envs, n_envs_per_proc, n_steps, action_fn, n_procs, seed=0):
    """env-steps/s with one Sync loop per process (the AsyncVectorEnv-style arrangement:
    envs partitioned over worker processes).  Each worker times its own stepping loop (env
    construction and process start-up excluded, as for the GPU arm); throughput = total
    steps / slowest worker."""
    ctx = mp.get_context("fork")
    with ctx.Pool(n_procs) as pool:
        times = pool.map(_worker, [(make_envs, n_envs_per_proc, n_steps, action_fn, seed + i)
                                   for i in range(n_procs)])
    return n_procs * n_envs_per_proc * n_steps / max(times)
