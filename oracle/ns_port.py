"""CPU restatement of the reference's own NS layer (TEST INFRASTRUCTURE -- oracle/__init__.py).

PARITY STATUS: **PINNED** against the reference imported verbatim in this container
(``oracle/ref_loader.py``; ``tests/test_oracle_vs_reference.py``) and against the golden
vectors under ``tests/golden/`` generated from that import.

This is not a copy of the reference's classes: the reference is an object protocol
(``UpdateFn.__call__`` -> ``Scheduler.__call__`` -> ``_update``); the port is a flat
interpreter over *descriptions* of those objects (``describe``), with all mutable state held
in ``SlotState`` so that one description can drive many envs.  Each rule cites the reference
line it restates.

Covered: a1 schedulers, a2 scalar update functions, a3 distribution update
functions + W1 delta, a4 NSClassicControlWrapper.step (constraints, CartPole resolver),
a8 NSFrozenLakeWrapper, a9 NSBridgeWrapper + Bridge env, a10 NSCliffWalkingWrapper,
a11 NSWrapper.step notification gating, a12 TimeLimit, a13 reset.
"""
from __future__ import annotations

import copy
import math
import warnings

import numpy as np

from . import gym_restated as G
from .streams import LANE_SCHED0, EnvStreams, geometric_from_uniform

INF = float("inf")

# --------------------------------------------------------------------------------------
# describe(): duck-typed snapshot of a scheduler / update-function object
# --------------------------------------------------------------------------------------

_SCHED_ATTRS = {
    "ContinuousScheduler": (),
    "PeriodicScheduler": ("period",),
    "DiscreteScheduler": ("event_list",),
    "BurstScheduler": ("on_duration", "off_duration", "cycle"),
    "WindowScheduler": ("windows",),
    "RandomScheduler": ("probability",),
    "DecayingProbabilityScheduler": ("initial_probability", "decay_rate"),
    "MemorylessScheduler": ("p", "transition_time"),
    "CustomScheduler": ("event_function",),
}

_UPD_ATTRS = {
    "IncrementUpdate": ("k",),
    "DecrementUpdate": ("k",),
    "DeterministicTrend": ("slope",),
    "PolynomialTrend": ("coeffs",),
    "GeometricProgression": ("r",),
    "ExponentialDecay": ("decay_rate",),
    "OscillatingUpdate": ("delta",),
    "SigmoidTransition": ("a", "b", "k", "t0"),
    "LinearInterpolation": ("start_val", "end_val", "T"),
    "StepWiseUpdate": ("param_list",),
    "CyclicUpdate": ("value_list",),
    "NoUpdate": (),
    "RandomWalk": ("mu", "sigma"),
    "RandomWalkWithDrift": ("alpha", "mu", "sigma"),
    "RandomWalkWithDriftAndTrend": ("alpha", "mu", "sigma", "slope"),
    "OrnsteinUhlenbeck": ("theta", "mu", "sigma"),
    "BoundedRandomWalk": ("mu", "sigma", "lo", "hi"),
    "DistributionIncrementUpdate": ("k",),
    "DistributionDecrementUpdate": ("k",),
    "DistributionStepWiseUpdate": ("update_values",),
    "DistributionCyclicUpdate": ("dist_list",),
    "DistributionNoUpdate": (),
    "UniformDrift": ("rate",),
    "TargetReversion": ("target", "theta"),
    "DistributionLinearInterpolation": ("start_dist", "end_dist", "T"),
    "RandomCategorical": (),
    "LCBoundedDistrubutionUpdate": ("L",),
}

STOCHASTIC_UPDATES = {
    "RandomWalk", "RandomWalkWithDrift", "RandomWalkWithDriftAndTrend",
    "OrnsteinUhlenbeck", "BoundedRandomWalk", "RandomCategorical", "LCBoundedDistrubutionUpdate",
}
STOCHASTIC_SCHEDS = {"RandomScheduler", "DecayingProbabilityScheduler", "MemorylessScheduler"}


def _rng_of(obj, stochastic: bool):
    """The object's own generator (reference classes), else one built from its ``seed``
    attribute (ns_gym_b200 descriptions), else None for deterministic rules."""
    if hasattr(obj, "rng"):
        return copy.deepcopy(obj.rng)
    if stochastic:
        return np.random.default_rng(seed=getattr(obj, "seed", None))
    return None


_STATEFUL_UPDATES = STOCHASTIC_UPDATES | {
    "StepWiseUpdate", "CyclicUpdate", "DistributionStepWiseUpdate", "DistributionCyclicUpdate"}


def check_no_stateful_aliasing(tunable_params: dict) -> None:
    """The reference lets one *stateful* object serve several parameters, interleaving calls
    on the shared cursor / RNG.  Neither the port nor the CUDA path emulates
    that; both reject it."""
    seen_fn, seen_sched = {}, {}
    for key, fn in tunable_params.items():
        if type(fn).__name__ in _STATEFUL_UPDATES:
            if id(fn) in seen_fn:
                raise NotImplementedError(
                    f"stateful update function shared by {seen_fn[id(fn)]!r} and {key!r}")
            seen_fn[id(fn)] = key
        sch = fn.scheduler
        if type(sch).__name__ in STOCHASTIC_SCHEDS | {"CustomScheduler"}:
            if id(sch) in seen_sched:
                raise NotImplementedError(
                    f"stateful scheduler shared by {seen_sched[id(sch)]!r} and {key!r}")
            seen_sched[id(sch)] = key


def describe(fn) -> dict:
    """Plain-dict description of an update function and its scheduler."""
    kind = type(fn).__name__
    if kind not in _UPD_ATTRS:
        raise TypeError(f"oracle: update function {kind} is not on the restated path")
    d = {"kind": kind, "obj_id": id(fn)}
    for a in _UPD_ATTRS[kind]:
        d[a] = copy.deepcopy(getattr(fn, a))
    d["rng"] = _rng_of(fn, kind in STOCHASTIC_UPDATES)
    if kind == "LCBoundedDistrubutionUpdate":       # distribution.py:155-164: inner rule, default RandomCategorical
        inner = getattr(fn, "update_fn", None)      # instance (reference) | class or None (descriptions)
        name = "RandomCategorical" if inner is None else (
            inner.__name__ if isinstance(inner, type) else type(inner).__name__)
        if name != "RandomCategorical":
            raise TypeError("oracle: LCBoundedDistrubutionUpdate is restated for its default inner rule only")
        d["inner"] = {"kind": "RandomCategorical"}
        d["rng"] = copy.deepcopy(inner.rng) if hasattr(inner, "rng") else np.random.default_rng()
    sch = fn.scheduler
    skind = type(sch).__name__
    if skind not in _SCHED_ATTRS:
        raise TypeError(f"oracle: scheduler {skind} is not on the restated path")
    s = {"kind": skind, "start": sch.start, "end": sch.end, "obj_id": id(sch)}
    for a in _SCHED_ATTRS[skind]:
        v = getattr(sch, a)
        s[a] = v if a == "event_function" else copy.deepcopy(v)
    s["rng"] = _rng_of(sch, skind in STOCHASTIC_SCHEDS)
    d["sched"] = s
    return d


class SlotState:
    """Mutable per-(env, parameter) state: list cursors, Memoryless clock, RNG handles."""

    __slots__ = ("queue", "index", "transition_time", "fn_rng", "sched_rng", "prev_param", "prev_time")

    def __init__(self, desc: dict):
        kind = desc["kind"]
        self.queue = None
        if kind == "StepWiseUpdate":
            self.queue = list(desc["param_list"])
        elif kind == "DistributionStepWiseUpdate":
            self.queue = [list(v) for v in desc["update_values"]]
        self.index = 0
        tt = desc["sched"].get("transition_time")
        self.transition_time = None if tt is None else int(np.asarray(tt).reshape(-1)[0])
        self.fn_rng = copy.deepcopy(desc.get("rng"))
        self.sched_rng = copy.deepcopy(desc["sched"].get("rng"))
        self.prev_param = None
        self.prev_time = -1


# --------------------------------------------------------------------------------------
# a1: scheduler fire test  (base.py:67-81 gate, schedulers.py _check rules)
# --------------------------------------------------------------------------------------


def sched_fires(s: dict, st: SlotState, t, streams: EnvStreams | None, slot: int) -> bool:
    if not (s["start"] <= t <= s["end"]):          # base.py:79-81, inclusive both ends
        return False
    k = s["kind"]
    if k == "ContinuousScheduler":                  # schedulers.py:52-53
        return True
    if k == "PeriodicScheduler":                    # schedulers.py:88-89
        return t % s["period"] == 0
    if k == "DiscreteScheduler":                    # schedulers.py:73-74
        return t in s["event_list"]
    if k == "BurstScheduler":                       # schedulers.py:139-140
        return (t % s["cycle"]) < s["on_duration"]
    if k == "WindowScheduler":                      # schedulers.py:197-198
        return any(a <= t <= b for a, b in s["windows"])
    if k == "CustomScheduler":                      # schedulers.py:42-43
        return bool(s["event_function"](t))

    def uni():
        if streams is not None:
            return streams.sched_uniform(slot, t)
        return st.sched_rng.random()

    if k == "RandomScheduler":                      # schedulers.py:27-28
        return uni() < s["probability"]
    if k == "DecayingProbabilityScheduler":         # schedulers.py:175-177
        p = s["initial_probability"] * np.exp(-s["decay_rate"] * t)
        return bool(uni() < p)
    if k == "MemorylessScheduler":                  # schedulers.py:110-116
        if t == st.transition_time:
            if streams is not None:
                g = geometric_from_uniform(streams.sched_uniform(slot, t), s["p"])
            else:
                g = int(st.sched_rng.geometric(p=s["p"], size=(1,))[0])
            st.transition_time = g + t
            return True
        return False
    raise TypeError(k)


# --------------------------------------------------------------------------------------
# a2 / a3: update rules
# --------------------------------------------------------------------------------------


def w1_index_distance(u, v) -> float:
    """utils.py:55-94 -> scipy.stats.wasserstein_distance(arange(n), arange(n), u, v).

    Closed form of scipy's CDF algorithm for values on the integer grid: unit gaps between
    distinct support points, CDFs normalised by each weight vector's own sum.  scipy raises
    ValueError for negative / non-finite / all-zero weights; so does this.
    """
    u = [float(x) for x in u]
    v = [float(x) for x in v]
    if len(u) != len(v):
        raise ValueError("wasserstein_distance: u and v must have the same shape")
    for w in (u, v):
        if any(x < 0 for x in w):
            raise ValueError("All weights must be non-negative.")
        if not (0 < sum(w) < INF):
            raise ValueError("Weight array-like sum must be positive and finite.")
    cu = np.cumsum(np.asarray(u))
    cv = np.cumsum(np.asarray(v))
    acc = 0.0
    for i in range(len(u) - 1):
        acc = acc + abs(cu[i] / cu[-1] - cv[i] / cv[-1])
    return float(acc)


def _normal(desc, st, streams, slot, mu, sigma):
    if streams is not None:
        return mu + sigma * streams.std_normal(slot)
    return st.fn_rng.normal(mu, sigma)


def apply_update(d: dict, st: SlotState, param, t, streams, slot):
    """``_update(copy.copy(param), t)`` of every class in single_param.py / distribution.py."""
    k = d["kind"]
    # ---- scalar ----
    if k == "IncrementUpdate":                      # single_param.py:173-175
        return param + d["k"]
    if k == "DecrementUpdate":                      # :197-199
        return param - d["k"]
    if k == "DeterministicTrend":                   # :38-40
        return param + d["slope"] * t
    if k == "PolynomialTrend":                      # :471-473  (python sum, starts from int 0)
        trend = sum(a * t ** (i + 1) for i, a in enumerate(d["coeffs"]))
        return param + trend
    if k == "GeometricProgression":                 # :305-307
        return param * d["r"]
    if k == "ExponentialDecay":                     # :285-287  (compounds on the current value)
        return param * np.exp(-d["decay_rate"] * t)
    if k == "OscillatingUpdate":                    # :262-264
        return param + d["delta"] * np.sin(t)
    if k == "SigmoidTransition":                    # :383-385
        sig = 1.0 / (1.0 + np.exp(-d["k"] * (t - d["t0"])))
        return d["a"] + (d["b"] - d["a"]) * sig
    if k == "LinearInterpolation":                  # :506-508
        frac = min(t / d["T"], 1.0)
        return d["start_val"] + (d["end_val"] - d["start_val"]) * frac
    if k in ("StepWiseUpdate", "DistributionStepWiseUpdate"):
        # :217-223 / distribution.py:116-130 -- pop(0); the ``finally: return`` swallows the
        # IndexError of an exhausted list, leaving the value unchanged (flag stays 1)
        if st.queue:
            return st.queue.pop(0)
        return param
    if k == "CyclicUpdate":                         # :405-408
        vals = d["value_list"]
        v = vals[st.index]
        st.index = (st.index + 1) % len(vals)
        return v
    if k in ("NoUpdate", "DistributionNoUpdate"):   # :239-240 / distribution.py:230-231
        return param
    if k == "RandomWalk":                           # :110-113
        return param + _normal(d, st, streams, slot, d["mu"], d["sigma"])
    if k == "RandomWalkWithDrift":                  # :148-151
        return d["alpha"] + param + _normal(d, st, streams, slot, d["mu"], d["sigma"])
    if k == "RandomWalkWithDriftAndTrend":          # :78-81
        wn = _normal(d, st, streams, slot, d["mu"], d["sigma"])
        return d["alpha"] + param + wn + d["slope"] * t
    if k == "OrnsteinUhlenbeck":                    # :344-346  (no draw when sigma == 0)
        noise = _normal(d, st, streams, slot, 0, d["sigma"]) if d["sigma"] > 0 else 0.0
        return param + d["theta"] * (d["mu"] - param) + noise
    if k == "BoundedRandomWalk":                    # :446-448
        noise = _normal(d, st, streams, slot, d["mu"], d["sigma"])
        return float(np.clip(param + noise, d["lo"], d["hi"]))
    # ---- distributions (param is a list) ----
    if k == "DistributionIncrementUpdate":          # distribution.py:61-67 (no lower clamp)
        p = list(param)
        p[0] = min(1, p[0] + d["k"])
        for i in range(1, len(p)):
            p[i] = (1 - p[0]) / (len(p) - 1)
        return p
    if k == "DistributionDecrementUpdate":          # :88-97
        p = list(param)
        p[0] = max(0, p[0] - d["k"])
        for i in range(1, len(p)):
            p[i] = (1 - p[0]) / (len(p) - 1)
        return p
    if k == "DistributionCyclicUpdate":             # :353-356
        vals = d["dist_list"]
        v = vals[st.index]
        st.index = (st.index + 1) % len(vals)
        return v
    if k == "UniformDrift":                         # :256-261
        n = len(param)
        uniform = 1.0 / n
        return [(1 - d["rate"]) * p + d["rate"] * uniform for p in param]
    if k == "TargetReversion":                      # :289-293
        return [p + d["theta"] * (tg - p) for p, tg in zip(param, d["target"])]
    if k == "DistributionLinearInterpolation":      # :326-331
        frac = min(t / d["T"], 1.0)
        return [s + (e - s) * frac for s, e in zip(d["start_dist"], d["end_dist"])]
    if k == "RandomCategorical":                    # :37-38
        if streams is not None:
            return streams.dirichlet(slot, len(param))
        return list(st.fn_rng.dirichlet(np.ones(len(param))))
    if k == "LCBoundedDistrubutionUpdate":          # :166-183 rejection loop around the inner rule
        inner = d["inner"]
        cur = np.asarray(param, dtype=float)
        bound = d["L"] * abs(t - st.prev_time)
        for _ in range(int(1e5)):
            cand = apply_update(inner, st, list(param), t, streams, slot)
            if w1_index_distance(cur, np.asarray(cand, dtype=float)) <= bound:
                return cand
        raise ValueError("Could not find a Lipschitz-continuous update")
    raise TypeError(k)


def call_update(d: dict, st: SlotState, param, t, streams, slot):
    """UpdateFn.__call__ (base.py:124-149): (new, flag, delta)."""
    is_dist = isinstance(param, list)
    if sched_fires(d["sched"], st, t, streams, slot):
        new = apply_update(d, st, copy.copy(param), t, streams, slot)
        if is_dist:                                  # base.py:192-203
            delta = w1_index_distance(param, new)
        else:                                        # base.py:172-182
            delta = new - param
        st.prev_param, st.prev_time = param, t
        return new, 1, delta
    st.prev_param, st.prev_time = param, t
    return param, 0, 0.0


# --------------------------------------------------------------------------------------
# a4: constraint rules (classic_control.py:193-422) and resolver (:424-444)
# --------------------------------------------------------------------------------------


def constraint_violations(env_name: str, env, new_vals: dict) -> dict:
    out = {}
    if env_name == "CartPoleEnv":                    # :208-235
        for p, v in new_vals.items():
            out[p] = bool(
                (p in ("length", "masscart", "masspole") and v <= 0) or (p == "gravity" and v < 0)
            )
    elif env_name == "AcrobotEnv":                   # :237-357
        for p, v in new_vals.items():
            bad = False
            if p == "LINK_LENGTH_1":
                if v <= 0:
                    bad = True
                elif "LINK_COM_POS_1" in new_vals and new_vals["LINK_COM_POS_1"] > v:
                    bad = True
                elif v < env.LINK_COM_POS_1:
                    bad = True
            elif p == "LINK_LENGTH_2":               # :267 -- only the <= 0 branch is reachable
                bad = v <= 0
            elif p in ("LINK_MASS_1", "LINK_MASS_2"):
                bad = v <= 0
            elif p in ("LINK_COM_POS_1", "LINK_COM_POS_2"):
                partner = "LINK_LENGTH_1" if p.endswith("1") else "LINK_LENGTH_2"
                if v <= 0:
                    bad = True
                elif partner in new_vals and new_vals[partner] < v:
                    bad = True
                elif v > getattr(env, partner):
                    bad = True
            out[p] = bool(bad)
    elif env_name == "MountainCarEnv":               # :359-376
        for p, v in new_vals.items():
            out[p] = bool(p in ("gravity", "force") and v <= 0)
    elif env_name == "Continuous_MountainCarEnv":    # :378-387
        for p, v in new_vals.items():
            out[p] = bool(p == "power" and v <= 0)
    elif env_name == "PendulumEnv":                  # :389-420
        for p, v in new_vals.items():
            out[p] = bool((p in ("m", "l", "dt") and v <= 0) or (p == "g" and v < 0))
    return out


def resolve_dependencies(env_name: str, env) -> None:
    if env_name == "CartPoleEnv":                    # :426-444
        env.total_mass = env.masspole + env.masscart
        env.polemass_length = env.length * env.masspole


class ConstraintViolationWarning(Warning):
    pass


# --------------------------------------------------------------------------------------
# Bridge env (envs/Bridge.py) -- in-tree dynamics, restated
# --------------------------------------------------------------------------------------

BRIDGE_MAP = ["HHHHHHHH", "FFFFFHHH", "GFHFSFFG", "FFFFFHHH", "HHHHHHHH"]   # Bridge.py:12
_BRIDGE_DELTA = {0: (0, -1), 1: (1, 0), 2: (0, 1), 3: (-1, 0)}                # Bridge.py:14-17,82-87


def legacy_choice_index(p, u: float) -> int:
    """numpy legacy ``RandomState.choice(a, p=p)`` for one draw: validate p, then
    ``searchsorted(cumsum(p)/cumsum(p)[-1], u, side='right')``."""
    p = np.array(p, dtype=np.float64)
    if np.isnan(p).any() or (p < 0).any():
        raise ValueError("probabilities are not non-negative")
    if abs(float(np.sum(p)) - 1.0) > math.sqrt(np.finfo(np.float64).eps):
        raise ValueError("probabilities do not sum to 1")
    cdf = p.cumsum()
    cdf /= cdf[-1]
    return int(cdf.searchsorted(u, side="right"))


class BridgePort(G.Env):
    def __init__(self, global_init_probs=(1, 0, 0)):
        self.map = np.asarray(BRIDGE_MAP, dtype="c")
        self.nrow, self.ncol = self.map.shape
        self.nS, self.nA = self.map.size, 4
        self.P = list(global_init_probs)
        self.P_left = list(global_init_probs)
        self.P_right = list(global_init_probs)
        self.split_probs = False
        self.observation_space = G.Discrete(self.nS)
        self.action_space = G.Discrete(self.nA)
        self.s = None
        self.streams = None

    def step(self, action):
        if self.split_probs:                         # Bridge.py:90-93, 148-157
            P = self.P_left if (self.s % self.ncol) < self.ncol // 2 else self.P_right
        else:
            P = self.P
        u = self.streams.uniform(0) if self.streams is not None else np.random.random_sample()
        i = legacy_choice_index(P, u)                # Bridge.py:95-97
        a = [action, (action + 1) % 4, (action - 1) % 4][i]
        row, col = divmod(self.s, self.ncol)
        dr, dc = _BRIDGE_DELTA[a]
        nr, nc = row + dr, col + dc
        if nr < 0 or nr >= self.nrow or nc < 0 or nc >= self.ncol:   # :125-126, 136-141
            nr, nc = row, col
        cell = self.map[nr, nc]                      # :159-174 reward from the destination
        if cell == b"H":
            reward, done = -1, True
        elif cell == b"G":
            reward, done = 1, True
        else:
            reward, done = 0, False
        self.s = nr * self.ncol + nc
        return int(self.s), int(reward), done, False, {"prob": P}

    def reset(self, *, seed=None, options=None):
        super().reset(seed=seed)
        self.s = 2 * self.ncol + 4                   # Bridge.py:110
        return int(self.s), {"prob": self.P[0]}


G.register("ns_gym_port/Bridge-v0", BridgePort, 100)      # ns_gym/__init__.py:17-21

# --------------------------------------------------------------------------------------
# the NS env (wrappers + NSWrapper.step/reset), one object per env
# --------------------------------------------------------------------------------------

_CLASSIC = {"CartPoleEnv", "AcrobotEnv", "MountainCarEnv", "Continuous_MountainCarEnv", "PendulumEnv"}

TUNABLE = {
    "CartPoleEnv": ["gravity", "masscart", "masspole", "force_mag", "tau", "length"],
    "AcrobotEnv": ["dt", "LINK_LENGTH_1", "LINK_LENGTH_2", "LINK_MASS_1", "LINK_MASS_2",
                   "LINK_COM_POS_1", "LINK_COM_POS_2", "LINK_MOI"],
    "MountainCarEnv": ["gravity", "force"],
    "Continuous_MountainCarEnv": ["power"],
    "PendulumEnv": ["m", "l", "dt", "g"],
    "FrozenLakeEnv": ["P"],
    "CliffWalkingEnv": ["P"],
    "BridgePort": ["P", "P_left", "P_right"],
}                                                    # base.py:611-635, 1161-1165


class NSEnvPort:
    """One non-stationary env: the reference's wrapper stack restated as a flat object.

    ``streams`` (an ``EnvStreams``) switches every random draw to the injected tables; with
    ``streams=None`` numpy generators are used exactly where the reference uses them.
    """

    def __init__(self, env_id: str, tunable_params: dict, *, change_notification=False,
                 delta_change_notification=False, in_sim_change=False, scalar_reward=True,
                 persistent_params=False, initial_prob_dist=None, modified_rewards=None,
                 terminal_cliff=False, max_episode_steps=None, streams=None, **make_kwargs):
        if delta_change_notification:               # base.py:252-255
            assert change_notification
        self._ctor = dict(env_id=env_id, tunable_params=tunable_params, change_notification=change_notification,
                          delta_change_notification=delta_change_notification, in_sim_change=in_sim_change,
                          scalar_reward=scalar_reward, persistent_params=persistent_params,
                          initial_prob_dist=initial_prob_dist, modified_rewards=modified_rewards,
                          terminal_cliff=terminal_cliff, max_episode_steps=max_episode_steps, streams=streams,
                          **make_kwargs)
        self.has_reset = False
        if env_id in ("ns_gym/Bridge-v0", "Bridge"):
            env_id = "ns_gym_port/Bridge-v0"
        if "FrozenLake" in env_id:
            make_kwargs.setdefault("is_slippery", False)
        self.env = G.make(env_id, max_episode_steps=max_episode_steps, **make_kwargs)
        self.base = self.env.unwrapped
        self.name = type(self.base).__name__
        assert set(tunable_params) <= set(TUNABLE[self.name]), (   # base.py:257-261
            f"Tunable parameters {list(tunable_params)} not all in {TUNABLE[self.name]}")
        check_no_stateful_aliasing(tunable_params)
        self.keys = list(tunable_params)
        self.desc = {k: describe(fn) for k, fn in tunable_params.items()}
        self.slot = {k: i for i, k in enumerate(self.keys)}
        self.change_notification = change_notification
        self.delta_change_notification = delta_change_notification
        self.in_sim_change = in_sim_change
        self.scalar_reward = scalar_reward
        self.persistent_params = persistent_params
        self.frozen = False
        self.is_sim_env = False
        self.streams = streams
        self.t = 0
        self.states = {k: SlotState(d) for k, d in self.desc.items()}
        if streams is not None:
            from .streams import EnvNpRandom
            self.base.np_random = EnvNpRandom(streams)
            if self.name == "BridgePort":
                self.base.streams = streams
        # --- env-kind specific initial values ---
        if self.name in _CLASSIC:                    # classic_control.py:50-58
            self.initial = {k: getattr(self.base, k) for k in self.keys}
        else:
            n_out = 4 if self.name == "CliffWalkingEnv" else 3
            default = [1, 0, 0, 0][:n_out]
            ipd = default if initial_prob_dist is None else initial_prob_dist
            self.modified_rewards = modified_rewards
            self.terminal_cliff = terminal_cliff
            if self.name == "BridgePort":            # toy_text.py:570-597
                self.split = ("P_left" in tunable_params) or ("P_right" in tunable_params)
                if isinstance(ipd, tuple) and len(ipd) == 2:
                    self.init_left, self.init_right = list(ipd[0]), list(ipd[1])
                    self.init_uniform = list(ipd[0])
                else:
                    self.init_left = list(ipd)
                    self.init_right = list(ipd)
                    self.init_uniform = list(ipd)
                self.base.split_probs = self.split
                self.base.P = list(self.init_uniform)
                self.base.P_left = list(self.init_left)
                self.base.P_right = list(self.init_right)
            else:
                if self.name == "FrozenLakeEnv":     # toy_text.py:329-334
                    assert sum(ipd) == 1 or math.isclose(sum(ipd), 1)
                    assert len(ipd) == 3
                self.initial_prob_dist = ipd
                self.transition_prob = copy.deepcopy(ipd)
                # probabilities baked into the table the base env samples from; rebuilt ONLY
                # when the update fires (toy_text.py:181-187, 365-367) and NOT on reset, so a
                # new episode keeps sampling from the stale table until the next fire.
                self.table_prob = list(self.transition_prob)

    # ---- gridworld tables -------------------------------------------------------------
    def _frozenlake_outcome(self, s, b):             # toy_text.py:449-469
        nrow, ncol, desc = self.base.nrow, self.base.ncol, self.base.desc
        row, col = divmod(s, ncol)
        if b == 0:
            col = max(col - 1, 0)
        elif b == 1:
            row = min(row + 1, nrow - 1)
        elif b == 2:
            col = min(col + 1, ncol - 1)
        else:
            row = max(row - 1, 0)
        letter = desc[row, col]
        if self.modified_rewards:
            reward = float(self.modified_rewards[letter.decode("utf-8")])
        else:
            reward = float(letter == b"G")
        return row * ncol + col, reward, bytes(letter) in b"GH"

    def _cliff_outcome(self, s, b):                  # toy_text.py:86-138
        rewards = self.modified_rewards or {"H": -100, "G": 0, "F": -1, "S": -1}
        dr, dc = [(-1, 0), (0, 1), (1, 0), (0, -1)][b]
        row, col = divmod(s, 12)
        nr = min(max(row + dr, 0), 3)
        nc = min(max(col + dc, 0), 11)
        is_cliff = nr == 3 and 1 <= nc <= 10
        is_goal = (nr, nc) == (3, 11)
        if is_cliff:
            return 36, rewards["H"], bool(self.terminal_cliff)
        return nr * 12 + nc, (rewards["G"] if is_goal else rewards["F"]), is_goal

    def _grid_step(self, action):
        """FrozenLakeEnv.step / CliffWalkingEnv.step over the wrapper-built table."""
        s = int(self.base.s)
        u = self.base.np_random.random()
        if self.name == "FrozenLakeEnv":
            if bytes(self.base.desc.ravel()[s]) in b"GH":   # toy_text.py:435-436 absorbing row
                self.base.lastaction = action
                return s, 0, True, {"prob": 1.0}
            outs = [action, (action + 1) % 4, (action - 1) % 4]
            fn = self._frozenlake_outcome
        else:
            outs = [action, (action + 1) % 4, (action - 1) % 4, (action + 2) % 4]
            fn = self._cliff_outcome
        cs = np.cumsum(np.asarray(self.table_prob))
        i = int(np.argmax(cs > u))                   # first index with cumsum > u, else 0
        ns, r, term = fn(s, outs[i])
        self.base.s = ns
        self.base.lastaction = action
        return int(ns), r, term, {"prob": self.table_prob[i]}

    # ---- step ---------------------------------------------------------------------------
    def _base_step(self, action):
        if self.name in ("FrozenLakeEnv", "CliffWalkingEnv"):
            # go through the TimeLimit / OrderEnforcing chain with the table-driven step
            tl = self.env
            obs, r, term, info = self._grid_step(action)
            trunc = False
            if isinstance(tl, G.TimeLimit):
                tl._elapsed_steps += 1
                trunc = tl._elapsed_steps >= tl._max_episode_steps
            return obs, r, term, trunc, info
        return self.env.step(action)

    def step(self, action):
        frozen_sim = self.is_sim_env and not self.in_sim_change
        env_change = {k: 0 for k in self.keys}
        delta_change = {k: 0.0 for k in self.keys}
        if not frozen_sim:
            if self.name in _CLASSIC:                # classic_control.py:77-94
                new_vals = {}
                for k in self.keys:
                    cur = getattr(self.base, k)
                    new, flag, delta = call_update(self.desc[k], self.states[k], cur, self.t,
                                                   self.streams, self.slot[k])
                    new_vals[k], env_change[k], delta_change[k] = new, flag, delta
                for k, bad in constraint_violations(self.name, self.base, new_vals).items():
                    if not bad:
                        setattr(self.base, k, new_vals[k])
                    else:
                        warnings.warn(f"{k} not updated", ConstraintViolationWarning)
                        delta_change[k] = 0.0
                        env_change[k] = 0
                resolve_dependencies(self.name, self.base)
            elif self.name == "BridgePort":          # toy_text.py:605-631
                order = ("P_left", "P_right") if self.split else ("P",)
                for k in order:
                    if k not in self.desc:
                        continue
                    cur = list(getattr(self.base, k))
                    new, flag, delta = call_update(self.desc[k], self.states[k], cur, self.t,
                                                   self.streams, self.slot[k])
                    setattr(self.base, k, list(new))
                    env_change[k], delta_change[k] = flag, delta
            else:                                    # toy_text.py:178-187, 362-373
                new, flag, delta = call_update(self.desc["P"], self.states["P"],
                                               self.transition_prob, self.t, self.streams, 0)
                self.transition_prob = new
                if flag:
                    self.table_prob = list(new)
                env_change["P"], delta_change["P"] = flag, delta
        state, reward, terminated, truncated, info = self._base_step(action)
        if not frozen_sim and self.name in _CLASSIC:
            info["prob"] = 1.0                       # classic_control.py:98
        if self.name in ("FrozenLakeEnv", "CliffWalkingEnv"):
            info["transition_prob"] = self.transition_prob    # toy_text.py:192,379
        return self._package(state, reward, terminated, truncated, info, env_change, delta_change)

    def _package(self, state, reward, terminated, truncated, info, env_change, delta_change):
        """NSWrapper.step (base.py:296-363)."""
        self.t += 1
        zero_c = {k: 0 for k in self.keys}
        zero_d = {k: 0.0 for k in self.keys}
        gt_c = {k: int(v) for k, v in env_change.items()}
        gt_d = {k: float(v) for k, v in delta_change.items()}
        muted = self.frozen or (self.is_sim_env and not self.in_sim_change)
        out_c = zero_c if (not self.change_notification or muted) else {**zero_c, **gt_c}
        out_d = zero_d if (not self.delta_change_notification or muted) else {**zero_d, **gt_d}
        obs = {"state": state, "env_change": out_c, "delta_change": out_d, "relative_time": self.t}
        if not self.scalar_reward:
            reward = {"reward": reward, "env_change": out_c, "delta_change": out_d,
                      "relative_time": self.t}
        info["Ground Truth Env Change"] = gt_c
        info["Ground Truth Delta Change"] = gt_d
        return obs, reward, terminated, truncated, info

    # ---- reset --------------------------------------------------------------------------
    def reset(self, *, seed=None, options=None):
        state, info = self.env.reset(seed=seed, options=options)   # base.py:377
        self.has_reset = True
        self.t = 0
        if not self.persistent_params:               # base.py:381-391
            old = self.states
            self.states = {k: SlotState(d) for k, d in self.desc.items()}
            if self.name == "BridgePort":
                # toy_text.py:665 re-clones from its own template AFTER the base class seeded /
                # transplanted, so update-fn RNGs rewind as well
                pass
            elif seed is not None:
                self._seed_update_fns(seed)
            else:                                    # base.py:423-431
                for k in self.keys:
                    if old[k].fn_rng is not None:
                        self.states[k].fn_rng = old[k].fn_rng
        elif seed is not None:                       # base.py:392-395
            self._seed_update_fns(seed)
        if not self.persistent_params:
            if self.name in _CLASSIC:                # classic_control.py:105-107
                for k, v in self.initial.items():
                    setattr(self.base, k, copy.deepcopy(v))
            elif self.name == "BridgePort":          # toy_text.py:657-664
                if self.split:
                    self.base.P_left = list(self.init_left)
                    self.base.P_right = list(self.init_right)
                else:
                    self.base.P = list(self.init_uniform)
            else:                                    # toy_text.py:206-209, 395-398
                self.transition_prob = copy.deepcopy(self.initial_prob_dist)
        zeros = {k: 0 for k in self.keys}
        obs = {"state": state, "env_change": dict(zeros), "delta_change": dict(zeros),
               "relative_time": self.t}
        info["Ground Truth Env Change"] = dict(zeros)
        info["Ground Truth Delta Change"] = dict(zeros)
        return obs, info

    def _seed_update_fns(self, seed):                # base.py:412-421
        children = np.random.SeedSequence(seed).spawn(len(self.keys))
        for child, k in zip(children, self.keys):
            if self.states[k].fn_rng is not None:
                self.states[k].fn_rng = np.random.default_rng(seed=child)

    # ---- transition tables ---------------------------------------------------------------
    def transition_table(self):
        """The table the base env samples from right now, as arrays ``prob / next / reward / done``
        of shape [S, A, D] (outcome k of action a in cell s; unused outcomes of single-outcome rows
        have prob 0): FrozenLake ``toy_text.py:426-447`` (absorbing G / H rows ``(1.0, s, 0, True)``),
        CliffWalking ``toy_text.py:86-138`` (no absorbing rows), Bridge ``envs/Bridge.py:189-221``
        (``transition_matrix``: H / G rows absorbing with the cell's own reward, per-side slip
        distributions in split mode)."""
        assert self.name in ("FrozenLakeEnv", "CliffWalkingEnv", "BridgePort")
        if self.name == "BridgePort":
            nrow, ncol, D = self.base.nrow, self.base.ncol, 3
        elif self.name == "FrozenLakeEnv":
            nrow, ncol, D = self.base.nrow, self.base.ncol, 3
        else:
            nrow, ncol, D = 4, 12, 4
        S = nrow * ncol
        prob = np.zeros((S, 4, D))
        nxt = np.zeros((S, 4, D), dtype=np.int64)
        rew = np.zeros((S, 4, D))
        done = np.zeros((S, 4, D), dtype=bool)
        for s in range(S):
            row, col = divmod(s, ncol)
            for a in range(4):
                if self.name == "FrozenLakeEnv":
                    if bytes(self.base.desc[row, col]) in b"GH":
                        prob[s, a, 0], nxt[s, a, 0], rew[s, a, 0], done[s, a, 0] = 1.0, s, 0, True
                        continue
                    for k, b in enumerate([a, (a + 1) % 4, (a - 1) % 4]):
                        ns, r, term = self._frozenlake_outcome(s, b)
                        prob[s, a, k], nxt[s, a, k], rew[s, a, k], done[s, a, k] = self.table_prob[k], ns, r, term
                elif self.name == "CliffWalkingEnv":
                    for k, b in enumerate([a, (a + 1) % 4, (a - 1) % 4, (a + 2) % 4]):
                        ns, r, term = self._cliff_outcome(s, b)
                        prob[s, a, k], nxt[s, a, k], rew[s, a, k], done[s, a, k] = self.table_prob[k], ns, r, term
                else:
                    def cell_reward(r_, c_):             # Bridge.py:159-174
                        cell = bytes(self.base.map[r_, c_])
                        return (-1, True) if cell == b"H" else (1, True) if cell == b"G" else (0, False)

                    if bytes(self.base.map[row, col]) in (b"H", b"G"):
                        r, d = cell_reward(row, col)
                        prob[s, a, 0], nxt[s, a, 0], rew[s, a, 0], done[s, a, 0] = 1.0, s, r, d
                        continue
                    if self.split:
                        cell_p = self.base.P_left if col < ncol // 2 else self.base.P_right
                    else:
                        cell_p = self.base.P
                    for k, b in enumerate([a, (a + 1) % 4, (a - 1) % 4]):
                        nr, nc = row + _BRIDGE_DELTA[b][0], col + _BRIDGE_DELTA[b][1]
                        if not (0 <= nr < nrow and 0 <= nc < ncol):
                            nr, nc = row, col
                        r, d = cell_reward(nr, nc)
                        prob[s, a, k], nxt[s, a, k], rew[s, a, k], done[s, a, k] = cell_p[k], nr * ncol + nc, r, d
        return {"prob": prob, "next": nxt, "reward": rew, "done": done}

    # ---- planning copies -----------------------------------------------------------------
    _PLAN_LIMIT = {"FrozenLakeEnv": 100, "CliffWalkingEnv": 1000, "BridgePort": 1000}

    def _deepcopy(self):
        """``__deepcopy__`` of the reference wrappers (classic_control.py:137-186, toy_text.py:
        220-251, 483-511, 684-711): a NEW gym.make chain (fresh TimeLimit count, registered limit /
        100 / 1000), reset, then state, t, current parameters and update-function state copied in."""
        c = dict(self._ctor)
        c["max_episode_steps"] = self._PLAN_LIMIT.get(self.name)       # None: registered limit
        sim = NSEnvPort(**c)
        sim.env.reset()
        sim.has_reset = True
        sim.states = copy.deepcopy(self.states)          # cursors, Memoryless times (rngs: see harness)
        sim.t = copy.deepcopy(self.t)
        if self.name in _CLASSIC:
            sim.base.state = copy.deepcopy(self.base.state)
            for k in self.keys:
                setattr(sim.base, k, copy.deepcopy(getattr(self.base, k)))
            resolve_dependencies(self.name, sim.base)
        elif self.name == "BridgePort":
            sim.base.s = copy.deepcopy(self.base.s)
            if self.split:
                sim.base.P_left, sim.base.P_right = list(self.base.P_left), list(self.base.P_right)
            else:
                sim.base.P = list(self.base.P)
        else:
            sim.base.s = copy.deepcopy(self.base.s)
            sim.transition_prob = copy.deepcopy(self.transition_prob)
            sim.table_prob = list(self.table_prob)
        sim.is_sim_env = True
        return sim

    def get_planning_env(self):
        """classic_control.py:120-135, toy_text.py:212-218, 471-481, 669-682."""
        assert self.has_reset, "The environment must be reset before getting the planning environment."
        plan = self._deepcopy()
        if not (self.is_sim_env or self.delta_change_notification):
            if self.name in _CLASSIC:
                # NOTE no _dependency_resolver here (classic_control.py:131-135): CartPole's total_mass /
                # polemass_length keep the values derived from the TRUE parameters until a step of a
                # copy with in_sim_change True recomputes them
                for k, v in self.initial.items():
                    setattr(plan.base, k, copy.deepcopy(v))
            elif self.name == "BridgePort":
                if self.split:
                    plan.base.P_left, plan.base.P_right = list(self.init_left), list(self.init_right)
                else:
                    plan.base.P = list(self.init_uniform)
            else:
                plan.transition_prob = copy.deepcopy(self.initial_prob_dist)
                plan.table_prob = list(self.initial_prob_dist)
        return plan

    # ---- introspection used by the tests ---------------------------------------------------
    def theta(self) -> dict:
        if self.name in _CLASSIC:
            return {k: getattr(self.base, k) for k in self.keys}
        if self.name == "BridgePort":
            return {k: list(getattr(self.base, k)) for k in self.keys}
        return {"P": list(self.transition_prob)}

    def freeze(self, mode: bool = True):            # base.py:443-451
        if not isinstance(mode, bool):
            raise TypeError(f"Expected mode to be a boolean, got {type(mode)}")
        self.frozen = mode
        return self

    def unfreeze(self):
        return self.freeze(False)
