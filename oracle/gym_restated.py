"""Restated gymnasium-1.2.1 base environments (TEST INFRASTRUCTURE -- see oracle/__init__.py).

PARITY STATUS: **UNPINNED** (except the CartPole step, see below).  gymnasium 1.2.1 (reference ``uv.lock:958-959``) is neither
vendored in /root/reference nor installed in this image, and no reference test asserts a
numeric post-step state.  Everything below restates the *published upstream algorithm*
 and is anchored on:

* attribute names the reference reads/writes: ``ns_gym/base.py:611-635`` (ATTRIBUTE_MAP),
  ``ns_gym/wrappers/toy_text.py:51-54,65-69,314-319``;
* CartPole equations: PINNED to one ulp against the in-tree legacy copy
  ``ns_gym/benchmark_algorithms/rats-experiments/code/envs/nscartpole_v0.py:76-135`` (loaded in
  place by ``tests/golden/make_cartpole_anchor.py``; ``tests/test_oracle_cartpole_anchor.py``
  checks this module and the CUDA kernel against its 4 096 recorded transitions);
* default parameter tables ``docs/source/env_pages/classic_control/*.md``;
* registered episode limits (``ns_gym/__init__.py:17-21`` for Bridge).

The module doubles as a minimal ``gymnasium`` import shim (``install_shim``) so that the
reference's own NS layer can be imported verbatim in this container (oracle/ref_loader.py).
"""
from __future__ import annotations

import importlib
import math
import sys
import types

import numpy as np

# --------------------------------------------------------------------------------------
# spaces
# --------------------------------------------------------------------------------------


class Space:
    def __init__(self, shape=None, dtype=None):
        self.shape = shape
        self.dtype = dtype
        self._rng = None

    @property
    def np_random(self):
        if self._rng is None:
            self._rng = np.random.default_rng()
        return self._rng

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)


class Discrete(Space):
    def __init__(self, n, start=0):
        super().__init__((), np.int64)
        self.n = int(n)
        self.start = int(start)

    def sample(self):
        return int(self.start + self.np_random.integers(self.n))

    def contains(self, x):
        try:
            xi = int(x)
        except (TypeError, ValueError):
            return False
        return xi == x and self.start <= xi < self.start + self.n

    def __eq__(self, other):
        return isinstance(other, Discrete) and other.n == self.n and other.start == self.start


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.shape(low)
        super().__init__(tuple(shape), dtype)
        self.low = np.broadcast_to(np.asarray(low, dtype=np.float64), self.shape).astype(dtype)
        self.high = np.broadcast_to(np.asarray(high, dtype=np.float64), self.shape).astype(dtype)

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return self.np_random.uniform(lo, hi, size=self.shape).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))


class Dict(Space):
    def __init__(self, spaces=None, **kw):
        super().__init__(None, None)
        self.spaces = dict(spaces or {}, **kw)

    def __getitem__(self, k):
        return self.spaces[k]

    def keys(self):
        return self.spaces.keys()

    def sample(self):
        return {k: s.sample() for k, s in self.spaces.items()}


# --------------------------------------------------------------------------------------
# Env / Wrapper / TimeLimit / registry
# --------------------------------------------------------------------------------------


class EnvSpec:
    def __init__(self, id, entry_point, max_episode_steps=None, kwargs=None):
        self.id = id
        self.entry_point = entry_point
        self.max_episode_steps = max_episode_steps
        self.kwargs = dict(kwargs or {})


class Env:
    metadata: dict = {}
    spec: EnvSpec | None = None
    render_mode = None
    _np_random = None

    @property
    def unwrapped(self):
        return self

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random = np.random.default_rng()
        return self._np_random

    @np_random.setter
    def np_random(self, v):
        self._np_random = v

    def reset(self, *, seed=None, options=None):
        # gymnasium.Env.reset: re-create the PCG64 generator only for an explicit seed
        if seed is not None:
            self._np_random = np.random.default_rng(seed)

    def step(self, action):
        raise NotImplementedError

    def render(self):
        return None

    def close(self):
        pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self._action_space = None
        self._observation_space = None

    @property
    def unwrapped(self):
        return self.env.unwrapped

    @property
    def spec(self):
        return self.env.spec

    @property
    def action_space(self):
        return self._action_space if self._action_space is not None else self.env.action_space

    @action_space.setter
    def action_space(self, v):
        self._action_space = v

    @property
    def observation_space(self):
        return (self._observation_space if self._observation_space is not None
                else self.env.observation_space)

    @observation_space.setter
    def observation_space(self, v):
        self._observation_space = v

    @property
    def np_random(self):
        return self.env.np_random

    @np_random.setter
    def np_random(self, v):
        self.env.np_random = v

    @property
    def render_mode(self):
        return self.env.render_mode

    def step(self, action):
        return self.env.step(action)

    def reset(self, *, seed=None, options=None):
        return self.env.reset(seed=seed, options=options)

    def render(self):
        return self.env.render()

    def close(self):
        return self.env.close()

    def get_wrapper_attr(self, name):
        if name in self.__dict__ or hasattr(type(self), name):
            return getattr(self, name)
        if isinstance(self.env, Wrapper):
            return self.env.get_wrapper_attr(name)
        return getattr(self.env, name)

    def __repr__(self):
        return f"<{type(self).__name__}{self.env!r}>"

    __str__ = __repr__


class TimeLimit(Wrapper):
    """gymnasium.wrappers.TimeLimit: truncated |= elapsed >= max."""

    def __init__(self, env, max_episode_steps):
        super().__init__(env)
        self._max_episode_steps = max_episode_steps
        self._elapsed_steps = None

    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        self._elapsed_steps += 1
        if self._elapsed_steps >= self._max_episode_steps:
            truncated = True
        return obs, reward, terminated, truncated, info

    def reset(self, *, seed=None, options=None):
        self._elapsed_steps = 0
        return self.env.reset(seed=seed, options=options)


class OrderEnforcing(Wrapper):
    def __init__(self, env):
        super().__init__(env)
        self._has_reset = False

    def step(self, action):
        if not self._has_reset:
            raise RuntimeError("Cannot call env.step() before calling env.reset()")
        return self.env.step(action)

    def reset(self, *, seed=None, options=None):
        self._has_reset = True
        return self.env.reset(seed=seed, options=options)


registry: dict[str, EnvSpec] = {}


def register(id, entry_point=None, max_episode_steps=None, kwargs=None, **_ignored):
    registry[id] = EnvSpec(id, entry_point, max_episode_steps, kwargs)


def _load_entry_point(ep):
    if callable(ep):
        return ep
    mod, _, attr = ep.partition(":")
    return getattr(importlib.import_module(mod), attr)


class _OpaqueEnv(Env):
    """Stand-in for ids whose engine is absent (MuJoCo).  base._generate_tunable_params
    (base.py:659-682) silently skips class names it does not know."""

    def __init__(self, **kw):
        self.observation_space = Box(-np.inf, np.inf, (1,), np.float64)
        self.action_space = Box(-1.0, 1.0, (1,), np.float32)

    def reset(self, *, seed=None, options=None):
        super().reset(seed=seed)
        return np.zeros(1), {}

    def step(self, action):
        return np.zeros(1), 0.0, False, False, {}


def make(id, max_episode_steps=None, **kwargs):
    if isinstance(id, EnvSpec):
        spec = id
    elif id in registry:
        spec = registry[id]
    elif id.endswith("-v5"):
        spec = EnvSpec(id, _OpaqueEnv, 1000)
    else:
        raise KeyError(f"gym_restated: unknown environment id {id!r}")
    kw = dict(spec.kwargs)
    kw.update(kwargs)
    env = _load_entry_point(spec.entry_point)(**kw)
    env.spec = EnvSpec(spec.id, spec.entry_point, spec.max_episode_steps, kw)
    limit = max_episode_steps if max_episode_steps is not None else spec.max_episode_steps
    env = OrderEnforcing(env)
    if limit is not None and limit > 0:
        env = TimeLimit(env, limit)
    return env


# --------------------------------------------------------------------------------------
# classic control
# --------------------------------------------------------------------------------------


class CartPoleEnv(Env):
    """A.1.  Euler integrator, float64 state, float32 observation."""

    def __init__(self, sutton_barto_reward=False, render_mode=None):
        self.gravity = 9.8
        self.masscart = 1.0
        self.masspole = 0.1
        self.total_mass = self.masspole + self.masscart
        self.length = 0.5  # half the pole's length
        self.polemass_length = self.masspole * self.length
        self.force_mag = 10.0
        self.tau = 0.02
        self.kinematics_integrator = "euler"
        self.theta_threshold_radians = 12 * 2 * math.pi / 360
        self.x_threshold = 2.4
        self._sutton_barto_reward = sutton_barto_reward
        high = np.array([self.x_threshold * 2, np.inf, self.theta_threshold_radians * 2, np.inf])
        self.action_space = Discrete(2)
        self.observation_space = Box(-high, high, dtype=np.float32)
        self.render_mode = render_mode
        self.state = None
        self.steps_beyond_terminated = None

    def step(self, action):
        assert self.action_space.contains(action), f"{action!r} invalid"
        assert self.state is not None, "Call reset before using step method."
        x, x_dot, theta, theta_dot = self.state
        force = self.force_mag if action == 1 else -self.force_mag
        costheta = np.cos(theta)
        sintheta = np.sin(theta)
        temp = (force + self.polemass_length * np.square(theta_dot) * sintheta) / self.total_mass
        thetaacc = (self.gravity * sintheta - costheta * temp) / (
            self.length * (4.0 / 3.0 - self.masspole * np.square(costheta) / self.total_mass)
        )
        xacc = temp - self.polemass_length * thetaacc * costheta / self.total_mass
        x = x + self.tau * x_dot
        x_dot = x_dot + self.tau * xacc
        theta = theta + self.tau * theta_dot
        theta_dot = theta_dot + self.tau * thetaacc
        self.state = np.array((x, x_dot, theta, theta_dot), dtype=np.float64)
        terminated = bool(
            x < -self.x_threshold
            or x > self.x_threshold
            or theta < -self.theta_threshold_radians
            or theta > self.theta_threshold_radians
        )
        if not terminated:
            reward = 0.0 if self._sutton_barto_reward else 1.0
        elif self.steps_beyond_terminated is None:
            self.steps_beyond_terminated = 0
            reward = -1.0 if self._sutton_barto_reward else 1.0
        else:
            self.steps_beyond_terminated += 1
            reward = -1.0 if self._sutton_barto_reward else 0.0
        return np.array(self.state, dtype=np.float32), reward, terminated, False, {}

    def reset(self, *, seed=None, options=None):
        super().reset(seed=seed)
        low, high = -0.05, 0.05
        self.state = self.np_random.uniform(low=low, high=high, size=(4,))
        self.steps_beyond_terminated = None
        return np.array(self.state, dtype=np.float32), {}


def _wrap(x, m, M):
    diff = M - m
    while x > M:
        x = x - diff
    while x < m:
        x = x + diff
    return x


def _bound(x, m, M):
    return min(max(x, m), M)


def _rk4(derivs, y0, t):
    yout = np.zeros((len(t), len(y0)), np.float64)
    yout[0] = y0
    for i in np.arange(len(t) - 1):
        this = t[i]
        dt = t[i + 1] - this
        dt2 = dt / 2.0
        y0 = yout[i]
        k1 = np.asarray(derivs(y0))
        k2 = np.asarray(derivs(y0 + dt2 * k1))
        k3 = np.asarray(derivs(y0 + dt2 * k2))
        k4 = np.asarray(derivs(y0 + dt * k3))
        yout[i + 1] = y0 + dt / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
    return yout[-1][:4]


class AcrobotEnv(Env):
    """A.2.  RK4 over one interval [0, dt], "book" dynamics, g = 9.8 hard-coded."""

    dt = 0.2
    LINK_LENGTH_1 = 1.0
    LINK_LENGTH_2 = 1.0
    LINK_MASS_1 = 1.0
    LINK_MASS_2 = 1.0
    LINK_COM_POS_1 = 0.5
    LINK_COM_POS_2 = 0.5
    LINK_MOI = 1.0
    MAX_VEL_1 = 4 * math.pi
    MAX_VEL_2 = 9 * math.pi
    AVAIL_TORQUE = [-1.0, 0.0, +1]
    torque_noise_max = 0.0
    book_or_nips = "book"

    def __init__(self, render_mode=None):
        self.render_mode = render_mode
        high = np.array([1.0, 1.0, 1.0, 1.0, self.MAX_VEL_1, self.MAX_VEL_2], dtype=np.float32)
        self.observation_space = Box(-high, high, dtype=np.float32)
        self.action_space = Discrete(3)
        self.state = None

    def reset(self, *, seed=None, options=None):
        super().reset(seed=seed)
        low, high = -0.1, 0.1
        self.state = self.np_random.uniform(low=low, high=high, size=(4,)).astype(np.float32)
        return self._get_ob(), {}

    def step(self, a):
        s = self.state
        assert s is not None, "Call reset before using AcrobotEnv object."
        torque = self.AVAIL_TORQUE[a]
        s_augmented = np.append(s, torque)
        ns = _rk4(self._dsdt, s_augmented, [0, self.dt])
        ns[0] = _wrap(ns[0], -math.pi, math.pi)
        ns[1] = _wrap(ns[1], -math.pi, math.pi)
        ns[2] = _bound(ns[2], -self.MAX_VEL_1, self.MAX_VEL_1)
        ns[3] = _bound(ns[3], -self.MAX_VEL_2, self.MAX_VEL_2)
        self.state = ns
        terminated = self._terminal()
        reward = -1.0 if not terminated else 0.0
        return self._get_ob(), reward, terminated, False, {}

    def _get_ob(self):
        s = self.state
        return np.array(
            [np.cos(s[0]), np.sin(s[0]), np.cos(s[1]), np.sin(s[1]), s[2], s[3]], dtype=np.float32
        )

    def _terminal(self):
        s = self.state
        return bool(-np.cos(s[0]) - np.cos(s[1] + s[0]) > 1.0)

    def _dsdt(self, s_augmented):
        m1 = self.LINK_MASS_1
        m2 = self.LINK_MASS_2
        l1 = self.LINK_LENGTH_1
        lc1 = self.LINK_COM_POS_1
        lc2 = self.LINK_COM_POS_2
        I1 = self.LINK_MOI
        I2 = self.LINK_MOI
        g = 9.8
        a = s_augmented[-1]
        s = s_augmented[:-1]
        theta1, theta2, dtheta1, dtheta2 = s
        cos, sin, pi = np.cos, np.sin, np.pi
        d1 = m1 * lc1**2 + m2 * (l1**2 + lc2**2 + 2 * l1 * lc2 * cos(theta2)) + I1 + I2
        d2 = m2 * (lc2**2 + l1 * lc2 * cos(theta2)) + I2
        phi2 = m2 * lc2 * g * cos(theta1 + theta2 - pi / 2.0)
        phi1 = (
            -m2 * l1 * lc2 * dtheta2**2 * sin(theta2)
            - 2 * m2 * l1 * lc2 * dtheta2 * dtheta1 * sin(theta2)
            + (m1 * lc1 + m2 * l1) * g * cos(theta1 - pi / 2)
            + phi2
        )
        if self.book_or_nips == "nips":
            ddtheta2 = (a + d2 / d1 * phi1 - phi2) / (m2 * lc2**2 + I2 - d2**2 / d1)
        else:
            ddtheta2 = (
                a + d2 / d1 * phi1 - m2 * l1 * lc2 * dtheta1**2 * sin(theta2) - phi2
            ) / (m2 * lc2**2 + I2 - d2**2 / d1)
        ddtheta1 = -(d2 * ddtheta2 + phi1) / d1
        return dtheta1, dtheta2, ddtheta1, ddtheta2, 0.0


class MountainCarEnv(Env):
    """A.3."""

    def __init__(self, render_mode=None, goal_velocity=0):
        self.min_position = -1.2
        self.max_position = 0.6
        self.max_speed = 0.07
        self.goal_position = 0.5
        self.goal_velocity = goal_velocity
        self.force = 0.001
        self.gravity = 0.0025
        self.low = np.array([self.min_position, -self.max_speed], dtype=np.float32)
        self.high = np.array([self.max_position, self.max_speed], dtype=np.float32)
        self.render_mode = render_mode
        self.action_space = Discrete(3)
        self.observation_space = Box(self.low, self.high, dtype=np.float32)
        self.state = None

    def step(self, action):
        assert self.action_space.contains(action), f"{action!r} invalid"
        position, velocity = self.state
        velocity += (action - 1) * self.force + math.cos(3 * position) * (-self.gravity)
        velocity = np.clip(velocity, -self.max_speed, self.max_speed)
        position += velocity
        position = np.clip(position, self.min_position, self.max_position)
        if position == self.min_position and velocity < 0:
            velocity = 0
        terminated = bool(position >= self.goal_position and velocity >= self.goal_velocity)
        reward = -1.0
        self.state = (position, velocity)
        return np.array(self.state, dtype=np.float32), reward, terminated, False, {}

    def reset(self, *, seed=None, options=None):
        super().reset(seed=seed)
        low, high = -0.6, -0.4
        self.state = np.array([self.np_random.uniform(low=low, high=high), 0])
        return np.array(self.state, dtype=np.float32), {}


class Continuous_MountainCarEnv(Env):
    """A.4.  State is stored as float32 after every step; hill term uses a literal 0.0025."""

    def __init__(self, render_mode=None, goal_velocity=0):
        self.min_action = -1.0
        self.max_action = 1.0
        self.min_position = -1.2
        self.max_position = 0.6
        self.max_speed = 0.07
        self.goal_position = 0.45
        self.goal_velocity = goal_velocity
        self.power = 0.0015
        self.low_state = np.array([self.min_position, -self.max_speed], dtype=np.float32)
        self.high_state = np.array([self.max_position, self.max_speed], dtype=np.float32)
        self.render_mode = render_mode
        self.action_space = Box(self.min_action, self.max_action, shape=(1,), dtype=np.float32)
        self.observation_space = Box(self.low_state, self.high_state, dtype=np.float32)
        self.state = None

    def step(self, action):
        position = self.state[0]
        velocity = self.state[1]
        force = min(max(action[0], self.min_action), self.max_action)
        velocity += force * self.power - 0.0025 * math.cos(3 * position)
        if velocity > self.max_speed:
            velocity = self.max_speed
        if velocity < -self.max_speed:
            velocity = -self.max_speed
        position += velocity
        if position > self.max_position:
            position = self.max_position
        if position < self.min_position:
            position = self.min_position
        if position == self.min_position and velocity < 0:
            velocity = 0
        terminated = bool(position >= self.goal_position and velocity >= self.goal_velocity)
        reward = 0
        if terminated:
            reward = 100.0
        reward -= math.pow(action[0], 2) * 0.1
        self.state = np.array([position, velocity], dtype=np.float32)
        return self.state, reward, terminated, False, {}

    def reset(self, *, seed=None, options=None):
        super().reset(seed=seed)
        low, high = -0.6, -0.4
        self.state = np.array([self.np_random.uniform(low=low, high=high), 0])
        return np.array(self.state, dtype=np.float32), {}


def _angle_normalize(x):
    return ((x + np.pi) % (2 * np.pi)) - np.pi


class PendulumEnv(Env):
    """A.5.  Pendulum-v1: g = 10.0; cost uses the pre-step state; never terminates."""

    def __init__(self, render_mode=None, g=10.0):
        self.max_speed = 8
        self.max_torque = 2.0
        self.dt = 0.05
        self.g = g
        self.m = 1.0
        self.l = 1.0
        self.render_mode = render_mode
        high = np.array([1.0, 1.0, self.max_speed], dtype=np.float32)
        self.action_space = Box(-self.max_torque, self.max_torque, shape=(1,), dtype=np.float32)
        self.observation_space = Box(-high, high, dtype=np.float32)
        self.state = None
        self.last_u = None

    def step(self, u):
        th, thdot = self.state
        g = self.g
        m = self.m
        l = self.l
        dt = self.dt
        u = np.clip(u, -self.max_torque, self.max_torque)[0]
        self.last_u = u
        costs = _angle_normalize(th) ** 2 + 0.1 * thdot**2 + 0.001 * (u**2)
        newthdot = thdot + (3 * g / (2 * l) * np.sin(th) + 3.0 / (m * l**2) * u) * dt
        newthdot = np.clip(newthdot, -self.max_speed, self.max_speed)
        newth = th + newthdot * dt
        self.state = np.array([newth, newthdot])
        return self._get_obs(), -costs, False, False, {}

    def reset(self, *, seed=None, options=None):
        super().reset(seed=seed)
        high = np.array([np.pi, 1.0])
        low = -high
        self.state = self.np_random.uniform(low=low, high=high)
        self.last_u = None
        return self._get_obs(), {}

    def _get_obs(self):
        theta, thetadot = self.state
        return np.array([np.cos(theta), np.sin(theta), thetadot], dtype=np.float32)


# --------------------------------------------------------------------------------------
# toy text
# --------------------------------------------------------------------------------------

FROZEN_LAKE_MAPS = {
    "4x4": ["SFFF", "FHFH", "FFFH", "HFFG"],
    "8x8": [
        "SFFFFFFF",
        "FFFFFFFF",
        "FFFHFFFF",
        "FFFFFHFF",
        "FFFHFFFF",
        "FHHFFFHF",
        "FHFFHFHF",
        "FFFHFFFG",
    ],
}


def categorical_sample(prob_n, np_random):
    """First index whose running sum exceeds one uniform draw; 0 when none does."""
    prob_n = np.asarray(prob_n)
    csprob_n = np.cumsum(prob_n)
    return np.argmax(csprob_n > np_random.random())


class FrozenLakeEnv(Env):
    LEFT, DOWN, RIGHT, UP = 0, 1, 2, 3

    def __init__(self, render_mode=None, desc=None, map_name="4x4", is_slippery=True,
                 success_rate=1.0 / 3.0, reward_schedule=(1, 0, 0)):
        if desc is None:
            desc = FROZEN_LAKE_MAPS[map_name]
        self.desc = desc = np.asarray(desc, dtype="c")
        self.nrow, self.ncol = nrow, ncol = desc.shape
        self.reward_range = (0, 1)
        nA = 4
        nS = nrow * ncol
        self.initial_state_distrib = np.array(desc == b"S").astype("float64").ravel()
        self.initial_state_distrib /= self.initial_state_distrib.sum()
        fail_rate = (1.0 - success_rate) / 2.0
        self.P = {s: {a: [] for a in range(nA)} for s in range(nS)}

        def to_s(row, col):
            return row * ncol + col

        def inc(row, col, a):
            if a == 0:
                col = max(col - 1, 0)
            elif a == 1:
                row = min(row + 1, nrow - 1)
            elif a == 2:
                col = min(col + 1, ncol - 1)
            elif a == 3:
                row = max(row - 1, 0)
            return row, col

        def outcome(row, col, action):
            nr, nc = inc(row, col, action)
            letter = desc[nr, nc]
            return to_s(nr, nc), float(letter == b"G"), bytes(letter) in b"GH"

        for row in range(nrow):
            for col in range(ncol):
                s = to_s(row, col)
                for a in range(4):
                    li = self.P[s][a]
                    if desc[row, col] in b"GH":
                        li.append((1.0, s, 0, True))
                    elif is_slippery:
                        for b in [(a - 1) % 4, a, (a + 1) % 4]:
                            p = success_rate if b == a else fail_rate
                            li.append((p, *outcome(row, col, b)))
                    else:
                        li.append((1.0, *outcome(row, col, a)))

        self.observation_space = Discrete(nS)
        self.action_space = Discrete(nA)
        self.render_mode = render_mode
        self.s = None
        self.lastaction = None

    def step(self, a):
        transitions = self.P[self.s][a]
        i = categorical_sample([t[0] for t in transitions], self.np_random)
        p, s, r, t = transitions[i]
        self.s = s
        self.lastaction = a
        return int(s), r, t, False, {"prob": p}

    def reset(self, *, seed=None, options=None):
        super().reset(seed=seed)
        self.s = categorical_sample(self.initial_state_distrib, self.np_random)
        self.lastaction = None
        return int(self.s), {"prob": 1}


class CliffWalkingEnv(Env):
    """4x12 grid, UP 0 / RIGHT 1 / DOWN 2 / LEFT 3, start 36, deterministic base table."""

    def __init__(self, render_mode=None, is_slippery=False):
        self.shape = (4, 12)
        self.start_state_index = np.ravel_multi_index((3, 0), self.shape)
        self.nS = int(np.prod(self.shape))
        self.nA = 4
        self.is_slippery = is_slippery
        self._cliff = np.zeros(self.shape, dtype=bool)
        self._cliff[3, 1:-1] = True
        deltas = {0: (-1, 0), 1: (0, 1), 2: (1, 0), 3: (0, -1)}
        self.P = {}
        for s in range(self.nS):
            pos = np.unravel_index(s, self.shape)
            self.P[s] = {}
            for a in range(self.nA):
                nr = min(max(pos[0] + deltas[a][0], 0), self.shape[0] - 1)
                nc = min(max(pos[1] + deltas[a][1], 0), self.shape[1] - 1)
                ns = int(np.ravel_multi_index((nr, nc), self.shape))
                if self._cliff[nr, nc]:
                    self.P[s][a] = [(1.0, int(self.start_state_index), -100, False)]
                else:
                    done = (nr, nc) == (self.shape[0] - 1, self.shape[1] - 1)
                    self.P[s][a] = [(1.0, ns, -1, done)]
        self.initial_state_distrib = np.zeros(self.nS)
        self.initial_state_distrib[self.start_state_index] = 1.0
        self.observation_space = Discrete(self.nS)
        self.action_space = Discrete(self.nA)
        self.render_mode = render_mode
        self.s = None
        self.lastaction = None

    def step(self, a):
        transitions = self.P[self.s][a]
        i = categorical_sample([t[0] for t in transitions], self.np_random)
        p, s, r, t = transitions[i]
        self.s = s
        self.lastaction = a
        return int(s), r, t, False, {"prob": p}

    def reset(self, *, seed=None, options=None):
        super().reset(seed=seed)
        self.s = categorical_sample(self.initial_state_distrib, self.np_random)
        self.lastaction = None
        return int(self.s), {"prob": 1}


# registered ids and episode limits
register("CartPole-v1", CartPoleEnv, 500)
register("CartPole-v0", CartPoleEnv, 200)
register("Acrobot-v1", AcrobotEnv, 500)
register("MountainCar-v0", MountainCarEnv, 200)
register("MountainCarContinuous-v0", Continuous_MountainCarEnv, 999)
register("Pendulum-v1", PendulumEnv, 200)
register("FrozenLake-v1", FrozenLakeEnv, 100, kwargs={"map_name": "4x4"})
register("FrozenLake8x8-v1", FrozenLakeEnv, 200, kwargs={"map_name": "8x8"})
register("CliffWalking-v1", CliffWalkingEnv, None)


# --------------------------------------------------------------------------------------
# import shim
# --------------------------------------------------------------------------------------


def install_shim() -> bool:
    """Expose this module as ``gymnasium`` (+ a bare ``mujoco``) unless the real packages
    are importable.  Returns True when the shim (not the real gymnasium) is active."""
    if "gymnasium" in sys.modules:
        return getattr(sys.modules["gymnasium"], "__nsgym_shim__", False)
    try:
        importlib.import_module("gymnasium")
        return False
    except ModuleNotFoundError:
        pass
    me = sys.modules[__name__]
    g = types.ModuleType("gymnasium")
    g.__nsgym_shim__ = True
    g.__version__ = "1.2.1+restated"
    g.__path__ = []
    for name in ("Env", "Wrapper", "make", "register", "registry", "Space"):
        setattr(g, name, getattr(me, name))
    spaces = types.ModuleType("gymnasium.spaces")
    for name in ("Space", "Discrete", "Box", "Dict"):
        setattr(spaces, name, getattr(me, name))
    envs = types.ModuleType("gymnasium.envs")
    envs.__path__ = []
    reg = types.ModuleType("gymnasium.envs.registration")
    for name in ("register", "registry", "make", "EnvSpec"):
        setattr(reg, name, getattr(me, name))
    wrappers = types.ModuleType("gymnasium.wrappers")
    wrappers.TimeLimit = TimeLimit
    wrappers.OrderEnforcing = OrderEnforcing
    g.spaces, g.envs, g.wrappers = spaces, envs, wrappers
    envs.registration = reg
    sys.modules.update({
        "gymnasium": g,
        "gymnasium.spaces": spaces,
        "gymnasium.envs": envs,
        "gymnasium.envs.registration": reg,
        "gymnasium.wrappers": wrappers,
    })
    if "mujoco" not in sys.modules:
        try:
            importlib.import_module("mujoco")
        except ModuleNotFoundError:
            sys.modules["mujoco"] = types.ModuleType("mujoco")
    return True
