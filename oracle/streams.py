"""Pre-drawn random-stream injection (TEST INFRASTRUCTURE -- see oracle/__init__.py).

The reference draws randomness in four places on the step path:

* update functions: ``self.rng.normal(mu, sigma)`` (``update_functions/single_param.py:79,111,149,345,446``)
* stochastic schedulers: ``self.rng.random()`` / ``self.rng.geometric(p, size=(1,))``
  (``schedulers.py:28,112-113,177``)
* gymnasium toy-text envs: ``self.np_random.random()`` inside ``categorical_sample``;
  classic-control resets: ``self.np_random.uniform(low, high, size)``
* Bridge: ``np.random.choice(..., p=P)`` on the process-global legacy RNG
  (``envs/Bridge.py:95-97``)

Injection is **positional**: one table per env, indexed by (global step index k, lane).
A lane is addressed only if the implementation actually draws at that step, so
conditional draws (scheduler did not fire, sigma == 0, ...) simply leave the entry unused.
The CUDA kernel uses exactly the same addressing (``include/nsgym_b200.h``):

    uniforms[k][lane][env]   lane 0          gridworld draw (slip; start-cell draw on reset)
                             lane 1..4       classic-control reset draws (initial state)
                             lane 5+j        scheduler draw of parameter slot j
                             lane 5+P+(j*16+r)*4+i  i-th uniform of the r-th Dirichlet draw (r mod 16) of
                                                   parameter slot j at this step (RandomCategorical; the
                                                   Lipschitz-bounded wrapper redraws until accepted)
    normals [k][j][env]      update-function standard normal of parameter slot j

Dirichlet(1, .., 1) convention under injection (numpy draws standard gammas and scales by the
reciprocal of their sum): e_i = -log1p(-u_i), p_i = e_i * (1 / (e_0 + .. + e_{n-1})).

``Generator.normal(mu, sigma)`` equals ``mu + sigma * standard_normal()`` bit for bit, so
feeding standard normals reproduces the reference arithmetic exactly.
"""
from __future__ import annotations

import numpy as np

LANE_DYN = 0
LANE_RESET0 = 1
N_RESET_LANES = 4
LANE_SCHED0 = LANE_RESET0 + N_RESET_LANES  # 5


DIR_TRIES = 16         # injected Dirichlet draws available per (step, slot); further tries wrap around
DIR_WIDTH = 4          # uniforms per draw (3 or 4 outcomes)


def n_uniform_lanes(n_slots: int) -> int:
    return LANE_SCHED0 + n_slots + n_slots * DIR_TRIES * DIR_WIDTH


def dirichlet_lane(n_slots: int, slot: int, attempt: int, i: int) -> int:
    return LANE_SCHED0 + n_slots + (slot * DIR_TRIES + attempt % DIR_TRIES) * DIR_WIDTH + i


def dirichlet_from_uniforms(us):
    """Dirichlet(1,..,1) from uniforms: standard exponentials scaled by the reciprocal of their sum."""
    e = [-np.log1p(-u) for u in us]
    acc = 0.0
    for v in e:
        acc = acc + v
    inv = 1.0 / acc
    return [float(v * inv) for v in e]


class Clock:
    """Shared global step index; deep copies alias the same clock on purpose."""

    def __init__(self):
        self.k = 0

    def __deepcopy__(self, memo):
        return self


class EnvStreams:
    """The pre-drawn tables of ONE env: uniforms[K, L] and normals[K, P]."""

    def __init__(self, uniforms: np.ndarray, normals: np.ndarray, clock: Clock):
        self.u = np.asarray(uniforms, dtype=np.float64)
        self.z = np.asarray(normals, dtype=np.float64)
        self.clock = clock
        self.n_slots = self.z.shape[1]        # lanes of the Dirichlet block start after the scheduler lanes
        self._dir_seen = {}                   # (k, slot) -> draws made so far at this step

    def __deepcopy__(self, memo):
        return self

    def uniform(self, lane: int) -> float:
        return float(self.u[self.clock.k, lane])

    def std_normal(self, slot: int) -> float:
        return float(self.z[self.clock.k, slot])

    def sched_uniform(self, slot: int, t: int) -> float:
        """Scheduler draw of parameter slot ``slot`` at episode time ``t``: positional here (the
        table row of the current step); a native-stream stand-in may key it by ``t`` instead."""
        return self.uniform(LANE_SCHED0 + slot)

    def dirichlet(self, slot: int, n: int):
        """The next Dirichlet(1,..,1) draw of parameter slot ``slot`` at the current step."""
        key = (self.clock.k, slot)
        attempt = self._dir_seen.get(key, 0)
        self._dir_seen = {key: attempt + 1} if key not in self._dir_seen else {**self._dir_seen, key: attempt + 1}
        us = [self.u[self.clock.k, dirichlet_lane(self.n_slots, slot, attempt, i)] for i in range(n)]
        return dirichlet_from_uniforms(us)


class SlotRng:
    """Drop-in for ``fn.rng`` / ``fn.scheduler.rng`` of parameter slot ``slot``."""

    def __init__(self, streams: EnvStreams, slot: int):
        self.s = streams
        self.slot = slot

    def __deepcopy__(self, memo):
        return self

    def normal(self, mu=0.0, sigma=1.0):
        return mu + sigma * self.s.std_normal(self.slot)

    def random(self):
        return self.s.sched_uniform(self.slot, -1)

    def dirichlet(self, alpha):
        return np.array(self.s.dirichlet(self.slot, len(alpha)))

    def geometric(self, p, size=None):
        # inverse-CDF geometric on {1,2,...}: ceil(log1p(-u)/log1p(-p)); shared convention
        u = self.s.uniform(LANE_SCHED0 + self.slot)
        g = geometric_from_uniform(u, p)
        return np.array([g]) if size is not None else g


def geometric_from_uniform(u: float, p: float) -> int:
    """Number of Bernoulli(p) trials up to and including the first success, by inversion.

    numpy's own ``Generator.geometric`` uses a search for p >= 1/3 and inversion with a
    different uniform for smaller p; under *injection* both implementations use this rule.
    """
    if p >= 1.0:
        return 1
    g = int(np.ceil(np.log1p(-u) / np.log1p(-p)))
    return max(g, 1)


class EnvNpRandom:
    """Drop-in for ``env.unwrapped.np_random`` of the restated gymnasium envs."""

    def __init__(self, streams: EnvStreams):
        self.s = streams

    def __deepcopy__(self, memo):
        return self

    def random(self):
        # toy-text step draw; the toy-text *reset* draw (start-cell categorical) lands on
        # the same lane -- a call is either a reset or a step, never both
        return self.s.uniform(LANE_DYN)

    def uniform(self, low=0.0, high=1.0, size=None):
        # Generator.uniform == low + (high - low) * next_double, element by element
        if size is None and np.ndim(low) == 0:
            return low + (high - low) * self.s.uniform(LANE_RESET0)
        n = int(np.prod(size)) if size is not None else int(np.size(low))
        u = np.array([self.s.uniform(LANE_RESET0 + i) for i in range(n)])
        return np.asarray(low) + (np.asarray(high) - np.asarray(low)) * u


def draw_tables(seed: int, n_envs: int, n_steps: int, n_slots: int):
    """uniforms[K, L, N] in [0,1) and normals[K, P, N], float64, for a whole batch."""
    rng = np.random.default_rng(seed)
    L = n_uniform_lanes(n_slots)
    u = rng.random((n_steps, L, n_envs))
    z = rng.standard_normal((n_steps, max(n_slots, 1), n_envs))
    return u, z
