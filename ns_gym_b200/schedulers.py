"""Scheduler descriptions (same names / constructor arguments / attributes as
``ns_gym/schedulers.py:9-198``).  Each lowers to one ``sched_op`` of the step kernel
(``ns_gym_b200/compile.py``); the fire rule is quoted from the reference line it follows.
"""
from __future__ import annotations

import numpy as np

from . import base


class ContinuousScheduler(base.Scheduler):
    """Fires at every step in range (``schedulers.py:46-53``)."""

    def __init__(self, start=0, end=np.inf) -> None:
        super().__init__(start, end)


class PeriodicScheduler(base.Scheduler):
    """Fires when ``t % period == 0`` (``schedulers.py:77-89``)."""

    def __init__(self, period: int, start=0, end=np.inf) -> None:
        super().__init__(start, end)
        self.period = period


class DiscreteScheduler(base.Scheduler):
    """Fires when ``t in event_list`` (``schedulers.py:56-74``)."""

    def __init__(self, event_list: set, start=0, end=np.inf) -> None:
        super().__init__(start, end)
        self.event_list = event_list
        assert min(event_list) >= start, "Scheduler start time occurs after first event in event list"
        assert max(event_list) <= end, "Scheduler end time occurs before last event in event list"


class BurstScheduler(base.Scheduler):
    """Fires while ``(t % (on+off)) < on`` (``schedulers.py:119-140``)."""

    def __init__(self, on_duration: int, off_duration: int, start=0, end=np.inf) -> None:
        super().__init__(start, end)
        self.on_duration = on_duration
        self.off_duration = off_duration
        self.cycle = on_duration + off_duration


class WindowScheduler(base.Scheduler):
    """Fires when ``t`` lies in any inclusive ``(start, end)`` window (``schedulers.py:180-198``)."""

    def __init__(self, windows: list, start=0, end=np.inf) -> None:
        super().__init__(start, end)
        self.windows = windows


class RandomScheduler(base.Scheduler):
    """Fires when ``uniform < probability``; draws only in range (``schedulers.py:9-28``)."""

    def __init__(self, probability: float = 0.5, start=0, end=np.inf, seed=None) -> None:
        super().__init__(start, end)
        self.probability = probability
        self.seed = seed


class DecayingProbabilityScheduler(base.Scheduler):
    """Fires when ``uniform < p0 * exp(-decay_rate * t)`` (``schedulers.py:143-177``)."""

    def __init__(self, initial_probability: float, decay_rate: float, start=0, end=np.inf,
                 seed=None) -> None:
        super().__init__(start, end)
        self.initial_probability = initial_probability
        self.decay_rate = decay_rate
        self.seed = seed


class MemorylessScheduler(base.Scheduler):
    """Fires when ``t == transition_time`` and then re-arms ``transition_time = t +
    Geometric(p)`` (``schedulers.py:92-116``).  The first transition time is drawn at
    construction, as in the reference."""

    def __init__(self, p: float, start=0, end=np.inf, seed=None) -> None:
        super().__init__(start, end)
        self.p = p
        self.seed = seed
        self.transition_time = np.random.default_rng(seed=seed).geometric(p=self.p, size=(1,))


class CustomScheduler(base.Scheduler):
    """Arbitrary ``event_function(t) -> bool`` (``schedulers.py:31-43``).  A Python callable
    cannot run on the device: the compiler pre-evaluates it for ``t in [0, horizon]`` into a
    fire bitmap (``horizon`` = the env's episode limit), calling it once per ``t`` in
    increasing order."""

    def __init__(self, event_function, start=0, end=np.inf) -> None:
        super().__init__(start, end)
        self.event_function = event_function


__all__ = [
    "ContinuousScheduler", "PeriodicScheduler", "DiscreteScheduler", "BurstScheduler",
    "WindowScheduler", "RandomScheduler", "DecayingProbabilityScheduler",
    "MemorylessScheduler", "CustomScheduler",
]
