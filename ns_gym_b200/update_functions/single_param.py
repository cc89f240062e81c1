"""Scalar update-function descriptions (same names / constructor arguments / attributes as
``ns_gym/update_functions/single_param.py:20-508``).  ``Y`` is the current value; the rule
runs only when the scheduler fires and the reported change is ``new - old``
(``ns_gym/base.py:172-182``).  Each lowers to one ``upd_op`` of the step kernel.
"""
from __future__ import annotations

from typing import Union

from .. import base


class IncrementUpdate(base.UpdateFn):
    """``Y + k`` (``single_param.py:154-175``)."""

    def __init__(self, scheduler, k: float) -> None:
        super().__init__(scheduler)
        self.k = k


class DecrementUpdate(base.UpdateFn):
    """``Y - k`` (``single_param.py:178-199``)."""

    def __init__(self, scheduler, k) -> None:
        super().__init__(scheduler)
        self.k = k


class DeterministicTrend(base.UpdateFn):
    """``Y + slope * t`` (``single_param.py:20-40``)."""

    def __init__(self, scheduler, slope: float) -> None:
        super().__init__(scheduler)
        self.slope = slope


class PolynomialTrend(base.UpdateFn):
    """``Y + sum_i coeffs[i] * t**(i+1)`` (``single_param.py:451-473``)."""

    def __init__(self, scheduler, coeffs: list) -> None:
        super().__init__(scheduler)
        self.coeffs = coeffs


class GeometricProgression(base.UpdateFn):
    """``Y * r`` (``single_param.py:290-307``)."""

    def __init__(self, scheduler, r):
        super().__init__(scheduler)
        self.r = r


class ExponentialDecay(base.UpdateFn):
    """``Y * exp(-decay_rate * t)``, compounding on the current value
    (``single_param.py:266-287``)."""

    def __init__(self, scheduler, decay_rate: float) -> None:
        super().__init__(scheduler)
        self.decay_rate = decay_rate


class OscillatingUpdate(base.UpdateFn):
    """``Y + delta * sin(t)`` (``single_param.py:243-264``)."""

    def __init__(self, scheduler, delta: float) -> None:
        super().__init__(scheduler)
        self.delta = delta


class SigmoidTransition(base.UpdateFn):
    """``a + (b - a) / (1 + exp(-k (t - t0)))``, replaces Y (``single_param.py:349-385``)."""

    def __init__(self, scheduler, a: float, b: float, k: float, t0: float) -> None:
        super().__init__(scheduler)
        self.a = a
        self.b = b
        self.k = k
        self.t0 = t0


class LinearInterpolation(base.UpdateFn):
    """``start + (end - start) * min(t / T, 1)``, replaces Y (``single_param.py:476-508``)."""

    def __init__(self, scheduler, start_val: float, end_val: float, T: int) -> None:
        super().__init__(scheduler)
        self.start_val = start_val
        self.end_val = end_val
        self.T = T


class StepWiseUpdate(base.UpdateFn):
    """Next value of ``param_list`` per fire; an exhausted list keeps Y but still reports
    a change flag (``single_param.py:202-223``)."""

    def __init__(self, scheduler, param_list: list) -> None:
        super().__init__(scheduler)
        self.param_list = param_list


class CyclicUpdate(base.UpdateFn):
    """Cycles through ``value_list`` (``single_param.py:388-408``)."""

    def __init__(self, scheduler, value_list: list) -> None:
        super().__init__(scheduler)
        self.value_list = value_list
        self._index = 0


class NoUpdate(base.UpdateFn):
    """Keeps Y; flag 1 and delta 0 when the scheduler fires (``single_param.py:226-240``)."""

    def __init__(self, scheduler) -> None:
        super().__init__(scheduler)


class RandomWalk(base.UpdateFn):
    """``Y + N(mu, sigma)`` (``single_param.py:84-113``)."""

    def __init__(self, scheduler, mu: Union[float, int] = 0, sigma: Union[float, int] = 1,
                 seed=None) -> None:
        super().__init__(scheduler)
        self.mu = mu
        self.sigma = sigma
        self.seed = seed


class RandomWalkWithDrift(base.UpdateFn):
    """``alpha + Y + N(mu, sigma)`` (``single_param.py:116-151``)."""

    def __init__(self, scheduler, alpha: float, mu: float, sigma: float,
                 seed: Union[int, None] = None) -> None:
        super().__init__(scheduler)
        self.mu = mu
        self.sigma = sigma
        self.alpha = alpha
        self.seed = seed


class RandomWalkWithDriftAndTrend(base.UpdateFn):
    """``alpha + Y + N(mu, sigma) + slope * t`` (``single_param.py:43-81``)."""

    def __init__(self, scheduler, alpha: float, mu: float, sigma: float, slope: float,
                 seed: Union[int, None] = None) -> None:
        super().__init__(scheduler)
        self.mu = mu
        self.sigma = sigma
        self.alpha = alpha
        self.slope = slope
        self.seed = seed


class OrnsteinUhlenbeck(base.UpdateFn):
    """``Y + theta (mu - Y) + N(0, sigma)``; no draw when ``sigma == 0``
    (``single_param.py:310-346``)."""

    def __init__(self, scheduler, theta: float, mu: float, sigma: float = 0.0,
                 seed: Union[int, None] = None) -> None:
        super().__init__(scheduler)
        self.theta = theta
        self.mu = mu
        self.sigma = sigma
        self.seed = seed


class BoundedRandomWalk(base.UpdateFn):
    """``clip(Y + N(mu, sigma), lo, hi)`` (``single_param.py:411-448``)."""

    def __init__(self, scheduler, mu: float, sigma: float, lo: float, hi: float,
                 seed: Union[int, None] = None) -> None:
        super().__init__(scheduler)
        self.mu = mu
        self.sigma = sigma
        self.lo = lo
        self.hi = hi
        self.seed = seed


__all__ = [
    "BoundedRandomWalk", "CyclicUpdate", "DecrementUpdate", "DeterministicTrend",
    "ExponentialDecay", "GeometricProgression", "IncrementUpdate", "LinearInterpolation",
    "NoUpdate", "OrnsteinUhlenbeck", "OscillatingUpdate", "PolynomialTrend", "RandomWalk",
    "RandomWalkWithDrift", "RandomWalkWithDriftAndTrend", "SigmoidTransition",
    "StepWiseUpdate",
]
