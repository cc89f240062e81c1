"""Slip-distribution update-function descriptions (same names / constructor arguments /
attributes as ``ns_gym/update_functions/distribution.py:11-356``).  ``p`` is the current
list of 3 (FrozenLake, Bridge) or 4 (CliffWalking) probabilities; the reported change is
the 1-Wasserstein distance on indices (``ns_gym/base.py:192-203``, ``ns_gym/utils.py:55-94``).
"""
from __future__ import annotations

from typing import Optional

from .. import base


class DistributionIncrementUpdate(base.UpdateDistributionFn):
    """``p0 = min(1, p0 + k)``, rest ``(1 - p0) / (n - 1)``; no lower clamp
    (``distribution.py:41-67``)."""

    def __init__(self, scheduler, k: float) -> None:
        super().__init__(scheduler)
        self.k = k


class DistributionDecrementUpdate(base.UpdateDistributionFn):
    """``p0 = max(0, p0 - k)``, rest ``(1 - p0) / (n - 1)`` (``distribution.py:70-97``)."""

    def __init__(self, scheduler, k: float) -> None:
        super().__init__(scheduler)
        self.k = k


class DistributionStepWiseUpdate(base.UpdateDistributionFn):
    """Next distribution of ``update_values`` per fire (``distribution.py:100-130``)."""

    def __init__(self, scheduler, update_values: list) -> None:
        super().__init__(scheduler)
        self.update_values = update_values


class DistributionCyclicUpdate(base.UpdateDistributionFn):
    """Cycles through ``dist_list`` (``distribution.py:334-356``)."""

    def __init__(self, scheduler, dist_list: list) -> None:
        super().__init__(scheduler)
        self.dist_list = dist_list
        self._index = 0


class DistributionNoUpdate(base.UpdateDistributionFn):
    """Keeps p (``distribution.py:217-231``)."""

    def __init__(self, scheduler) -> None:
        super().__init__(scheduler)


class UniformDrift(base.UpdateDistributionFn):
    """``(1 - rate) p + rate / n`` (``distribution.py:234-261``)."""

    def __init__(self, scheduler, rate: float) -> None:
        super().__init__(scheduler)
        self.rate = rate


class TargetReversion(base.UpdateDistributionFn):
    """``p + theta (target - p)`` (``distribution.py:264-293``)."""

    def __init__(self, scheduler, target: list, theta: float) -> None:
        super().__init__(scheduler)
        self.target = target
        self.theta = theta


class DistributionLinearInterpolation(base.UpdateDistributionFn):
    """``start + (end - start) * min(t / T, 1)`` (``distribution.py:296-331``)."""

    def __init__(self, scheduler, start_dist: list, end_dist: list, T: int) -> None:
        super().__init__(scheduler)
        self.start_dist = start_dist
        self.end_dist = end_dist
        self.T = T


class RandomCategorical(base.UpdateDistributionFn):
    """Fresh ``Dirichlet(1, .., 1)`` draw per fire (``distribution.py:11-38``).  On the device:
    standard exponentials from the env's Philox stream, scaled by the reciprocal of their sum."""

    def __init__(self, scheduler, seed: Optional[int] = None) -> None:
        super().__init__(scheduler)
        self.seed = seed


class LCBoundedDistrubutionUpdate(base.UpdateDistributionFn):
    """Redraws the inner rule until ``W1(p, p') <= L * |t - prev_time|`` (``distribution.py:133-183``;
    the class name keeps the reference's spelling).  The inner rule defaults to
    ``RandomCategorical``, the only one that is redrawn on the device (a deterministic inner rule
    either passes at once or can never pass: the reference raises, the device flags it)."""

    def __init__(self, scheduler, L: float, update_fn=None) -> None:
        super().__init__(scheduler)
        self.L = L
        if update_fn is not None:
            assert isinstance(update_fn, type) and issubclass(update_fn, base.UpdateDistributionFn), (
                "update_fn must be a subclass of base.UpdateDistributionFn")
        self.update_fn = update_fn


__all__ = [
    "DistributionCyclicUpdate", "DistributionDecrementUpdate", "DistributionIncrementUpdate",
    "DistributionLinearInterpolation", "DistributionNoUpdate", "DistributionStepWiseUpdate",
    "LCBoundedDistrubutionUpdate", "RandomCategorical", "TargetReversion", "UniformDrift",
]
