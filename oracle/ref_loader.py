"""Import the REAL reference (``/root/reference/ns_gym``) in this container.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference is pure Python but needs
``gymnasium`` and ``mujoco`` at import time (``ns_gym/base.py:5-6,1156``,
``ns_gym/wrappers/mujoco_env.py:1-2``); neither is installed here, so
``oracle.gym_restated.install_shim()`` supplies a minimal ``gymnasium`` whose ``make()``
returns the restated base envs, and a bare ``mujoco`` module.  Everything *inside*
``ns_gym`` then runs verbatim and unmodified from where it lies (never copied).

``/root/reference`` does not exist on the GPU box: callers must check ``available()``.
"""
from __future__ import annotations

import os
import sys

REFERENCE_ROOT = os.environ.get("NSGYM_REFERENCE_ROOT", "/root/reference")

_cached = None


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "ns_gym", "base.py"))


def load():
    """Return the imported reference package (module ``ns_gym``)."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError(f"reference tree not present at {REFERENCE_ROOT}")
    from . import gym_restated

    gym_restated.install_shim()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import ns_gym  # noqa: E402  (the reference itself)

    assert os.path.realpath(ns_gym.__file__).startswith(os.path.realpath(REFERENCE_ROOT)), (
        f"imported ns_gym from {ns_gym.__file__}, expected the reference tree"
    )
    _cached = ns_gym
    return ns_gym


def gym():
    """The ``gymnasium`` module the reference is running on (shim or real)."""
    load()
    return sys.modules["gymnasium"]
