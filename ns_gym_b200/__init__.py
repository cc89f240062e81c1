"""ns_gym_b200 -- B200-native batched simulator for ns_gym's non-stationary env-step path.

Host side: Python descriptions with the reference's plugin surface (``schedulers``,
``update_functions``, ``wrappers``) compiled into an opcode-and-coefficient table
(``compile``), executed by hand-written sm_100a CUDA kernels behind a C ABI
(``include/nsgym_b200.h``, ``ns_gym_b200/csrc``).  There is no CPU execution path: every
compute entry point raises if the CUDA library is missing.
"""
from . import base, schedulers, update_functions  # noqa: F401

__version__ = "0.1.0"

_LAZY = {"compile", "native", "vector_env", "wrappers", "distributed", "build", "spaces", "utils"}


def make(env_id: str, num_envs: int = 1, **kwargs):
    """Batched stand-in for ``gym.make``: see ``ns_gym_b200.wrappers.make``."""
    from .wrappers import make as _make

    return _make(env_id, num_envs, **kwargs)


def __getattr__(name):
    if name in _LAZY:
        import importlib

        mod = importlib.import_module(f".{name}", __name__)
        globals()[name] = mod
        return mod
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
