"""Known-answer vectors for the gridworld sampling rule and the clamped move from the reference's
OWN in-tree copies of them (run in the build container only).

    python tests/golden/make_gridworld_anchor.py

gymnasium is absent from /root/reference (third-party, uv.lock:958-959), so the two pieces of
gymnasium arithmetic the NS FrozenLake / CliffWalking wrappers lean on -- `categorical_sample`
(first index whose running sum exceeds one uniform, 0 when none does) and the clamp-to-grid move
-- are restated in oracle/gym_restated.py.  The reference carries its own copies of both:

    ns_gym/benchmark_algorithms/rats-experiments/code/envs/nsfrozenlake_v0.py:61-68, 215-228
    ns_gym/benchmark_algorithms/rats-experiments/code/envs/nscliff_v0.py:40-47, 101-114
    ns_gym/benchmark_algorithms/rats-experiments/code/envs/nsbridge_v0.py:30-37, 99-112

The three files are loaded IN PLACE (stub `gym`, `six`, `matplotlib` and the package-relative
`..utils.distribution`); the module-level `categorical_sample` and the classes' `inc` / `to_m` /
`to_s` are called on seeded inputs and the results go to tests/golden/anchors/gridworld_anchor.npz,
which travels to the GPU box.
"""
import importlib.util
import os
import sys
import types

import numpy as np

ENVS = "/root/reference/ns_gym/benchmark_algorithms/rats-experiments/code/envs"
FILES = {"frozenlake": "nsfrozenlake_v0.py", "cliff": "nscliff_v0.py", "bridge": "nsbridge_v0.py"}
CLASSES = {"frozenlake": "NSFrozenLakeV0", "cliff": "NSCliffV0", "bridge": "NSBridgeV0"}
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "anchors", "gridworld_anchor.npz")
N = 8192
# grids the product path runs on: FrozenLake 4x4 / 8x8, Bridge 5x8, CliffWalking 4x12
SHAPES = [(4, 4), (8, 8), (5, 8), (4, 12)]


def load_reference_modules():
    """{name: module} of the three reference files, third-party imports stubbed."""
    stubs = {}
    gym = types.ModuleType("gym")
    gym.Env = object
    gym.spaces = types.ModuleType("gym.spaces")
    gym.utils = types.ModuleType("gym.utils")
    stubs.update({"gym": gym, "gym.spaces": gym.spaces, "gym.utils": gym.utils})
    six = types.ModuleType("six")
    six.StringIO = object
    stubs["six"] = six
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = types.ModuleType("matplotlib.pyplot")
    stubs.update({"matplotlib": mpl, "matplotlib.pyplot": mpl.pyplot})
    # package context for `from ..utils.distribution import ...`
    pkg = types.ModuleType("_refrats")
    pkg.__path__ = []
    envs = types.ModuleType("_refrats.envs")
    envs.__path__ = []
    utils = types.ModuleType("_refrats.utils")
    utils.__path__ = []
    dist = types.ModuleType("_refrats.utils.distribution")
    dist.wass_dual = None           # only used when a transition matrix is generated (not here)
    dist.__all__ = []
    stubs.update({"_refrats": pkg, "_refrats.envs": envs, "_refrats.utils": utils,
                  "_refrats.utils.distribution": dist})
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    mods = {}
    try:
        for name, fn in FILES.items():
            spec = importlib.util.spec_from_file_location(f"_refrats.envs.{fn[:-3]}", os.path.join(ENVS, fn))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mods[name] = mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mods


class _Rand:
    """np_random stand-in: `.rand()` hands out the injected uniform (RandomState API of the copies)."""

    def __init__(self, u):
        self.u = u

    def rand(self):
        return self.u


def inputs(seed=20261019):
    """(p3[N,3], p4[N,4], u[N]) with the edge cases the rule has: u equal to a running sum (strict
    `>`), running sums that never exceed u (-> index 0), zeros, one-hot, unnormalised weights."""
    r = np.random.default_rng(seed)
    p3 = r.dirichlet([1.0, 1.0, 1.0], N)
    p4 = r.dirichlet([1.0, 1.0, 1.0, 1.0], N)
    u = r.random(N)
    k = N // 16
    p3[0 * k:1 * k] = [1.0, 0.0, 0.0]
    p4[0 * k:1 * k] = [1.0, 0.0, 0.0, 0.0]
    p3[1 * k:2 * k, 1] = 0.0                                    # a zero in the middle (sum < 1)
    p4[1 * k:2 * k, 2] = 0.0
    sl = slice(2 * k, 3 * k)                                    # u exactly on a running sum
    u[sl] = np.where(r.random(k) < 0.5, p3[sl, 0], p3[sl, 0] + p3[sl, 1])
    sl = slice(3 * k, 4 * k)
    u[sl] = np.cumsum(p4[sl], 1)[np.arange(k), r.integers(0, 3, k)]
    sl = slice(4 * k, 5 * k)                                    # mass below u: falls through to 0
    p3[sl] *= r.uniform(0.0, 0.6, (k, 1)); p4[sl] *= r.uniform(0.0, 0.6, (k, 1))
    u[sl] = r.uniform(0.6, 1.0, k)
    sl = slice(5 * k, 6 * k)                                    # the drifting C2 / C5 shapes
    x = r.uniform(0.0, 1.0, k)
    p3[sl] = np.stack([x, (1.0 - x) / 2.0, (1.0 - x) / 2.0], 1)
    p4[sl] = np.stack([x, (1.0 - x) / 3.0, (1.0 - x) / 3.0, (1.0 - x) / 3.0], 1)
    sl = slice(6 * k, 7 * k)                                    # sums an ulp or two off 1
    p3[sl] = p3[sl] * (1.0 + r.integers(-3, 4, (k, 1)) * 2.0 ** -52)
    u[sl] = np.where(r.random(k) < 0.5, np.nextafter(1.0, 0.0), u[sl])
    sl = slice(7 * k, 8 * k)                                    # unnormalised weights
    p3[sl] *= r.uniform(1.0, 3.0, (k, 1)); p4[sl] *= r.uniform(1.0, 3.0, (k, 1))
    return p3, p4, u


def reference_outputs(p3, p4, u):
    mods = load_reference_modules()
    idx3 = np.zeros((3, len(u)), dtype=np.int64)
    idx4 = np.zeros((3, len(u)), dtype=np.int64)
    for m, name in enumerate(("frozenlake", "cliff", "bridge")):
        cs = mods[name].categorical_sample
        for k in range(len(u)):
            idx3[m, k] = cs(p3[k], _Rand(float(u[k])))
            idx4[m, k] = cs(p4[k], _Rand(float(u[k])))
    assert (idx3 == idx3[0]).all() and (idx4 == idx4[0]).all(), "the reference's three copies disagree"
    # clamped move: next cell of every (cell, direction) on each grid, by each copy's inc / to_m / to_s
    moves = {}
    for nrow, ncol in SHAPES:
        tabs = []
        for name in ("frozenlake", "cliff", "bridge"):
            cls = getattr(mods[name], CLASSES[name])
            obj = types.SimpleNamespace(nrow=nrow, ncol=ncol)
            tab = np.zeros((nrow * ncol, 4), dtype=np.int64)
            for s in range(nrow * ncol):
                row, col = cls.to_m(obj, s)
                for a in range(4):                   # LEFT DOWN RIGHT UP
                    r2, c2 = cls.inc(obj, row, col, a)
                    tab[s, a] = cls.to_s(obj, r2, c2)
            tabs.append(tab)
        assert all((t == tabs[0]).all() for t in tabs)
        moves[f"move_{nrow}x{ncol}"] = tabs[0]
    return idx3[0], idx4[0], moves


if __name__ == "__main__":
    p3, p4, u = inputs()
    i3, i4, moves = reference_outputs(p3, p4, u)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, p3=p3, p4=p4, u=u, idx3=i3, idx4=i4, **moves)
    print(f"wrote {OUT}: {len(u)} draws; idx3 counts {np.bincount(i3, minlength=3).tolist()}, "
          f"idx4 counts {np.bincount(i4, minlength=4).tolist()}; grids {sorted(moves)}")
