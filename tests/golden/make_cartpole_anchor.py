"""Known-answer vectors for the CartPole dynamics from the reference's OWN in-tree copy of them
(run in the build container only).

    python tests/golden/make_cartpole_anchor.py

gymnasium is absent from /root/reference (third-party, uv.lock:958-959), but the reference carries
a restatement of the CartPole transition: ns_gym/benchmark_algorithms/rats-experiments/code/envs/
nscartpole_v0.py:76-135 (`NSCartPoleV0.transition`).  With the cart's tilt switched off
(alpha_max_radians = 0: cos alpha = 1, sin alpha = 0, so force and gravity pass through exactly)
and two actions it is gymnasium's Euler step up to the association of one product
(`polemass_length * theta_dot * theta_dot` vs `polemass_length * square(theta_dot)`), i.e. to an
ulp.  The file is loaded IN PLACE with a stub `gym` module; outputs go to
tests/golden/anchors/cartpole_anchor.npz, which travels to the GPU box.
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference/ns_gym/benchmark_algorithms/rats-experiments/code/envs/nscartpole_v0.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "anchors", "cartpole_anchor.npz")
N = 4096


def load_reference_class():
    """NSCartPoleV0 from the reference file, `gym` stubbed (Env / spaces / seeding only)."""
    gym = types.ModuleType("gym")
    gym.Env = object
    spaces = types.ModuleType("gym.spaces")
    spaces.Discrete = lambda n: ("Discrete", n)
    spaces.Box = lambda lo, hi: ("Box", lo, hi)
    utils = types.ModuleType("gym.utils")
    seeding = types.ModuleType("gym.utils.seeding")
    seeding.np_random = lambda seed=None: (np.random.RandomState(seed), seed)
    gym.spaces, gym.utils, utils.seeding = spaces, utils, seeding
    saved = {k: sys.modules.get(k) for k in ("gym", "gym.spaces", "gym.utils", "gym.utils.seeding")}
    sys.modules.update({"gym": gym, "gym.spaces": spaces, "gym.utils": utils, "gym.utils.seeding": seeding})
    try:
        spec = importlib.util.spec_from_file_location("_ref_nscartpole_v0", REF)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod.NSCartPoleV0


def inputs(seed=20261018):
    r = np.random.default_rng(seed)
    theta = np.stack([r.uniform(5.0, 15.0, N),      # gravity
                      r.uniform(0.5, 2.0, N),       # masscart
                      r.uniform(0.05, 0.5, N),      # masspole
                      r.uniform(5.0, 15.0, N),      # force_mag
                      r.uniform(0.01, 0.05, N),     # tau
                      r.uniform(0.25, 1.0, N)], 1)  # length
    theta[: N // 8] = [9.8, 1.0, 0.1, 10.0, 0.02, 0.5]                     # the stock parameters
    state = np.stack([r.uniform(-2.6, 2.6, N), r.uniform(-3, 3, N), r.uniform(-0.25, 0.25, N), r.uniform(-3, 3, N)], 1)
    state[N // 2:] = r.uniform(-0.05, 0.05, (N - N // 2, 4))                # reset-like states
    action = r.integers(0, 2, N)
    return theta, state, action


def reference_outputs(theta, state, action):
    cls = load_reference_class()
    env = cls()
    env.alpha_max_radians = 0.0
    env.nb_actions = 2
    nxt = np.zeros_like(state)
    done = np.zeros(len(state), dtype=bool)
    for k in range(len(state)):
        env.gravity, env.masscart, env.masspole, env.force_mag, env.tau, env.length = (float(v) for v in theta[k])
        env.total_mass = env.masspole + env.masscart                         # classic_control.py:426-444
        env.polemass_length = env.masspole * env.length
        s, _, d = env.transition(tuple(float(v) for v in state[k]) + (0.0,), int(action[k]), False)
        nxt[k] = s[:4]
        done[k] = d
    return nxt, done


if __name__ == "__main__":
    th, st, ac = inputs()
    nx, dn = reference_outputs(th, st, ac)
    np.savez_compressed(OUT, theta=th, state=st, action=ac, next_state=nx, done=dn)
    print(f"wrote {OUT}: {len(st)} transitions, {int(dn.sum())} terminal")
