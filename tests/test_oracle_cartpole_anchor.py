"""CPU: the oracle's restated gymnasium CartPole step against the reference's OWN in-tree copy of
the CartPole transition (rats-experiments/code/envs/nscartpole_v0.py:76-135), through known-answer
vectors generated from that file (tests/golden/make_cartpole_anchor.py).

gymnasium itself cannot be imported here or on the GPU box, so its dynamics stay formally
unpinned; for CartPole this closes the gap up to the association of one product (an ulp)."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "anchors", "cartpole_anchor.npz")


def _oracle_step(theta, state, action):
    from oracle.gym_restated import CartPoleEnv

    env = CartPoleEnv()
    nxt = np.zeros_like(state)
    done = np.zeros(len(state), dtype=bool)
    for k in range(len(state)):
        env.gravity, env.masscart, env.masspole, env.force_mag, env.tau, env.length = (float(v) for v in theta[k])
        env.total_mass = env.masspole + env.masscart
        env.polemass_length = env.masspole * env.length
        env.state = np.array(state[k], dtype=np.float64)
        env.steps_beyond_terminated = None
        _, _, terminated, _, _ = env.step(int(action[k]))
        nxt[k] = env.state
        done[k] = terminated
    return nxt, done


def test_oracle_cartpole_step_matches_the_reference_copy():
    g = np.load(GOLDEN)
    nxt, done = _oracle_step(g["theta"], g["state"], g["action"])
    # one differently associated product: agreement to a few ulps of the largest term
    np.testing.assert_allclose(nxt, g["next_state"], rtol=2e-14, atol=1e-15)
    # termination: identical except where a coordinate sits within rounding of a threshold
    differ = done != g["done"]
    assert differ.sum() == 0, f"{differ.sum()} termination flags differ"
    assert 100 < int(done.sum()) < len(done) - 100          # both outcomes are exercised


@pytest.mark.skipif(not os.path.exists("/root/reference/ns_gym"), reason="needs the reference tree (build container)")
def test_anchor_vectors_come_from_the_reference_file():
    """The committed fixture is what the reference's own file computes today."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_cartpole_anchor", os.path.join(HERE, "golden", "make_cartpole_anchor.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    g = np.load(GOLDEN)
    th, st, ac = mk.inputs()
    assert np.array_equal(th, g["theta"]) and np.array_equal(st, g["state"]) and np.array_equal(ac, g["action"])
    nx, dn = mk.reference_outputs(th[:512], st[:512], ac[:512])
    assert np.array_equal(nx, g["next_state"][:512]) and np.array_equal(dn, g["done"][:512])


@pytest.mark.gpu
def test_kernel_cartpole_step_matches_the_reference_copy():
    """The CUDA path's fp64 CartPole dynamics against the same vectors (the transitions with the
    stock parameters: unbound parameters are launch constants)."""
    import torch

    from ns_gym_b200 import native as nv
    from ns_gym_b200.vector_env import NSVectorEnv

    g = np.load(GOLDEN)
    stock = np.all(g["theta"] == np.array([9.8, 1.0, 0.1, 10.0, 0.02, 0.5]), axis=1)
    state, action = g["state"][stock], g["action"][stock]
    n = int(stock.sum())
    assert n >= 256
    env = NSVectorEnv("CartPole-v1", {}, n, precision="fp64", autoreset="none", seed=0)
    env.reset(seed=0)
    env.buffers["state"].copy_(torch.as_tensor(state, device=env.device))
    env.step_raw(torch.as_tensor(action, dtype=torch.int32, device=env.device))
    torch.cuda.synchronize()
    np.testing.assert_allclose(env.buffers["state"].cpu().numpy(), g["next_state"][stock], rtol=2e-14, atol=1e-15)
    terminated = (env.buffers["flags"].cpu().numpy() & nv.FLAG_TERMINATED) != 0
    assert np.array_equal(terminated, g["done"][stock])
