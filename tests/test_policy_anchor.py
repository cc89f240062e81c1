"""Behavioural anchors for the Acrobot and Pendulum dynamics: the reference's own pre-trained agents
(SB3 PPO / A2C / DDPG weights under ns_gym/evaluate/evaluation_model_weights, fitted by the reference's
authors on the REAL gymnasium environments) still solve the tasks on the oracle's restated
environments (CPU) and on the CUDA kernels (GPU).  tests/golden/make_policy_anchor.py converts the
networks and records the oracle's returns.

Acrobot-v1 counts as solved around -100 per episode, Pendulum-v1 around -150 .. -200; a random policy
gets -500 and about -1200.  A wrong sign, gain, clip or reward in the restated dynamics collapses the
return of a policy fitted to the true ones -- this is corroboration of SURVEY rows a6 / a7 by a
reference-held artefact, not a bit-level pin."""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_policy_anchor", os.path.join(HERE, "golden", "make_policy_anchor.py"))
MK = importlib.util.module_from_spec(spec)
spec.loader.exec_module(MK)

SOLVED = {"acrobot_ppo": -90.0, "acrobot_a2c": -120.0, "pendulum_ppo": -250.0, "pendulum_ddpg": -200.0}


@pytest.mark.parametrize("name", sorted(SOLVED))
def test_reference_agents_solve_the_restated_envs(name):
    pol = MK.load_saved()[name]
    rets = MK.oracle_returns(pol, episodes=12)
    assert np.allclose(rets, pol["oracle_returns"][:12])           # the fixture is what the oracle does today
    assert np.mean(pol["oracle_returns"]) >= SOLVED[name], (name, np.mean(pol["oracle_returns"]))


def test_a_random_policy_does_not():
    """The thresholds separate trained from untrained behaviour on the restated envs."""
    from oracle import gym_restated as G

    r = np.random.default_rng(0)
    for env_id, thr in (("Acrobot-v1", -400.0), ("Pendulum-v1", -900.0)):
        env = G.make(env_id)
        rets = []
        for ep in range(5):
            env.reset(seed=ep)
            total, done = 0.0, False
            while not done:
                a = int(r.integers(0, 3)) if "Acrobot" in env_id else np.array([r.uniform(-2, 2)], dtype=np.float32)
                _, rew, term, trunc, _ = env.step(a)
                total += float(rew)
                done = term or trunc
            rets.append(total)
        assert np.mean(rets) < thr


@pytest.mark.skipif(not os.path.exists("/root/reference/ns_gym"), reason="needs the reference tree (build container)")
def test_policy_fixture_comes_from_the_reference_files():
    saved, fresh = MK.load_saved(), MK.load_policies()
    assert set(saved) == set(fresh)
    for name in saved:
        for (w0, b0), (w1, b1) in zip(saved[name]["layers"], fresh[name]["layers"]):
            assert np.array_equal(w0, w1) and np.array_equal(b0, b1)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "fp64"])
@pytest.mark.parametrize("name", sorted(SOLVED))
def test_reference_agents_solve_the_cuda_envs(name, precision):
    """2048 episodes per agent on the step kernels (stationary: no tunable parameter bound); the policy
    network runs in torch on the observations the kernel writes."""
    import torch

    from ns_gym_b200 import native as nv
    from ns_gym_b200.vector_env import NSVectorEnv

    pol = MK.load_saved()[name]
    n = 2048
    env = NSVectorEnv(pol["env_id"], {}, n, precision=precision, autoreset="none", seed=11, want_obs=True)
    obs, _ = env.reset()
    dev = env.device
    layers = [(torch.as_tensor(w, device=dev), torch.as_tensor(b, device=dev)) for w, b in pol["layers"]]
    total = torch.zeros(n, dtype=torch.float64, device=dev)
    alive = torch.ones(n, dtype=torch.bool, device=dev)
    limit = int(env.program.spec.max_episode_steps)
    for k in range(limit):
        x = env.observation().float()
        for j, (w, b) in enumerate(layers):
            x = x @ w.T + b
            if j + 1 < len(layers):
                x = torch.relu(x) if pol["act"] == "relu" else torch.tanh(x)
        if pol["head"] == "argmax":
            a = torch.argmax(x, dim=1).to(torch.int32)
        else:
            a = (torch.clamp(x, -2.0, 2.0) if pol["head"] == "mean" else 2.0 * torch.tanh(x)).reshape(n).to(env.real)
        env.step_raw(a.contiguous())
        total += torch.where(alive, env.buffers["reward"].double(), torch.zeros_like(total))
        alive &= (env.buffers["flags"] & (nv.FLAG_TERMINATED | nv.FLAG_TRUNCATED)) == 0
        if not bool(alive.any()):
            break
    mean = float(total.mean())
    assert mean >= SOLVED[name], (name, precision, mean)
    # and the same behaviour as on the oracle: 30 oracle episodes against 2048 kernel episodes
    ref = np.array(pol["oracle_returns"])
    assert abs(mean - ref.mean()) < 4.0 * ref.std() / np.sqrt(len(ref)) + 5.0, (mean, ref.mean())
