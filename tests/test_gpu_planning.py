"""GPU: planning copies (nsgym_fanout / get_planning_env) against the oracle port and the golden
vectors generated from the REAL reference; fan-out rollouts; snapshot / restore."""
import glob
import os

import numpy as np
import pytest

from oracle import harness
from tests import parity_util as pu
from tests.planning_cases import PLAN_CASES

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "planning")
GPU_CASES = sorted(k for k, v in PLAN_CASES.items() if v["gpu"])


@pytest.mark.parametrize("name", GPU_CASES)
def test_planning_copies_match_oracle(name):
    import warnings

    sc = PLAN_CASES[name]
    n = 40
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref0, ref1, (actions, u, z) = pu.oracle_planning_trace(harness.port_envs, sc, n, seed=77)
    got0, got1 = pu.gpu_planning_trace(sc, n, actions, u, z)
    pu.compare(ref0, got0, float_obs_rtol=1e-6, name=name + " (roots)")
    pu.compare(ref1, got1, float_obs_rtol=1e-6, name=name + " (copies)")


@pytest.mark.parametrize("name", GPU_CASES)
def test_planning_copies_match_golden(name):
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    ref = {k: g[k] for k in g.files if k not in ("actions", "uniforms", "normals")}
    n = g["uniforms"].shape[2]
    _, got = pu.gpu_planning_trace(PLAN_CASES[name], n, g["actions"], g["uniforms"], g["normals"])
    pu.compare(ref, got, float_obs_rtol=1e-6, name=name)


def _cartpole(n, **kw):
    import ns_gym_b200.schedulers as PS
    import ns_gym_b200.update_functions as PU
    from ns_gym_b200.vector_env import NSVectorEnv

    tp = {"masspole": PU.IncrementUpdate(PS.ContinuousScheduler(), k=0.01),
          "gravity": PU.RandomWalk(PS.PeriodicScheduler(3), mu=0.0, sigma=0.1)}
    env = NSVectorEnv("CartPole-v1", tp, n, precision="fp64", seed=5, **kw)
    env.reset(seed=5)
    return env


def test_fanout_rollout_equals_stepping_the_copies():
    """m lanes per root: the fused K-step rollout of the copies (uniform-random device policy, lanes
    stop at their first episode end) equals stepping the same copies one launch at a time."""
    import torch

    from tests import philox_np

    n, m, K = 256, 8, 40
    env = _cartpole(n, change_notification=True, delta_change_notification=True)
    a = torch.zeros(n, dtype=torch.int32, device=env.device)
    for k in range(6):
        env.step(a + (k & 1))
    plan_a = env.get_planning_env(fanout=m, seed=99)
    plan_b = env.get_planning_env(fanout=m, seed=99)
    # copies start from their root
    root_state = env.buffers["state"].repeat_interleave(m, dim=0)
    assert torch.equal(plan_a.buffers["state"], root_state)
    assert torch.equal(plan_a.buffers["theta"], env.buffers["theta"].repeat_interleave(m, dim=1))
    assert torch.equal(plan_a.relative_time(), env.relative_time().repeat_interleave(m))
    step0 = int(plan_a.lib.nsgym_step_index(plan_a._h))
    ret, length = plan_a.rollout(K, gamma=0.97)
    gids = np.arange(n * m, dtype=np.uint64)
    acc = torch.zeros(n * m, dtype=torch.float32, device=env.device)
    disc = torch.ones(n * m, dtype=torch.float32, device=env.device)
    alive = torch.ones(n * m, dtype=torch.bool, device=env.device)
    steps = torch.zeros(n * m, dtype=torch.int32, device=env.device)
    frozen_state = plan_b.buffers["state"].clone()
    for k in range(K):
        act = torch.as_tensor(philox_np.policy_actions("cartpole", gids, step0 + k, 99)).to(env.device)
        obs, r, term, trunc, info = plan_b.step(act)
        acc += torch.where(alive, disc * r, torch.zeros_like(acc))
        steps += alive.int()
        disc = disc * 0.97
        newly = alive & (term | trunc)
        frozen_state = torch.where((alive)[:, None], plan_b.buffers["state"], frozen_state)
        alive = alive & ~newly
    torch.cuda.synchronize()
    assert torch.equal(length, steps)
    np.testing.assert_allclose(ret.cpu().numpy(), acc.cpu().numpy(), rtol=1e-6, atol=1e-6)
    assert torch.equal(plan_a.buffers["state"], frozen_state)       # stopped lanes keep their final state
    # theta frozen in the copies (in_sim_change False), muted notifications
    assert torch.equal(plan_a.buffers["theta"], env.buffers["theta"].repeat_interleave(m, dim=1))


def test_untold_copy_starts_from_initial_parameters():
    import torch

    env = _cartpole(64)                     # no notifications: the agent does not know theta
    a = torch.zeros(64, dtype=torch.int32, device=env.device)
    for _ in range(5):
        env.step(a)
    plan = env.get_planning_env(fanout=2)
    th = plan.theta()
    assert torch.all(th["masspole"] == 0.1) and torch.all(th["gravity"] == 9.8)
    assert not torch.all(env.theta()["masspole"] == 0.1)


def test_snapshot_restore_replays_bit_for_bit():
    import torch

    env = _cartpole(512, change_notification=True)
    a = torch.ones(512, dtype=torch.int32, device=env.device)
    for _ in range(4):
        env.step(a)
    snap = env.snapshot()
    outs = []
    for rep in range(2):
        env.restore(snap)
        for _ in range(25):
            env.step(a)
        outs.append({k: v.clone() for k, v in env.buffers.items() if v is not None})
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), k
