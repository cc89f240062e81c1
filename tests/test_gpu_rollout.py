"""GPU: the K-step fused rollout kernel equals K single-step launches fed with the same actions.

The rollout draws its actions from Philox block 13 of each env-step; tests/philox_np.py restates
the stream in NumPy, so the single-step path can be driven with identical actions.  Both paths
then run the same device arithmetic with the same native Philox draws (same seed, same step
indices, same global env ids): states, theta, t, returns and lengths must agree bit for bit."""
import numpy as np
import pytest

from tests import philox_np
from tests.cases import CASES

pytestmark = pytest.mark.gpu

ROLLOUT_CASES = [("c1_cartpole_readme", "cartpole", "fp32"), ("c1_cartpole_readme", "cartpole", "fp64"),
                 ("c3_acrobot", "acrobot", "fp32"), ("c3_mountaincar", "mountaincar", "fp64"),
                 ("c3_pendulum", "pendulum", "fp32"), ("c5_bridge_uniform", "grid", "fp64"),
                 ("c5_bridge_split", "grid", "fp64"), ("c2_frozenlake8_drift", "grid", "fp64"),
                 ("cliff_terminal", "grid", "fp64")]


def _make(case, n, precision, seed, offset=0, specialize=0):
    import ns_gym_b200.schedulers as PS
    import ns_gym_b200.update_functions as PU
    from ns_gym_b200.vector_env import NSVectorEnv

    env = NSVectorEnv(case["env_id"], case["params"](PS, PU), n, precision=precision, seed=seed,
                      env_id_offset=offset, **case.get("wrapper", {}), **case.get("make", {}))
    env.set_option("specialize", specialize)
    env.reset(seed=seed)
    return env


# specialize = 1: the program-specialised rollout kernel (nsgym_jit.cu) against the program-specialised
# single-step kernel -- the same device functions, compiled with the program as a constant
@pytest.mark.parametrize("specialize", [0, 1])
@pytest.mark.parametrize("name,kind,precision", ROLLOUT_CASES)
def test_rollout_equals_single_steps(name, kind, precision, specialize):
    import torch

    case, n, K, seed, offset = CASES[name], 4096, 37, 1234, 10_000
    a = _make(case, n, precision, seed, offset, specialize)
    b = _make(case, n, precision, seed, offset, specialize)
    step0 = int(a.lib.nsgym_step_index(a._h))
    assert step0 == int(b.lib.nsgym_step_index(b._h))
    ret, length = a.rollout(K, gamma=1.0)
    assert a.last_kernel_specialized == bool(specialize)
    gids = np.arange(offset, offset + n, dtype=np.uint64)
    acc = torch.zeros(n, dtype=torch.float32, device=b.device)
    steps_alive = torch.zeros(n, dtype=torch.int32, device=b.device)
    first = torch.ones(n, dtype=torch.bool, device=b.device)
    for k in range(K):
        act = philox_np.policy_actions(kind, gids, step0 + k, seed, precision)
        obs, r, term, trunc, info = b.step(torch.as_tensor(act))
        assert b.last_kernel_specialized == bool(specialize)
        was_reset = info["was_reset"]
        first &= ~was_reset
        steps_alive += (first & ~was_reset).int()
        acc += r
    torch.cuda.synchronize()
    for key in ("state", "theta", "t", "istate"):
        x, y = a.buffers[key], b.buffers[key]
        if x is not None:
            assert torch.equal(x, y), f"{name}: {key} differs after the rollout"
    assert torch.equal(ret, acc)
    assert torch.equal(length, steps_alive)
    assert int(a.lib.nsgym_step_index(a._h)) == step0 + K


def test_rollout_is_shard_invariant():
    """Global-id Philox keys: two half-size shards reproduce one full-size batch."""
    import torch

    case, n, K, seed = CASES["c5_bridge_uniform"], 2048, 25, 7
    full = _make(case, n, "fp64", seed, 0)
    lo = _make(case, n // 2, "fp64", seed, 0)
    hi = _make(case, n // 2, "fp64", seed, n // 2)
    rf, lf = full.rollout(K)
    r0, l0 = lo.rollout(K)
    r1, l1 = hi.rollout(K)
    assert torch.equal(rf, torch.cat([r0, r1])) and torch.equal(lf, torch.cat([l0, l1]))
    assert torch.equal(full.buffers["state"], torch.cat([lo.buffers["state"], hi.buffers["state"]]))


def test_frozen_planning_env_keeps_theta():
    """is_sim_env with in_sim_change False: theta frozen, t still advances (classic_control.py:70-75)."""
    import torch

    env = _make(CASES["c1_cartpole_readme"], 512, "fp64", 3)
    env.is_sim_env = True
    th0 = env.buffers["theta"].clone()
    t0 = env.relative_time().clone()
    a = torch.zeros(512, dtype=torch.int32, device=env.device)
    for _ in range(3):
        obs, r, term, trunc, info = env.step(a)
    assert torch.equal(env.buffers["theta"], th0)
    assert torch.equal(env.relative_time(), t0 + 3)
    assert all(int(v.sum()) == 0 for v in obs["env_change"].values())
    assert all(int(v.sum()) == 0 for v in info["Ground Truth Env Change"].values())


def test_device_philox_matches_numpy_restatement():
    """CartPole fp32 reset state = -0.05 + 0.1 * (word >> 8) * 2^-24 for the four words of block 0."""
    import torch

    n, seed, offset = 1000, 99, 123456789012
    env = _make(CASES["c1_cartpole_readme"], n, "fp32", seed, offset)
    step_of_reset = int(env.lib.nsgym_step_index(env._h)) - 1
    words = philox_np.block(np.arange(offset, offset + n, dtype=np.uint64), step_of_reset, 0, seed)
    want = np.stack([np.float32(-0.05) + (np.float32(0.05) - np.float32(-0.05)) *
                     ((w >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)) for w in words], 1)
    got = env.buffers["state"].cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-8)


@pytest.mark.parametrize("name,kind", [("c4_cartpole_rows", "cartpole"), ("het_cartpole_wide", "cartpole"),
                                       ("c4_frozenlake8_rows", "grid"), ("het_bridge_split", "grid")])
def test_heterogeneous_rollout_equals_single_steps(name, kind):
    """Per-env rows: the fused rollout and K single-step launches run the same device arithmetic."""
    import torch

    import ns_gym_b200.schedulers as PS
    import ns_gym_b200.update_functions as PU
    from ns_gym_b200.vector_env import NSVectorEnv

    case, n, K, seed = CASES[name], 512, 29, 77
    per_env = [case["params_of"](PS, PU, e) for e in range(n)]
    kw = dict(precision="fp64", seed=seed, **case["wrapper"], **case["make"])
    a = NSVectorEnv.heterogeneous(case["env_id"], per_env, **kw)
    b = NSVectorEnv.heterogeneous(case["env_id"], per_env, **kw)
    a.reset(seed=seed)
    b.reset(seed=seed)
    step0 = int(a.lib.nsgym_step_index(a._h))
    ret, length = a.rollout(K, gamma=1.0)
    gids = np.arange(n, dtype=np.uint64)
    acc = torch.zeros(n, dtype=torch.float32, device=b.device)
    for k in range(K):
        act = philox_np.policy_actions(kind, gids, step0 + k, seed)
        obs, r, term, trunc, info = b.step(torch.as_tensor(act))
        acc += r
    torch.cuda.synchronize()
    for key in ("state", "theta", "t", "istate"):
        x, y = a.buffers[key], b.buffers[key]
        if x is not None:
            assert torch.equal(x, y), f"{name}: {key} differs after the rollout"
    assert torch.equal(ret, acc)


def test_planning_fanout_on_a_heterogeneous_batch():
    import torch

    import ns_gym_b200.schedulers as PS
    import ns_gym_b200.update_functions as PU
    from ns_gym_b200.vector_env import NSVectorEnv

    case, n, m = CASES["c4_cartpole_rows"], 128, 4
    env = NSVectorEnv.heterogeneous(case["env_id"], [case["params_of"](PS, PU, e) for e in range(n)], bucket=True,
                                    precision="fp32", seed=3, change_notification=True,
                                    delta_change_notification=True, in_sim_change=True)
    env.reset(seed=3)
    a = torch.zeros(n, dtype=torch.int32, device=env.device)
    for _ in range(5):
        env.step(a)
    plan = env.get_planning_env(fanout=m)
    assert torch.equal(plan.buffers["theta"], env.buffers["theta"].repeat_interleave(m, dim=1))
    assert np.array_equal(plan.rows, np.repeat(env.rows, m, axis=0))
    ret, length = plan.rollout(40, gamma=0.95)
    assert bool(torch.isfinite(ret).all()) and int(length.min()) >= 1
    # the copies keep evolving (in_sim_change): some parameter moved away from its root's value
    assert not torch.equal(plan.buffers["theta"], env.buffers["theta"].repeat_interleave(m, dim=1))


def _linear_actions(w, obs, box):
    """Host restatement of the device-side linear policy: float32 multiply-adds in index order."""
    n, n_obs = obs.shape
    wn = np.broadcast_to(w, (n,) + w.shape[-2:]).astype(np.float32)
    v = wn[:, :, n_obs].copy()
    for q in range(n_obs):
        v = (v + (wn[:, :, q] * obs[:, None, q]).astype(np.float32)).astype(np.float32)
    return v[:, 0] if box else np.argmax(v, axis=1).astype(np.int32)


@pytest.mark.parametrize("specialize", [0, 1])
@pytest.mark.parametrize("name,precision,box,per_env", [
    ("c1_cartpole_readme", "fp32", False, False), ("c1_cartpole_readme", "fp64", False, True),
    ("c3_acrobot", "fp32", False, True), ("c3_mountaincar", "fp64", False, False),
    ("c3_pendulum", "fp32", True, True), ("mountaincar_continuous", "fp64", True, False)])
def test_linear_policy_rollout_equals_single_steps(name, precision, box, per_env, specialize):
    """nsgym_rollout_linear: K fused steps under a device-side linear policy on the float32
    observation = K launches with the actions a host restatement of that policy picks."""
    import torch

    case, n, K, seed = CASES[name], 2048, 33, 4321
    a = _make(case, n, precision, seed, specialize=specialize)
    b = _make(case, n, precision, seed, specialize=specialize)
    r = np.random.default_rng([5, len(name), int(per_env)])
    shape = a.policy_shape(per_env)
    w = r.normal(0, 1, shape).astype(np.float32)
    if box:
        w *= np.float32(0.5)
    wt = torch.as_tensor(w, device=a.device)
    ret, length = a.rollout(K, gamma=1.0, policy=wt)
    assert a.last_kernel_specialized == bool(specialize)
    acc = torch.zeros(n, dtype=torch.float32, device=b.device)
    for _ in range(K):
        obs = b.observation().float().cpu().numpy()
        act = _linear_actions(w, obs, box)
        dtype = b.real if box else torch.int32
        _, rew, _, _, _ = b.step(torch.as_tensor(act, device=b.device).to(dtype))
        acc += rew
    torch.cuda.synchronize()
    for key in ("state", "theta", "t"):
        assert torch.equal(a.buffers[key], b.buffers[key]), f"{name}: {key} differs after the linear-policy rollout"
    assert torch.equal(ret, acc)


@pytest.mark.parametrize("specialize", [0, 1])
@pytest.mark.parametrize("name,per_env", [("c5_bridge_uniform", False), ("c2_frozenlake8_drift", True),
                                          ("cliff_terminal", False)])
def test_tabular_policy_rollout_equals_single_steps(name, per_env, specialize):
    """Gridworlds: the linear policy on the one-hot cell is an action table."""
    import torch

    case, n, K, seed = CASES[name], 2048, 41, 99
    a = _make(case, n, "fp64", seed, specialize=specialize)
    b = _make(case, n, "fp64", seed, specialize=specialize)
    r = np.random.default_rng([6, len(name)])
    table = r.integers(0, 4, a.policy_shape(per_env)).astype(np.uint8)
    ret, length = a.rollout(K, gamma=1.0, policy=torch.as_tensor(table, device=a.device))
    assert a.last_kernel_specialized == bool(specialize)
    acc = torch.zeros(n, dtype=torch.float32, device=b.device)
    rows = np.arange(n)
    for _ in range(K):
        cell = b.buffers["state"].reshape(-1).cpu().numpy()
        act = (table[rows, cell] if per_env else table[cell]).astype(np.int32)
        _, rew, _, _, _ = b.step(torch.as_tensor(act, device=b.device))
        acc += rew
    torch.cuda.synchronize()
    for key in ("state", "theta", "t", "istate"):
        x, y = a.buffers[key], b.buffers[key]
        if x is not None:
            assert torch.equal(x, y), f"{name}: {key} differs after the tabular-policy rollout"
    assert torch.equal(ret, acc)


def test_policy_argument_is_validated():
    import torch

    env = _make(CASES["c1_cartpole_readme"], 64, "fp32", 1)
    assert env.policy_shape() == (2, 5) and env.policy_shape(True) == (64, 2, 5)
    with pytest.raises(ValueError):
        env.rollout(4, policy=torch.zeros(3, 5, device=env.device))
    with pytest.raises(ValueError):
        env.rollout(4, policy=torch.zeros(2, 5, device=env.device, dtype=torch.float64))
