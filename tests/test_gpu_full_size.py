"""GPU: BASELINE.json's full sizes, checked through size-independent properties (the oracle cannot
run millions of envs): shard invariance of the counter-based streams, replay determinism,
episode bookkeeping, probability mass, state ranges."""
import numpy as np
import pytest

from tests.cases import CASES

pytestmark = pytest.mark.gpu


def _env(case, n, precision, seed=11, offset=0, **kw):
    import ns_gym_b200.schedulers as PS
    import ns_gym_b200.update_functions as PU
    from ns_gym_b200.vector_env import NSVectorEnv

    env = NSVectorEnv(case["env_id"], case["params"](PS, PU), n, precision=precision, seed=seed,
                      env_id_offset=offset, **{**case["wrapper"], **case["make"], **kw})
    env.reset(seed=seed)
    return env


def _actions(env, seed):
    import torch

    g = torch.Generator(device=env.device)
    g.manual_seed(seed)
    return torch.randint(0, env.action_space_n, (env.num_envs,), generator=g, device=env.device, dtype=torch.int32)


def test_c1_cartpole_16m_envs_shard_invariance_and_bookkeeping():
    """C1 at 2^24 envs: one full batch == two half batches with global env ids (bit for bit), and the
    episode bookkeeping is consistent after 30 steps of native Philox draws."""
    import torch

    case, n, K = CASES["c1_cartpole_readme"], 1 << 24, 30
    full = _env(case, n, "fp32")
    a = _actions(full, 3)
    t_prev = full.relative_time().clone()
    ended_prev = torch.zeros(n, dtype=torch.bool, device=full.device)
    for k in range(K):
        obs, r, term, trunc, info = full.step(a)
        reset_now = info["was_reset"]
        assert torch.equal(reset_now, ended_prev), "an env resets exactly one call after it ended"
        t_now = full.relative_time()
        assert torch.equal(t_now, torch.where(reset_now, torch.zeros_like(t_now), t_prev + 1))
        assert bool((r[reset_now] == 0).all()) and bool((r[~reset_now] == 1).all())
        assert not bool((term & reset_now).any())
        # masspole = 0.1 + 0.1 t for every env (IncrementUpdate on a ContinuousScheduler)
        assert torch.allclose(full.theta()["masspole"], 0.1 + 0.1 * t_now.float(), rtol=1e-5, atol=1e-6)
        ended_prev, t_prev = (term | trunc), t_now.clone()
    assert bool(torch.isfinite(full.buffers["state"]).all())
    assert 0.0 < float(ended_prev.float().mean()) < 0.5
    state_full, theta_full = full.buffers["state"].clone(), full.buffers["theta"].clone()
    del full
    torch.cuda.empty_cache()
    half = n // 2
    for part in range(2):
        shard = _env(case, half, "fp32", offset=part * half)
        for k in range(K):
            shard.step_raw(a[part * half:(part + 1) * half])
        assert torch.equal(shard.buffers["state"], state_full[part * half:(part + 1) * half])
        assert torch.equal(shard.buffers["theta"], theta_full[:, part * half:(part + 1) * half])
        del shard
        torch.cuda.empty_cache()


def test_c2_frozenlake_1m_envs_properties():
    """C2 at 2^20 envs: cells stay on the 8x8 map, P stays a distribution, the step change at t = 12
    arrives in every env at the same time, replays are deterministic."""
    import torch

    case, n = CASES["c2_frozenlake8_stepchange"], 1 << 20
    runs = []
    for rep in range(2):
        env = _env(case, n, "fp64", autoreset="none")
        a = _actions(env, 5)
        for k in range(14):
            obs, r, term, trunc, info = env.step(a)
            p = env.transition_prob()["P"]
            assert torch.allclose(p.sum(0), torch.ones(n, dtype=torch.float64, device=env.device))
            want = [1.0, 0.0, 0.0] if k < 12 else [0.0, 0.5, 0.5]
            assert bool((p == torch.tensor(want, dtype=torch.float64, device=env.device)[:, None]).all())
            fired = info["Ground Truth Env Change"]["P"]
            assert int(fired.sum()) == (n if k == 12 else 0)
        s = env.buffers["state"]
        assert int(s.min()) >= 0 and int(s.max()) < 64
        runs.append((s.clone(), env.buffers["flags"].clone()))
        del env
    assert torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[0][1], runs[1][1])


def test_c5_bridge_rollout_16m_envs_matches_steps_on_a_slice():
    """C5: the fused K-step rollout over 2^24 envs; a 4096-env slice of it (same global ids) is
    reproduced bit for bit by a small handle."""
    import torch

    case, n, K = CASES["c5_bridge_uniform"], 1 << 24, 16
    big = _env(case, n, "fp64")
    ret, length = big.rollout(K)
    lo = 5 * 4096
    small = _env(case, 4096, "fp64", offset=lo)
    r2, l2 = small.rollout(K)
    assert torch.equal(ret[lo:lo + 4096], r2) and torch.equal(length[lo:lo + 4096], l2)
    assert torch.equal(big.buffers["state"][lo:lo + 4096], small.buffers["state"])
    assert int(big.buffers["state"].min()) >= 0 and int(big.buffers["state"].max()) < 40
    assert int(length.min()) >= 1 and int(length.max()) <= K
