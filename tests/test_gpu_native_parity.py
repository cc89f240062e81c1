"""GPU: the kernels THROUGHPUT runs launch -- native Philox draws, lean instantiations -- against the
oracle directly.

The parity tests proper (test_gpu_parity.py) inject numpy-drawn tables into both sides, which selects
the general, injection-capable kernel instantiations.  Here nothing is injected on the GPU: the batch
runs exactly as bench.py runs it, and the ORACLE is fed the numbers the kernels draw natively
(tests/philox_np.py restates the Philox counter layout and the draw transforms; test_gpu_native_draws.py
holds that restatement to the device value by value).  Every case asserts which instantiation ran.

Tolerances: fp64 mode 1e-9 relative (north_star); integer gridworld states, flags, fire indices and
change masks bit-exact; fp32 fast mode as stated in test_gpu_fp32.py, over each env's first episode.
"""
import numpy as np
import pytest

from tests import parity_util as pu
from tests import philox_np as PN
from tests.cases import CASES

pytestmark = pytest.mark.gpu

SEED = 77
N_ENVS = 64


def _run(name, precision, steps, offset=0, specialize=0, **extra):
    from ns_gym_b200 import native as nv

    case = CASES[name]
    grid = any(k in case["env_id"] for k in ("FrozenLake", "Cliff", "Bridge"))
    persistent = bool(case.get("wrapper", {}).get("persistent_params", False))
    actions = pu.harness.draw_actions(case, 5, steps, N_ENVS)
    clock, per_env, _, _ = PN.native_streams(N_ENVS, steps + 1, pu.n_slots_of(case), SEED, precision, grid,
                                             replay=not persistent, gid_offset=offset)
    ref = pu.oracle_trace_streams(case, clock, per_env, actions)
    env = pu.gpu_env(case, N_ENVS, precision=precision, seed=SEED, env_id_offset=offset, **extra)
    env.set_option("specialize", specialize)      # 1: the program-specialised kernels large batches run by default
    got = pu.gpu_run(env, actions, None, None)
    return ref, got, nv


LEAN_FP64 = [("c1_cartpole_readme", "LEAN_FAST"), ("cartpole_silent", "LEAN_FAST"), ("cartpole_persistent", "GENERAL"),
             ("c3_acrobot", "LEAN_FAST"), ("c3_mountaincar", "LEAN_MEDIUM"), ("c3_pendulum", "LEAN_MEDIUM"),
             ("mountaincar_continuous", "LEAN_FAST"), ("mountaincar_constraint", "LEAN_FAST"),
             ("pendulum_all", "LEAN_MEDIUM"), ("cartpole_stochastic", "GENERAL"), ("cartpole_stochastic_scheds", "GENERAL"),
             ("c2_frozenlake8_stepchange", "LEAN_FAST"), ("c2_frozenlake8_drift", "LEAN_FAST"), ("cliff_drift", "LEAN_FAST"),
             ("frozenlake5_multi_start", "LEAN_FAST"),
             ("c5_bridge_uniform", "LEAN_FAST"), ("c5_bridge_split", "LEAN_FAST"),
             ("frozenlake4_random_categorical", "GENERAL"), ("bridge_lipschitz_bounded", "GENERAL"),
             ("c4_cartpole_rows", "ROWS_LEAN"), ("het_cartpole_lean", "ROWS_LEAN"), ("c4_frozenlake8_rows", "ROWS_LEAN")]


# specialize = 1: the kernels bench.py's large batches actually launch (compiled at run time around the program);
# every kind of native-draw program specialises except batches with general (non-lean) per-env rows
@pytest.mark.parametrize("specialize", [0, 1])
@pytest.mark.parametrize("name,klass", LEAN_FP64)
def test_native_draw_kernels_match_the_oracle_fp64(name, klass, specialize):
    steps = min(CASES[name]["steps"], 120)
    ref, got, nv = _run(name, "fp64", steps, offset=(1 << 32) + 12345, specialize=specialize)
    assert got["_kernel_class"] == getattr(nv, "KERNEL_" + klass), (name, got["_kernel_class"])
    assert got["_specialized"] == bool(specialize), name
    assert not got["_bad_dist"]
    pu.compare(ref, got, float_obs_rtol=1e-6, name=name)


FP32 = [("c1_cartpole_readme", "LEAN_FAST"), ("c3_acrobot", "LEAN_FAST"), ("c3_mountaincar", "LEAN_MEDIUM"),
        ("c3_pendulum", "LEAN_MEDIUM"), ("mountaincar_continuous", "LEAN_FAST"), ("het_cartpole_lean", "ROWS_LEAN"),
        ("cartpole_stochastic", "GENERAL")]


@pytest.mark.parametrize("specialize", [0, 1])
@pytest.mark.parametrize("name,klass", FP32)
def test_native_draw_kernels_track_the_oracle_fp32(name, klass, specialize):
    """The benched fp32 kernels (C1 headline: classic_step_kernel<float, CartPole, 2, 0>) on their own
    draws vs the fp64 oracle fed the same draws, each env up to its first episode end."""
    steps = min(CASES[name]["steps"], 60)
    ref, got, nv = _run(name, "fp32", steps, specialize=specialize)
    assert got["_kernel_class"] == getattr(nv, "KERNEL_" + klass), (name, got["_kernel_class"])
    assert got["_specialized"] == bool(specialize), name
    ended = (ref["terminated"] | ref["truncated"] | ref["was_reset"])
    alive = np.cumsum(ended, axis=0) == 0
    assert alive[:5].all() and alive.sum() > 10 * N_ENVS
    np.testing.assert_allclose(got["raw0"], ref["raw0"], rtol=1e-6, atol=1e-7)
    for k in range(steps):
        m = alive[k]
        if not m.any():
            continue
        assert np.array_equal(ref["gt_change"][k][m], got["gt_change"][k][m]), f"{name}: fire flags differ at step {k}"
        np.testing.assert_allclose(got["theta"][k][m], ref["theta"][k][m], rtol=2e-5, atol=2e-5,
                                   err_msg=f"{name}: theta step {k}")
        np.testing.assert_allclose(got["raw"][k][m], ref["raw"][k][m], rtol=2e-3, atol=2e-3,
                                   err_msg=f"{name}: state step {k}")
        np.testing.assert_allclose(got["reward"][k][m], ref["reward"][k][m], rtol=2e-3, atol=2e-3)
    # episode boundaries: a threshold crossing may move by a step on a last-bit difference, no more
    first_ref = np.argmax(ended, axis=0)
    got_ended = (got["terminated"] | got["truncated"] | got["was_reset"])
    first_got = np.argmax(got_ended, axis=0)
    both = ended.any(0) & got_ended.any(0)
    assert np.array_equal(ended.any(0), got_ended.any(0)) or (ended.any(0) != got_ended.any(0)).mean() < 0.05
    if both.any():
        assert (np.abs(first_ref[both] - first_got[both]) <= 1).mean() > 0.95


def test_stochastic_schedulers_replay_their_pattern_every_episode():
    """NSWrapper.reset re-clones the update functions from the init-time template and never
    reseeds fn.scheduler.rng (base.py:381-395), so Random / Decaying / Memoryless schedulers fire the SAME
    pattern in every episode of an env.  Native draws: keyed by (env, episode time)."""
    import torch

    case = dict(CASES["cartpole_stochastic_scheds"])
    n, T = 512, 24
    env = pu.gpu_env(case, n, precision="fp64", autoreset="none", seed=5, max_episode_steps=T)
    pats = []
    for ep in range(3):
        env.reset()
        rows = []
        for k in range(T):
            env.step_raw(torch.zeros(n, dtype=torch.int32, device=env.device))
            rows.append(env.buffers["change"].clone())
        pats.append(torch.stack(rows).cpu().numpy())
    assert np.array_equal(pats[0], pats[1]) and np.array_equal(pats[1], pats[2])
    assert pats[0].any() and not pats[0].all()
    assert len({pats[0][:, i].tobytes() for i in range(n)}) > n // 2        # envs differ from each other
    # persistent_params keeps the scheduler objects across resets: the stream runs on
    case["wrapper"] = dict(case["wrapper"], persistent_params=True)
    env = pu.gpu_env(case, n, precision="fp64", autoreset="none", seed=5, max_episode_steps=T)
    pats = []
    for ep in range(2):
        env.reset()
        rows = []
        for k in range(T):
            env.step_raw(torch.zeros(n, dtype=torch.int32, device=env.device))
            rows.append(env.buffers["change"].clone())
        pats.append(torch.stack(rows).cpu().numpy())
    assert not np.array_equal(pats[0], pats[1])


def test_seeded_reset_is_reproducible():
    """reset(seed=s) reseeds every generator from s (base.py:386-388, 412-421): the same seed replays
    the same episode whatever ran before; another seed does not."""
    import torch

    env = pu.gpu_env(CASES["cartpole_stochastic"], 256, precision="fp32", seed=1)
    a = torch.randint(0, 2, (256,), device=env.device, dtype=torch.int32)

    def episode(seed, pre=0):
        for _ in range(pre):
            env.step_raw(a)
        env.reset(seed=seed)
        out = [env.buffers["state"].clone()]
        for _ in range(12):
            env.step_raw(a)
            out += [env.buffers["state"].clone(), env.buffers["theta"].clone()]
        return out

    env.reset(seed=1)
    x, y, z = episode(42), episode(42, pre=7), episode(43)
    assert all(torch.equal(p, q) for p, q in zip(x, y))
    assert not all(torch.equal(p, q) for p, q in zip(x, z))
