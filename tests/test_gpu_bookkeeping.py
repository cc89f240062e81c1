"""GPU: host-side bookkeeping around the step kernels -- stream ordering of the host-buffer step,
snapshot / restore of planning copies, planning-copy streams, the fused episode-statistics kernel,
the constraint-rejection flag."""
import ctypes as C
import warnings

import numpy as np
import pytest

from tests import parity_util as pu
from tests.cases import CASES

pytestmark = pytest.mark.gpu


def test_host_step_is_ordered_after_the_callers_stream():
    """reset / device step / host step back to back with no synchronisation in between: the host
    pipeline runs on the handle's own non-blocking streams and must wait for the work the caller
    queued (a 2^22-env reset takes long enough to lose the race otherwise)."""
    import torch

    case, n = CASES["c1_cartpole_readme"], 1 << 22
    a_env, b_env = pu.gpu_env(case, n, precision="fp32"), pu.gpu_env(case, n, precision="fp32")
    acts = a_env.action_space.sample()
    h_act, h_out = b_env.make_host_io()
    h_act.copy_(acts.cpu())
    torch.cuda.synchronize()
    for rep in range(3):
        a_env.reset(seed=rep)
        torch.cuda.synchronize()
        a_env.step_raw(acts)
        torch.cuda.synchronize()
        a_env.step_raw(acts)
        torch.cuda.synchronize()
        b_env.reset(seed=rep)                       # no synchronize from here on
        b_env.step_raw(acts)
        b_env.step_host(h_act, h_out, n_chunks=8)
        torch.cuda.synchronize()
        for key in ("state", "theta", "t", "reward", "flags"):
            assert torch.equal(a_env.buffers[key], b_env.buffers[key]), (rep, key)
        assert torch.equal(h_out["obs"], a_env.buffers["obs"].cpu())


@pytest.mark.parametrize("name", ["c1_cartpole_readme", "c5_bridge_uniform"])
def test_snapshot_restore_rewinds_a_planning_copy(name):
    """snapshot -> rollout(K) -> restore on a planning copy, many times: the copy's TimeLimit count
    (steps since the copy) rewinds with the state, so every pass returns the same thing."""
    import torch

    case = CASES[name]
    root = pu.gpu_env(case, 128, precision="fp64")
    root.reset(seed=3)
    a = root.action_space.sample()
    for _ in range(3):
        root.step_raw(a)
    plan = root.get_planning_env(fanout=8, seed=99)
    limit = int(plan.program.spec.max_episode_steps)
    K = max(limit // 3, 8)
    snap = plan.snapshot()
    first = None
    for rep in range(8):                            # 8 * K steps > the TimeLimit: a drifting counter truncates early
        ret, length = plan.rollout(K, gamma=0.97)
        torch.cuda.synchronize()
        if first is None:
            first = (ret.clone(), length.clone(), plan.buffers["state"].clone())
        else:
            assert torch.equal(ret, first[0]) and torch.equal(length, first[1]), rep
            assert torch.equal(plan.buffers["state"], first[2])
        plan.restore(snap)
    # stepping after a restore counts the limit from the snapshot point
    for k in range(limit):
        plan.step_raw(torch.zeros(plan.num_envs, dtype=plan.buffers["action"].dtype, device=plan.device))
    torch.cuda.synchronize()
    assert bool(((plan.buffers["flags"] & 2) != 0).all())         # truncated exactly at the limit
    plan.restore(snap)
    plan.step_raw(torch.zeros(plan.num_envs, dtype=plan.buffers["action"].dtype, device=plan.device))
    torch.cuda.synchronize()
    assert not bool(((plan.buffers["flags"] & 2) != 0).any())


def test_planning_copies_draw_their_own_streams():
    """_reseed_planning_env_rngs (base.py:433-441): every copy gets fresh generators.  Two copies of the
    same root step roll out differently; `seed=` pins a copy; copies of different shards of one job
    use disjoint global env ids."""
    import torch

    case = CASES["c1_cartpole_readme"]
    root = pu.gpu_env(case, 256, precision="fp32", seed=11, env_id_offset=512)
    root.reset()
    p1, p2 = root.get_planning_env(fanout=4), root.get_planning_env(fanout=4)
    r1, _ = p1.rollout(40)
    r2, _ = p2.rollout(40)
    assert not torch.equal(r1, r2)
    q1, q2 = root.get_planning_env(fanout=4, seed=5), root.get_planning_env(fanout=4, seed=5)
    s1, _ = q1.rollout(40)
    s2, _ = q2.rollout(40)
    assert torch.equal(s1, s2)
    assert int(p1.program.spec.env_id_offset) == 512 * 4
    # shard layout invariance of the copies: lanes [1024, 2048) of a 512-root job == the copies of shard 1
    whole = pu.gpu_env(case, 512, precision="fp32", seed=11, env_id_offset=256)
    whole.reset()
    w = whole.get_planning_env(fanout=4, seed=5)
    ws, _ = w.rollout(40)
    assert torch.equal(ws[1024:], s1)


def test_episode_stats_kernel_matches_a_numpy_restatement():
    import torch

    from ns_gym_b200 import native as nv
    from ns_gym_b200.distributed import EpisodeStats

    case, n = CASES["cartpole_constraint"], 4099
    env = pu.gpu_env(case, n, precision="fp32", seed=2)
    env.reset()
    stats = EpisodeStats(env)
    run_ret, run_len, tot = np.zeros(n), np.zeros(n, dtype=np.int64), np.zeros(len(nv.STAT_KEYS))
    for k in range(80):
        env.step_raw(env.action_space.sample())
        stats.update()
        f = env.buffers["flags"].cpu().numpy()
        r = env.buffers["reward"].cpu().numpy().astype(np.float64)
        stepped, ended = (f & 4) == 0, (f & 3) != 0
        run_ret += np.where(stepped, r, 0.0)
        run_len += stepped
        tot += [stepped.sum(), ended.sum(), run_ret[ended].sum(), run_len[ended].sum(), ((f & 1) != 0).sum(),
                ((f & 2) != 0).sum(), ((f & 8) != 0).sum(), ((f & 128) != 0).sum()]
        run_ret[ended] = 0
        run_len[ended] = 0
    torch.cuda.synchronize()
    np.testing.assert_allclose(stats.totals.cpu().numpy(), tot, rtol=1e-12)
    np.testing.assert_allclose(stats.running_return.cpu().numpy(), run_ret, rtol=1e-12)
    assert np.array_equal(stats.running_length.cpu().numpy(), run_len)
    red = stats.reduce()
    assert red["episodes"] > 0 and red["rejected_updates"] > 0
    assert red["mean_length"] == pytest.approx(tot[3] / tot[1])


def test_constraint_rejections_are_flagged():
    """classic_control.py:87-92: a fired update that violates a constraint keeps the old value and the
    reference warns (ConstraintViolationWarning).  masspole 0.1 - 0.03 (t + 1) is rejected from t = 3 on."""
    import torch

    from ns_gym_b200 import native as nv
    from ns_gym_b200.wrappers import ConstraintViolationWarning

    env = pu.gpu_env(CASES["cartpole_constraint"], 64, precision="fp64", autoreset="none")
    env.reset(seed=0)
    counts = []
    for k in range(6):
        env.step_raw(torch.zeros(64, dtype=torch.int32, device=env.device))
        counts.append(env.constraint_violations())
    # t = 0: gravity 9.8 - 4 ok, length <- -1 rejected (StepWise list [-1, 0.6, 0, 0.7] at t = 0, 3): every env
    assert counts[0] == 64 and counts[1] == 0 and counts[3] == 64
    with pytest.warns(ConstraintViolationWarning):
        env.constraint_violations(warn=True)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        env2 = pu.gpu_env(CASES["cartpole_silent"], 64, precision="fp64")
        env2.reset(seed=0)
        env2.step_raw(torch.zeros(64, dtype=torch.int32, device=env2.device))
        assert env2.constraint_violations(warn=True) == 0
    assert nv.FLAG_REJECTED == 8


def test_host_memory_helpers_and_error_paths():
    import torch

    from ns_gym_b200 import native as nv

    lib = nv.load()
    p = lib.nsgym_alloc_host(1 << 20, 1)
    assert p
    C.memset(p, 7, 1 << 20)
    lib.nsgym_free_host(p)
    env = pu.gpu_env(CASES["cartpole_silent"], 300, precision="fp32")
    h_act, h_out = env.make_host_io()
    with pytest.raises(nv.NsgymError):                 # step before reset
        env.step_host(h_act, h_out)
    env.reset(seed=0)
    env.step_host(h_act, h_out, n_chunks=1000)         # more chunks than 256-env blocks: clamps
    torch.cuda.synchronize()
    assert torch.equal(h_out["obs"], env.buffers["obs"].cpu())


def test_mixed_batch_steps_all_shards_in_one_call():
    """MixedVectorEnv.step_raw (nsgym_step_many) == stepping every shard on its own."""
    import torch

    from ns_gym_b200.vector_env import MixedVectorEnv

    names = ("c4_cartpole_rows", "c4_frozenlake8_rows")
    one = [pu.gpu_env(CASES[n], 300, precision="fp32", seed=4) for n in names]
    many = [pu.gpu_env(CASES[n], 300, precision="fp32", seed=4) for n in names]
    mixed = MixedVectorEnv(many)
    for e in one:
        e.reset()
    mixed.reset()
    for k in range(25):
        acts = [e.action_space.sample() for e in one]
        for e, a in zip(one, acts):
            e.step_raw(a)
        mixed.step_raw(acts)
    torch.cuda.synchronize()
    for a, b in zip(one, many):
        for key in ("state", "theta", "t", "reward", "flags", "change"):
            assert torch.equal(a.buffers[key], b.buffers[key]), key
    assert mixed.launch_count == sum(e.launch_count for e in one)
