"""CPU: host logic of the compiler (tunable_params -> opcode table)."""
import numpy as np
import pytest

import ns_gym_b200.schedulers as S
import ns_gym_b200.update_functions as U
from ns_gym_b200 import native as nv
from ns_gym_b200.compile import CompileError, compile_program
from tests.cases import CASES


@pytest.mark.parametrize("name", sorted(CASES))
def test_every_case_compiles(name):
    c = CASES[name]
    p = compile_program(c["env_id"], c["params"](S, U), 16, precision="fp64", **c["wrapper"], **c["make"])
    assert p.spec.n_slots == len(p.keys) == len(c["params"](S, U))
    assert [p.spec.slots[j].theta_index for j in range(p.spec.n_slots)] == sorted(
        {p.spec.slots[j].theta_index for j in range(p.spec.n_slots)}, key=lambda i: [
            p.spec.slots[j].theta_index for j in range(p.spec.n_slots)].index(i))


def test_slot_order_is_dict_order():
    tp = {"length": U.NoUpdate(S.ContinuousScheduler()), "gravity": U.NoUpdate(S.ContinuousScheduler())}
    p = compile_program("CartPole-v1", tp, 4)
    assert p.keys == ["length", "gravity"]
    assert [p.spec.slots[j].theta_index for j in range(2)] == [5, 0]


def test_scheduler_lowering():
    tp = {"gravity": U.IncrementUpdate(S.PeriodicScheduler(3, start=2.5, end=9.9), k=1.0)}
    s = compile_program("CartPole-v1", tp, 4).spec.slots[0]
    assert (s.sched_op, s.si[0], s.start, s.end) == (nv.SCHED_PERIODIC, 3, 3, 9)
    tp = {"gravity": U.IncrementUpdate(S.ContinuousScheduler(), k=1.0)}
    s = compile_program("CartPole-v1", tp, 4).spec.slots[0]
    assert (s.start, s.end) == (0, nv.INT32_MAX)
    tp = {"gravity": U.IncrementUpdate(S.DiscreteScheduler({1, 33}), k=1.0)}
    p = compile_program("CartPole-v1", tp, 4)
    s = p.spec.slots[0]
    assert s.sched_op == nv.SCHED_BITMAP and s.si[1] == 34
    words = [p.spec.bitmap[i] for i in range(p.spec.n_bitmap_words)]
    assert words == [2, 2]
    # one run of event times = a Continuous scheduler on that range (fast class, no bitmap)
    for events, kw, want in (({12}, {}, (12, 12)), ({4, 5, 6}, {}, (4, 6)), ({3, 4, 5}, dict(start=2, end=8), (3, 5))):
        s = compile_program("CartPole-v1", {"gravity": U.IncrementUpdate(S.DiscreteScheduler(events, **kw), k=1.0)},
                            4).spec.slots[0]
        assert (s.sched_op, s.start, s.end) == (nv.SCHED_CONTINUOUS, *want)
    s = compile_program("CartPole-v1", {"gravity": U.IncrementUpdate(S.WindowScheduler([(3, 7)], start=5), k=1.0)},
                        4).spec.slots[0]
    assert (s.sched_op, s.start, s.end) == (nv.SCHED_CONTINUOUS, 5, 7)
    tp = {"gravity": U.IncrementUpdate(S.CustomScheduler(lambda t: t in (0, 499, 500)), k=1.0)}
    # the bitmap covers every reachable t: an episode (TimeLimit 500) plus a planning copy stepping on
    # for another limit of its own (<= 1000): 2 * 500 + 500 + 2, one bit more for t = horizon
    p = compile_program("CartPole-v1", tp, 4)
    assert p.spec.slots[0].si[1] == 1503
    # no TimeLimit to bound t (autoreset "none" may step past the end): the documented custom_horizon
    p = compile_program("CartPole-v1", tp, 4, autoreset="none")
    assert p.spec.slots[0].si[1] == 65537
    # a Memoryless scheduler may drive a list update: next-fire time and cursor share the slot's word
    tp = {"gravity": U.StepWiseUpdate(S.MemorylessScheduler(p=0.3, seed=3), [9.0, 10.0])}
    s = compile_program("CartPole-v1", tp, 4).spec.slots[0]
    assert (s.sched_op, s.upd_op, s.ui[3], s.istate_plane) == (nv.SCHED_MEMORYLESS, nv.UPD_STEPWISE, 1, 0)
    with pytest.raises(CompileError):
        compile_program("CartPole-v1", {"gravity": U.CyclicUpdate(S.MemorylessScheduler(p=0.3), list(range(200)))}, 4)
    tp = {"gravity": U.IncrementUpdate(S.MemorylessScheduler(p=0.3, seed=3), k=1.0)}
    s = compile_program("CartPole-v1", tp, 4).spec.slots[0]
    assert s.sched_op == nv.SCHED_MEMORYLESS and s.istate_plane == 0
    assert s.istate_init == int(np.random.default_rng(3).geometric(p=0.3, size=(1,))[0])


def test_update_lowering_coefficients():
    tp = {"gravity": U.DecrementUpdate(S.ContinuousScheduler(), k=0.25),
          "length": U.LinearInterpolation(S.ContinuousScheduler(), 0.5, 0.8, T=25),
          "tau": U.SigmoidTransition(S.ContinuousScheduler(), a=0.02, b=0.03, k=0.5, t0=10)}
    sp = compile_program("CartPole-v1", tp, 4).spec
    assert (sp.slots[0].upd_op, sp.slots[0].uf[0]) == (nv.UPD_ADD, -0.25)
    assert list(sp.slots[1].uf)[:3] == [0.5, 0.8 - 0.5, 25.0]
    assert list(sp.slots[2].uf)[:4] == [0.02, 0.03 - 0.02, 0.5, 10.0]
    assert sp.slots[0].constraint == nv.CONS_REJECT_LT0 and sp.slots[1].constraint == nv.CONS_REJECT_LE0
    assert sp.slots[2].constraint == nv.CONS_NONE


def test_acrobot_partner_slots():
    tp = {"LINK_COM_POS_1": U.NoUpdate(S.ContinuousScheduler()), "LINK_LENGTH_1": U.NoUpdate(S.ContinuousScheduler()),
          "LINK_COM_POS_2": U.NoUpdate(S.ContinuousScheduler())}
    sp = compile_program("Acrobot-v1", tp, 4).spec
    assert (sp.slots[0].constraint, sp.slots[0].partner_slot, sp.slots[0].partner_index) == (nv.CONS_ACRO_COM, 1, 1)
    assert (sp.slots[1].constraint, sp.slots[1].partner_slot, sp.slots[1].partner_index) == (nv.CONS_ACRO_LENGTH1, 0, 5)
    assert (sp.slots[2].partner_slot, sp.slots[2].partner_index) == (-1, 2)


def test_reference_assertions_are_kept():
    with pytest.raises(AssertionError):                               # base.py:257-261
        compile_program("CartPole-v1", {"not_a_param": U.NoUpdate(S.ContinuousScheduler())}, 4)
    with pytest.raises(AssertionError):                               # toy_text.py:329-334
        compile_program("FrozenLake-v1", {"P": U.DistributionNoUpdate(S.ContinuousScheduler())}, 4,
                        initial_prob_dist=[0.5, 0.2, 0.2])
    with pytest.raises(AssertionError):                               # base.py:114-119
        U.IncrementUpdate("not_a_scheduler", k=1.0)
    with pytest.raises(AssertionError):                               # schedulers.py:66-71
        S.DiscreteScheduler({1, 50}, start=0, end=10)


def test_rejections():
    fn = U.RandomWalk(S.ContinuousScheduler())
    with pytest.raises(CompileError):
        compile_program("CartPole-v1", {"gravity": fn, "length": fn}, 4)   # shared stateful object (S13)
    with pytest.raises(CompileError):     # prev_time would carry across resets
        compile_program("FrozenLake-v1", {"P": U.LCBoundedDistrubutionUpdate(S.ContinuousScheduler(), L=0.5)}, 4,
                        persistent_params=True)
    with pytest.raises(CompileError):     # only rules constructible from the scheduler alone can be inner rules
        compile_program("FrozenLake-v1", {"P": U.LCBoundedDistrubutionUpdate(
            S.ContinuousScheduler(), L=0.5, update_fn=U.UniformDrift)}, 4)
    p = compile_program("FrozenLake-v1", {"P": U.LCBoundedDistrubutionUpdate(S.ContinuousScheduler(), L=0.5)}, 4)
    assert p.spec.slots[0].upd_op == nv.UPD_D_RANDOM and p.spec.slots[0].ui[2] == 1 and p.spec.slots[0].uf[5] == 0.5
    with pytest.raises(CompileError):
        compile_program("CartPole-v1", {"gravity": U.UniformDrift(S.ContinuousScheduler(), 0.1)}, 4)
    with pytest.raises(CompileError):
        compile_program("Nope-v0", {}, 4)
    # stateless objects may be shared (the reference's own fixtures do it: test_step_reset.py:36-44)
    inc = U.IncrementUpdate(S.ContinuousScheduler(), k=0.1)
    assert compile_program("CartPole-v1", {"masspole": inc, "gravity": inc}, 4).spec.n_slots == 2


def test_gridworld_maps():
    p = compile_program("FrozenLake-v1", {"P": U.DistributionNoUpdate(S.ContinuousScheduler())}, 4, map_name="8x8")
    sp = p.spec
    assert (sp.nrow, sp.ncol, sp.start_cell, sp.n_dist) == (8, 8, 0, 3)
    assert sp.goal_mask == 1 << 63 and bin(sp.hole_mask).count("1") == 10
    p = compile_program("ns_gym/Bridge-v0", {"P_left": U.DistributionNoUpdate(S.ContinuousScheduler())}, 4,
                        initial_prob_dist=([0.9, 0.05, 0.05], [1, 0, 0]))
    sp = p.spec
    assert sp.split_mode == 1 and sp.start_cell == 20 and (sp.nrow, sp.ncol) == (5, 8)
    assert list(sp.theta_init[1])[:3] == [0.9, 0.05, 0.05] and list(sp.theta_init[2])[:3] == [1, 0, 0]
    assert bin(sp.goal_mask).count("1") == 2
    p = compile_program("CliffWalking-v1", {"P": U.DistributionNoUpdate(S.ContinuousScheduler())}, 4)
    assert p.spec.start_cell == 36 and p.spec.n_dist == 4 and p.spec.max_episode_steps == 0


def test_reference_objects_compile_too():
    """Duck typing: dictionaries built from the reference's own classes compile unchanged."""
    from oracle import ref_loader

    if not ref_loader.available():
        pytest.skip("reference tree not present")
    ref_loader.load()
    import ns_gym.schedulers as RS
    import ns_gym.update_functions as RU

    for name in ("c1_cartpole_readme", "cartpole_all_params", "c5_bridge_split", "cartpole_stochastic_scheds"):
        c = CASES[name]
        a = compile_program(c["env_id"], c["params"](RS, RU), 8, **c["wrapper"], **c["make"]).spec
        b = compile_program(c["env_id"], c["params"](S, U), 8, **c["wrapper"], **c["make"]).spec
        for j in range(a.n_slots):
            sa, sb = a.slots[j], b.slots[j]
            assert (sa.sched_op, sa.upd_op, sa.theta_index, sa.start, sa.end) == (
                sb.sched_op, sb.upd_op, sb.theta_index, sb.start, sb.end)
            assert list(sa.uf) == list(sb.uf) and list(sa.sf) == list(sb.sf)
            assert list(sa.si) == list(sb.si) and sa.istate_init == sb.istate_init


# ---- heterogeneous batches: compile_rows (BASELINE config C4) -----------------------------------
def test_compile_rows_layout_and_cursor_planes():
    from ns_gym_b200.compile import compile_rows, rows_dtype

    per_env = [
        {"masspole": U.IncrementUpdate(S.PeriodicScheduler(2 + e), k=0.1 * (e + 1)),
         "gravity": (U.StepWiseUpdate(S.ContinuousScheduler(), [1.0 + e, 2.0]) if e % 2
                     else U.RandomWalk(S.BurstScheduler(1, 3), sigma=0.1 * e))}
        for e in range(6)]
    prog, rows = compile_rows("CartPole-v1", per_env, precision="fp64")
    assert rows.shape == (6, 2) and rows.dtype == rows_dtype()
    assert rows.dtype.itemsize == nv.C.sizeof(nv.NsgymSlot)
    assert rows["sched_op"][:, 0].tolist() == [nv.SCHED_PERIODIC] * 6
    assert rows["si"][:, 0, 0].tolist() == [2, 3, 4, 5, 6, 7]
    assert np.allclose(rows["uf"][:, 0, 0], [0.1 * (e + 1) for e in range(6)])
    assert rows["upd_op"][:, 1].tolist() == [nv.UPD_RW, nv.UPD_STEPWISE] * 3
    # a cursor plane exists for slot 1 because SOME env needs one; rows that do not need it keep -1
    assert prog.spec.slots[1].istate_plane == 0 and prog.spec.slots[0].istate_plane == -1
    assert rows["istate_plane"][:, 1].tolist() == [-1, 0] * 3
    # every env's list lives in the shared pool at its own offset
    offs = rows["ui"][1::2, 1, 0].tolist()
    assert len(set(offs)) == 3 and prog.spec.n_pool_f >= 6
    # shared key set
    assert (rows["theta_index"] == rows["theta_index"][0]).all()
    assert (rows["constraint"] == rows["constraint"][0]).all()


def test_compile_rows_rejects_mismatched_key_sets():
    from ns_gym_b200.compile import compile_rows

    a = {"masspole": U.IncrementUpdate(S.ContinuousScheduler(), k=0.1)}
    b = {"gravity": U.IncrementUpdate(S.ContinuousScheduler(), k=0.1)}
    with pytest.raises(CompileError):
        compile_rows("CartPole-v1", [a, b])
    with pytest.raises(CompileError):
        compile_rows("CartPole-v1", [])


def test_heterogeneous_cases_compile_per_env():
    from ns_gym_b200.compile import compile_rows

    for name, c in CASES.items():
        if "params_of" not in c:
            continue
        prog, rows = compile_rows(c["env_id"], [c["params_of"](S, U, e) for e in range(12)],
                                  precision="fp64", **c["wrapper"], **c["make"])
        assert rows.shape == (12, prog.spec.n_slots)
        assert len({tuple(r) for r in rows["upd_op"].tolist()}) > 1, f"{name}: rows are all alike"


def test_type_mismatch_checker_matches_the_reference_helper():
    """ns_gym/utils.py:122-152."""
    from ns_gym_b200.base import Reward
    from ns_gym_b200.utils import type_mismatch_checker

    obs = {"state": [1, 2], "env_change": {}, "delta_change": {}, "relative_time": 3}
    rew = Reward(reward=1.5, env_change={}, delta_change={}, relative_time=3)
    assert type_mismatch_checker(obs, rew) == ([1, 2], 1.5)
    assert type_mismatch_checker([1, 2], 0.5) == ([1, 2], 0.5)
    assert type_mismatch_checker() == (None, None)
    with pytest.raises(AssertionError):
        type_mismatch_checker({"not_state": 1}, None)


def test_batch_size_limits():
    """empty and oversized batches fail at compile time, like nsgym_create does (status -1)"""
    from ns_gym_b200.compile import CompileError, compile_program, compile_rows

    for n in (0, -1, (1 << 28) + 1):
        with pytest.raises(CompileError):
            compile_program("CartPole-v1", {}, n)
    with pytest.raises(CompileError):
        compile_rows("CartPole-v1", [])
    assert compile_program("CartPole-v1", {}, 1).spec.n_envs == 1
    assert compile_program("CartPole-v1", {}, 1 << 28).spec.n_envs == 1 << 28
