"""CPU: the oracle port against the reference's own known-answer unit tests (transcribed in
tests/known_answers.py) and against scipy for the W1 closed form."""
import numpy as np
import pytest

import ns_gym_b200.schedulers as PS
import ns_gym_b200.update_functions as PU
from oracle import ns_port
from tests.known_answers import DIST_KA, SCHED_KA, UPDATE_KA


def _run(builder, calls, chain):
    d = ns_port.describe(builder(PS, PU))
    st = ns_port.SlotState(d)
    out, cur = [], None
    for param, t in calls:
        if cur is not None and chain:
            param = cur
        new, flag, delta = ns_port.call_update(d, st, param, t, None, 0)
        out.append((new, flag, delta))
        cur = new
    return out


@pytest.mark.parametrize("ka", UPDATE_KA, ids=[k[0] for k in UPDATE_KA])
def test_scalar_update_known_answers(ka):
    _, builder, calls, want, chain, _ref = ka
    got = _run(builder, calls, chain)
    prev = None
    for (new, flag, delta), (w_new, w_flag), (param, _t) in zip(got, want, calls):
        assert flag == w_flag
        assert np.isclose(new, w_new, rtol=1e-12, atol=1e-12)
        src = prev if (chain and prev is not None) else param
        if flag:
            assert np.isclose(delta, new - src)
        else:
            assert delta == 0.0
        prev = new


@pytest.mark.parametrize("ka", DIST_KA, ids=[k[0] for k in DIST_KA])
def test_distribution_update_known_answers(ka):
    _, builder, calls, want, chain, _ref = ka
    got = _run(builder, calls, chain)
    for (new, flag, _delta), (w_new, w_flag) in zip(got, want):
        assert flag == w_flag
        assert np.allclose(new, w_new)


@pytest.mark.parametrize("ka", SCHED_KA, ids=[k[0] for k in SCHED_KA])
def test_scheduler_known_answers(ka):
    _, builder, times, want, _ref = ka
    fn = PU.NoUpdate(builder(PS))
    d = ns_port.describe(fn)
    st = ns_port.SlotState(d)
    got = [ns_port.sched_fires(d["sched"], st, t, None, 0) for t in times]
    assert got == want
    assert all(isinstance(x, bool) for x in got)          # test_schedulers.py:595-706


def test_w1_closed_form_matches_scipy():
    from scipy.stats import wasserstein_distance

    rng = np.random.default_rng(0)
    for n in (2, 3, 4):
        idx = np.arange(n, dtype=float)
        for _ in range(200):
            u, v = rng.random(n), rng.random(n)
            if rng.random() < 0.3:
                u[rng.integers(n)] = 0.0
            want = float(wasserstein_distance(idx, idx, u_weights=u, v_weights=v))
            assert ns_port.w1_index_distance(list(u), list(v)) == want
    assert ns_port.w1_index_distance([1, 0, 0], [0, 0, 1]) == 2.0   # utils.py:66-72
    with pytest.raises(ValueError):
        ns_port.w1_index_distance([1.1, -0.05, -0.05], [1, 0, 0])   # test_gridworld_wrappers.py:192-199


def test_bridge_transitions_known_answers():
    """tests/test_bridge.py:135-224 -- action->cell map, out of bounds, forced slip, split mode."""
    b = ns_port.BridgePort()
    from oracle.streams import Clock, EnvStreams

    clock = Clock()
    b.streams = EnvStreams(np.array([[0.5]]), np.zeros((1, 1)), clock)
    expect = {0: 19, 1: 28, 2: 21, 3: 12}                          # LEFT DOWN RIGHT UP from (2, 4)
    for a, cell in expect.items():
        b.reset()
        b.P = [1.0, 0.0, 0.0]
        s, r, done, trunc, info = b.step(a)
        assert s == cell and trunc is False
    b.reset(); b.s = 0; b.P = [1.0, 0.0, 0.0]
    assert b.step(3)[0] == 0 and b.step(0)[0] == 0                 # out of bounds -> stay
    b.reset(); b.P = [0.0, 1.0, 0.0]                               # forced slip to (a + 1) % 4
    assert b.step(0)[0] == 28                                      # LEFT -> DOWN
    b.reset(); b.split_probs = True                                # (2, 4): col 4 >= ncol // 2 -> right side
    b.P_left, b.P_right = [0.0, 1.0, 0.0], [1.0, 0.0, 0.0]
    assert b.step(2)[0] == 21
    b.s = 19                                                       # col 3 -> left side: RIGHT slips to UP
    assert b.step(2)[0] == 11
    b.reset(); b.split_probs = False; b.P = [1.0, 0.0, 0.0]
    s, r, done, _, _ = b.step(3)                                   # (1, 4) is F
    assert (s, r, done) == (12, 0, False)
    s, r, done, _, _ = b.step(3)                                   # (0, 4) is H
    assert (s, r, done) == (4, -1, True)


def test_numpy_philox_known_answer_vectors():
    """Random123 kat_vectors for philox4x32_10 (tests/philox_np.py restates the device stream)."""
    from tests import philox_np as p

    z = np.uint32([0])
    assert [int(x[0]) for x in p.philox4x32_10(z, z, z, z, 0)] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = np.uint32([0xffffffff])
    assert [int(x[0]) for x in p.philox4x32_10(f, f, f, f, 0xffffffffffffffff)] == [
        0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
