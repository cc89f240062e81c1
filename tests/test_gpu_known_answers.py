"""GPU: the reference's known-answer unit tests (tests/known_answers.py) evaluated by the CUDA
interpreter's fire-test + advance stages alone (nsgym_eval_update), fp64 and fp32."""
import ctypes as C

import numpy as np
import pytest

from tests.known_answers import DIST_KA, SCHED_KA, UPDATE_KA

pytestmark = pytest.mark.gpu


def _evaluator(fn, precision, dist_len=0):
    import torch

    from ns_gym_b200 import native as nv
    from ns_gym_b200.compile import compile_program

    lib = nv.load()
    if dist_len:
        env_id = "CliffWalking-v1" if dist_len == 4 else "FrozenLake-v1"
        prog = compile_program(env_id, {"P": fn}, 1, initial_prob_dist=[1.0] + [0.0] * (dist_len - 1))
    else:
        prog = compile_program("CartPole-v1", {"force_mag": fn}, 1, precision=precision)
    h = C.c_void_p()
    nv.check(lib.nsgym_create(C.byref(prog.spec), C.byref(h)))
    dev = torch.device("cuda")
    real = torch.float64 if (precision == "fp64" or dist_len) else torch.float32
    ist = torch.full((1,), prog.spec.slots[0].istate_init, dtype=torch.int32, device=dev)

    def call(param, t):
        if dist_len:
            p = torch.tensor(param, dtype=real, device=dev).reshape(dist_len, 1).contiguous()
        else:
            p = torch.tensor([param], dtype=real, device=dev)
        tt = torch.tensor([t], dtype=torch.int32, device=dev)
        flag = torch.zeros(1, dtype=torch.uint8, device=dev)
        delta = torch.zeros(1, dtype=real, device=dev)
        nv.check(lib.nsgym_eval_update(h, 0, p.data_ptr(), tt.data_ptr(), ist.data_ptr(), flag.data_ptr(),
                                       delta.data_ptr(), None, None, 1, None))
        torch.cuda.synchronize()
        new = p.reshape(-1).double().cpu().tolist()
        return (new if dist_len else new[0]), int(flag.item()), float(delta.item())

    call.destroy = lambda: lib.nsgym_destroy(h)
    return call


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("ka", UPDATE_KA, ids=[k[0] for k in UPDATE_KA])
def test_scalar_update_known_answers(ka, precision):
    import ns_gym_b200.schedulers as PS
    import ns_gym_b200.update_functions as PU

    _, builder, calls, want, chain, _ref = ka
    call = _evaluator(builder(PS, PU), precision)
    tol = dict(rtol=1e-12, atol=1e-12) if precision == "fp64" else dict(rtol=2e-6, atol=1e-6)
    cur = None
    for (param, t), (w_new, w_flag) in zip(calls, want):
        if chain and cur is not None:
            param = cur
        new, flag, delta = call(param, t)
        assert flag == w_flag
        assert np.isclose(new, w_new, **tol)
        assert np.isclose(delta, (new - param) if flag else 0.0, **tol)
        cur = new
    call.destroy()


@pytest.mark.parametrize("ka", DIST_KA, ids=[k[0] for k in DIST_KA])
def test_distribution_update_known_answers(ka):
    import ns_gym_b200.schedulers as PS
    import ns_gym_b200.update_functions as PU
    from oracle.ns_port import w1_index_distance

    _, builder, calls, want, chain, _ref = ka
    call = _evaluator(builder(PS, PU), "fp64", dist_len=len(calls[0][0]))
    cur = None
    for (param, t), (w_new, w_flag) in zip(calls, want):
        if chain and cur is not None:
            param = cur
        new, flag, delta = call(param, t)
        assert flag == w_flag
        assert np.allclose(new, w_new, rtol=1e-12, atol=1e-15)
        if flag:
            assert delta == w1_index_distance(param, new)           # base.py:192-203
        cur = new
    call.destroy()


@pytest.mark.parametrize("ka", SCHED_KA, ids=[k[0] for k in SCHED_KA])
def test_scheduler_known_answers(ka):
    import ns_gym_b200.schedulers as PS
    import ns_gym_b200.update_functions as PU

    _, builder, times, want, _ref = ka
    call = _evaluator(PU.NoUpdate(builder(PS)), "fp64")
    got = [bool(call(1.0, t)[1]) for t in times]
    assert got == want
    call.destroy()


def test_bridge_transitions_known_answers():
    """The reference's own Bridge tests (tests/test_bridge.py:135-224): action -> cell map, out of
    bounds, forced slip, the side-local distribution in split mode, reward / done from the
    destination cell -- one env per case, cells planted in the state buffer."""
    import torch

    import ns_gym_b200 as nsb
    import ns_gym_b200.schedulers as PS
    import ns_gym_b200.update_functions as PU
    from ns_gym_b200.wrappers import NSBridgeWrapper

    ncol = 8

    def st(r, c):
        return r * ncol + c

    def run(tunable, init, cases):
        env = NSBridgeWrapper(nsb.make("ns_gym/Bridge-v0", num_envs=len(cases)), tunable, initial_prob_dist=init,
                              autoreset="none")
        env.reset(seed=0)
        env.buffers["state"].copy_(torch.tensor([c[0] for c in cases], dtype=torch.int32))
        obs, r, term, trunc, info = env.step(torch.tensor([c[1] for c in cases], dtype=torch.int32))
        return obs["state"].cpu().tolist(), r.cpu().tolist(), term.cpu().tolist()

    nop = lambda: PU.DistributionNoUpdate(PS.ContinuousScheduler())            # noqa: E731
    # test_bridge.py:135-175 deterministic P = [1, 0, 0]
    cases = [(st(2, 4), 0, st(2, 3)), (st(2, 4), 1, st(3, 4)), (st(2, 4), 2, st(2, 5)), (st(2, 4), 3, st(1, 4)),
             (st(1, 1), 0, st(1, 0)), (st(1, 1), 1, st(2, 1)), (st(1, 1), 2, st(1, 2)),
             (st(1, 0), 0, st(1, 0)), (st(2, 7), 2, st(2, 7)), (st(0, 0), 3, st(0, 0)), (st(0, 0), 0, st(0, 0))]
    got, _, _ = run({"P": nop()}, [1.0, 0.0, 0.0], cases)
    assert got == [c[2] for c in cases]
    # :178-197 forced slip to (a + 1) % 4 from S = (2, 4)
    cases = [(st(2, 4), a, st(*rc)) for a, rc in {0: (3, 4), 1: (2, 5), 2: (1, 4), 3: (2, 3)}.items()]
    got, _, _ = run({"P": nop()}, [0.0, 1.0, 0.0], cases)
    assert got == [c[2] for c in cases]
    # :200-224 split mode: the side of the CURRENT cell selects the distribution (col < ncol // 2 = left)
    cases = [(st(2, 1), 0, st(2, 0)),          # left side, P_left = [1, 0, 0]: LEFT goes left
             (st(2, 6), 0, st(3, 6)),          # right side, P_right = [0, 1, 0]: LEFT slips to DOWN
             (st(2, 3), 2, st(2, 4)),          # col 3 is still the left side
             (st(2, 4), 2, st(1, 4))]          # col 4 is the right side: RIGHT slips to UP
    got, _, _ = run({"P_left": nop(), "P_right": nop()}, ([1.0, 0.0, 0.0], [0.0, 1.0, 0.0]), cases)
    assert got == [c[2] for c in cases]
    # envs/Bridge.py:159-174 reward / done from the destination; no absorbing cells (S16)
    cases = [(st(2, 4), 3, st(1, 4)), (st(1, 4), 3, st(0, 4)), (st(0, 4), 1, st(1, 4)), (st(2, 6), 2, st(2, 7)),
             (st(2, 1), 0, st(2, 0))]
    got, rew, term = run({"P": nop()}, [1.0, 0.0, 0.0], cases)
    assert got == [c[2] for c in cases]
    assert rew == [0.0, -1.0, 0.0, 1.0, 1.0] and term == [False, True, False, True, True]
