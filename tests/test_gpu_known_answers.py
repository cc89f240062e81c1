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
