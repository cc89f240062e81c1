"""CPU: the program specialiser (nsgym_jit.cu) generates and compiles -- NVRTC, no device needed -- the
step kernel of every case program that stays in the lean kernel classes, in both precisions."""
import ctypes as C

import pytest

import ns_gym_b200.schedulers as S
import ns_gym_b200.update_functions as U
from ns_gym_b200 import native as nv
from ns_gym_b200.compile import compile_program
from tests.cases import CASES

CLASSIC = sorted(n for n, c in CASES.items()
                 if "params" in c and c["env_id"].split("-")[0] in
                 ("CartPole", "Acrobot", "MountainCar", "MountainCarContinuous", "Pendulum"))
GRID = sorted(n for n, c in CASES.items() if "params" in c and n not in CLASSIC)


def _check(spec, want_delta, want_obs, rollout=0):
    lib = nv.load()
    src = C.create_string_buffer(1 << 16)
    log = C.create_string_buffer(1 << 14)
    rc = lib.nsgym_jit_check(C.byref(spec), rollout, want_delta, want_obs, src, len(src), log, len(log))
    return rc, src.value.decode(), log.value.decode(), (lib.nsgym_last_error() or b"").decode()


def test_nvrtc_is_available():
    p = compile_program("CartPole-v1", {"masspole": U.IncrementUpdate(S.ContinuousScheduler(), k=0.1)}, 16)
    rc, src, log, err = _check(p.spec, 0, 0)
    assert rc > 0, (rc, err, log)
    assert "classic_step_body<float, 0, 1, 0, nsg::SpecFix>" in src
    assert "prefetch = 0, want_delta = 0, has_obs = 0, root = 1" in src


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
@pytest.mark.parametrize("name", CLASSIC)
def test_every_lean_classic_case_specialises(name, precision):
    c = CASES[name]
    p = compile_program(c["env_id"], c["params"](S, U), 16, precision=precision, **c["wrapper"], **c["make"])
    rc, src, log, err = _check(p.spec, 1, 1)
    assert rc > 0, (name, rc, err, log)     # lean classes and programs with slow-class slots alike
    real = "float" if precision == "fp32" else "double"
    assert f"nsg::StepIO<{real}>" in src and "want_delta = 1" in src
    if len(c["params"](S, U)):     # the fused rollout of the same program
        rc, src, log, err = _check(p.spec, 0, 1, rollout=1)
        assert rc > 0 and "classic_rollout_body<" in src, (name, rc, err, log)


@pytest.mark.parametrize("name", GRID)
def test_every_lean_gridworld_case_specialises(name):
    c = CASES[name]
    p = compile_program(c["env_id"], c["params"](S, U), 16, precision="fp64", **c["wrapper"], **c["make"])
    rc, src, log, err = _check(p.spec, 1, 0)
    assert rc > 0, (name, rc, err, log)     # deterministic (lean) and stochastic (general class) programs alike
    assert "grid_step_body<" in src and "GridPtrs ptrs" in src
    rc, src, log, err = _check(p.spec, 0, 0, rollout=1)
    assert rc > 0 and "grid_rollout_body<" in src, (name, rc, err, log)


def test_distinct_programs_give_distinct_sources():
    a = compile_program("CartPole-v1", {"masspole": U.IncrementUpdate(S.ContinuousScheduler(), k=0.1)}, 16)
    b = compile_program("CartPole-v1", {"masspole": U.IncrementUpdate(S.ContinuousScheduler(), k=0.2)}, 16)
    sa, sb = _check(a.spec, 0, 0)[1], _check(b.spec, 0, 0)[1]
    assert sa and sb and sa != sb
    # the batch size and the seed are launch arguments, not part of the program
    c = compile_program("CartPole-v1", {"masspole": U.IncrementUpdate(S.ContinuousScheduler(), k=0.1)}, 4096, seed=7)
    assert _check(c.spec, 0, 0)[1] == sa
