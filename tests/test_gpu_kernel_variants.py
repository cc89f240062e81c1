"""GPU: the lean kernel instantiations (fast / medium rule classes only, injection folded away,
lean per-env kernels) against the general instantiations the oracle parity tests exercise.

The parity tests inject random tables, which selects the general kernels; throughput runs use
native Philox draws and the lean kernels.  Both are launched here on the same native draws:
fp64 mode must agree bit for bit (same -fmad=false arithmetic), fp32 mode within the contraction
differences of two separately compiled kernels."""
import numpy as np
import pytest

from tests.cases import CASES

pytestmark = pytest.mark.gpu

CASES_UNDER_TEST = ["c1_cartpole_readme", "cartpole_silent", "cartpole_persistent", "c3_acrobot", "c3_mountaincar",
                    "c3_pendulum", "mountaincar_continuous", "mountaincar_constraint", "acrobot_constraints",
                    "c2_frozenlake8_drift", "c2_frozenlake8_stepchange", "frozenlake8_lerp", "cliff_terminal",
                    "c5_bridge_uniform", "c5_bridge_split", "bridge_stepwise", "het_cartpole_lean",
                    "c4_cartpole_rows", "c4_frozenlake8_rows"]


def _run(case, precision, general, n=4096, steps=40, seed=9):
    import torch

    from tests import parity_util as pu

    env = pu.gpu_env(case, n, precision=precision)
    env.set_option("general_kernels", general)
    env.reset(seed=seed)
    g = torch.Generator(device=env.device)
    g.manual_seed(1)
    out = []
    for k in range(steps):
        if env.action_space_n is None:
            a = torch.rand(n, generator=g, device=env.device, dtype=env.real) * 2 - 1
        else:
            a = torch.randint(0, env.action_space_n, (n,), generator=g, device=env.device, dtype=torch.int32)
        env.step_raw(a)
        out.append({k2: v.clone() for k2, v in env.buffers.items() if v is not None and k2 != "action"})
    return out


@pytest.mark.parametrize("name", CASES_UNDER_TEST)
def test_lean_kernels_equal_general_kernels_fp64(name):
    import torch

    case = CASES[name]
    lean, general = _run(case, "fp64", False), _run(case, "fp64", True)
    for k, (x, y) in enumerate(zip(lean, general)):
        for key in x:
            assert torch.equal(x[key], y[key]), f"{name}: {key} differs at step {k}"


@pytest.mark.parametrize("name", ["frozenlake8_cyclic_stale", "c2_frozenlake8_stepchange", "bridge_stepwise",
                                  "c5_bridge_uniform", "cliff_drift"])
def test_lean_gridworld_kernels_with_persistent_params_fp64(name):
    """persistent_params=True: a reset keeps distributions and cursors, so the lean kernels (which
    write a plane back only when its update fired or a reset re-initialised it) store nothing on a
    reset step -- the general kernels store everything; both must leave identical buffers."""
    import torch

    case = dict(CASES[name])
    case["wrapper"] = dict(case.get("wrapper", {}), persistent_params=True)
    lean, general = _run(case, "fp64", False, steps=60), _run(case, "fp64", True, steps=60)
    for k, (x, y) in enumerate(zip(lean, general)):
        for key in x:
            assert torch.equal(x[key], y[key]), f"{name}: {key} differs at step {k}"


@pytest.mark.parametrize("name", [n for n in CASES_UNDER_TEST if "frozenlake" not in n and "bridge" not in n
                                  and "cliff" not in n])
def test_lean_kernels_track_general_kernels_fp32(name):
    """fp32: the first steps agree to rounding (later, chaotic dynamics amplify the last-bit
    differences of two separately contracted kernels); integer outputs of step 0 are identical."""
    import torch

    case = CASES[name]
    lean, general = _run(case, "fp32", False, steps=6), _run(case, "fp32", True, steps=6)
    for key in ("flags", "change", "t"):
        assert torch.equal(lean[0][key], general[0][key])
    for k in range(6):
        for key in ("state", "theta"):
            if key in lean[k]:
                np.testing.assert_allclose(lean[k][key].cpu().numpy(), general[k][key].cpu().numpy(),
                                           rtol=2e-4, atol=2e-5, err_msg=f"{name}: {key} step {k}")
