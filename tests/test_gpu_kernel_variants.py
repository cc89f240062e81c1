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


def _run(case, precision, general, n=4096, steps=40, seed=9, specialize=0, info=None, keep_last=None, **env_kw):
    import torch

    from tests import parity_util as pu

    env = pu.gpu_env(case, n, precision=precision, **env_kw)
    env.set_option("general_kernels", general)
    env.set_option("specialize", specialize)     # off unless a test asks: the precompiled kernels are under test
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
        if info is not None:
            info.setdefault("specialized", []).append(env.last_kernel_specialized)
            info.setdefault("class", []).append(int(env.lib.nsgym_last_kernel_class(env._h)))
        if keep_last is None or k >= steps - keep_last:      # (large batches: keep only the last snapshots)
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


# ---- program-specialised kernels (nsgym_jit.cu): the same device code compiled at run time with the
# ---- program as a compile-time constant, against the precompiled lean kernels on the same native draws
SPECIALISING = ["c1_cartpole_readme", "cartpole_silent", "cartpole_persistent", "c3_acrobot", "c3_mountaincar",
                "c3_pendulum", "mountaincar_continuous", "mountaincar_constraint", "acrobot_constraints"]
# programs with slow-class slots (stochastic schedulers, list / cursor rules, OU / bounded random walks, ...):
# general kernel class, specialised with the slow class unrolled over the slots
SPECIALISING_SLOW = ["cartpole_lists", "cartpole_stochastic", "cartpole_stochastic_scheds", "cartpole_custom_sched",
                     "cartpole_memoryless_lists", "cartpole_all_params"]
SPECIALISING_GRID = ["c2_frozenlake8_drift", "c2_frozenlake8_stepchange", "frozenlake8_lerp", "frozenlake8_cyclic_stale",
                     "frozenlake5_multi_start",
                     "cliff_terminal", "cliff_drift", "c5_bridge_uniform", "c5_bridge_split", "bridge_stepwise"]


def _assert_specialised_where_lean(name, info):
    """Every native-draw program of these lists runs a specialised kernel: the lean classes, and -- with the
    slow class unrolled over the slots -- programs with cursor / list rules too (cartpole_persistent:
    CyclicUpdate, general kernel class)."""
    from ns_gym_b200 import native as nv

    assert all(info["specialized"]), f"{name}: the specialised kernel did not run"
    want = nv.KERNEL_GENERAL if name == "cartpole_persistent" else None
    if want is not None:
        assert all(c == want for c in info["class"])
    else:
        assert all(c in (nv.KERNEL_LEAN_FAST, nv.KERNEL_LEAN_MEDIUM) for c in info["class"]), info["class"][:3]


@pytest.mark.parametrize("name", ["frozenlake8_cyclic_stale", "bridge_stepwise", "c5_bridge_uniform"])
def test_specialised_gridworld_kernels_with_persistent_params_fp64(name):
    import torch

    case = dict(CASES[name])
    case["wrapper"] = dict(case.get("wrapper", {}), persistent_params=True)
    info = {}
    spec = _run(case, "fp64", False, steps=60, specialize=1, info=info)
    general = _run(case, "fp64", True, steps=60)
    assert all(info["specialized"])
    for k, (x, y) in enumerate(zip(spec, general)):
        for key in x:
            assert torch.equal(x[key], y[key]), f"{name}: {key} differs at step {k}"


@pytest.mark.parametrize("buffers", [dict(), dict(want_delta=False, want_obs=False)])
@pytest.mark.parametrize("name", SPECIALISING + SPECIALISING_GRID)
def test_specialised_kernels_equal_precompiled_kernels_fp64(name, buffers):
    import torch

    case = CASES[name]
    info = {}
    spec = _run(case, "fp64", False, specialize=1, info=info, **buffers)
    lean = _run(case, "fp64", False, specialize=0, **buffers)
    _assert_specialised_where_lean(name, info)
    for k, (x, y) in enumerate(zip(spec, lean)):
        for key in x:
            assert torch.equal(x[key], y[key]), f"{name}: {key} differs at step {k}"


@pytest.mark.parametrize("name", SPECIALISING)
def test_specialised_kernels_track_precompiled_kernels_fp32(name):
    """fp32: folded constants may change an FMA contraction, so the two kernels agree to rounding over the
    first steps; the integer outputs of step 0 are identical."""
    import torch

    case = CASES[name]
    info = {}
    spec = _run(case, "fp32", False, steps=6, specialize=1, info=info)
    lean = _run(case, "fp32", False, steps=6, specialize=0)
    _assert_specialised_where_lean(name, info)
    for key in ("flags", "change", "t"):
        assert torch.equal(spec[0][key], lean[0][key])
    for k in range(6):
        for key in ("state", "theta", "reward", "obs", "delta"):
            if key in spec[k]:
                np.testing.assert_allclose(spec[k][key].cpu().numpy(), lean[k][key].cpu().numpy(),
                                           rtol=2e-4, atol=2e-5, err_msg=f"{name}: {key} step {k}")


def test_specialisation_is_automatic_for_large_batches_only():
    import torch

    from tests import parity_util as pu

    for n, want in ((4096, False), (32768, True)):
        env = pu.gpu_env(CASES["c1_cartpole_readme"], n, precision="fp32")
        env.reset(seed=1)
        env.step_raw(torch.zeros(n, dtype=torch.int32, device=env.device))
        assert env.last_kernel_specialized == want
    # general kernels (injection, slow rule classes) never specialise
    env = pu.gpu_env(CASES["c1_cartpole_readme"], 32768, precision="fp32")
    env.set_option("general_kernels", 1)
    env.reset(seed=1)
    env.step_raw(torch.zeros(32768, dtype=torch.int32, device=env.device))
    assert not env.last_kernel_specialized


def _het_frozenlake_stochastic(S, U, e):
    """per-env rows of the general class on a gridworld: a stochastic scheduler per env (the Memoryless one
    seeded: its first transition time is drawn at construction, schedulers.py:92-116)"""
    r = np.random.default_rng([8, e])
    if int(r.integers(0, 2)):
        return {"P": U.DistributionDecrementUpdate(S.RandomScheduler(float(r.uniform(0.2, 0.9))), k=float(r.uniform(0.01, 0.05)))}
    return {"P": U.UniformDrift(S.MemorylessScheduler(p=float(r.uniform(0.2, 0.6)), seed=e), rate=float(r.uniform(0.02, 0.1)))}


EXTRA_CASES = {
    "het_frozenlake_stochastic": dict(env_id="FrozenLake-v1", params_of=_het_frozenlake_stochastic,
                                      wrapper=dict(initial_prob_dist=[0.9, 0.05, 0.05], change_notification=True,
                                                   delta_change_notification=True), make={}, steps=60),
}


@pytest.mark.parametrize("name", ["het_cartpole_wide", "het_frozenlake_stochastic"])
def test_specialised_general_row_kernels_equal_precompiled_row_kernels(name):
    """Per-env rows of the general class (stochastic schedulers, cursor rules, slow updates per env): the kernel
    specialised on program and row layout (slot loop unrolled, row loads up front, no injection code) against the
    precompiled general per-env kernel, bit for bit in fp64."""
    import torch

    from ns_gym_b200 import native as nv

    case = CASES.get(name) or EXTRA_CASES[name]
    info = {}
    spec = _run(case, "fp64", False, specialize=1, info=info, steps=40)
    pre = _run(case, "fp64", False, specialize=0, steps=40)
    assert all(c == nv.KERNEL_ROWS_GENERAL for c in info["class"]) and all(info["specialized"]), (info["class"][:2], info["specialized"][:2])
    for k, (x, y) in enumerate(zip(spec, pre)):
        for key in x:
            assert torch.equal(x[key], y[key]), f"{name}: {key} differs at step {k}"


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("name", ["het_cartpole_lean", "c4_cartpole_rows", "c4_frozenlake8_rows"])
def test_specialised_row_kernels_equal_precompiled_row_kernels(name, precision):
    """Per-env rows: the kernel specialised on the program AND the row layout (which words vary, their
    planes, the shared defaults) against the precompiled lean per-env kernel."""
    import torch

    from ns_gym_b200 import native as nv

    case = CASES[name]
    if precision == "fp32" and "frozenlake" in name:
        pytest.skip("gridworld probabilities are always fp64")
    info = {}
    spec = _run(case, precision, False, specialize=1, info=info, steps=30)
    lean = _run(case, precision, False, specialize=0, steps=30)
    assert all(c == nv.KERNEL_ROWS_LEAN for c in info["class"]) and all(info["specialized"])
    for k, (x, y) in enumerate(zip(spec, lean)):
        for key in x:
            if precision == "fp64" or key in ("flags", "change", "t", "istate"):
                if precision == "fp32" and k > 0:
                    continue
                assert torch.equal(x[key], y[key]), f"{name}: {key} differs at step {k}"
            elif k < 6:
                np.testing.assert_allclose(x[key].cpu().numpy(), y[key].cpu().numpy(), rtol=2e-4, atol=2e-5,
                                           err_msg=f"{name}: {key} step {k}")


@pytest.mark.parametrize("n", [(1 << 21) + 100, 1 << 21, (1 << 21) + 101])
@pytest.mark.parametrize("name", ["c5_bridge_uniform", "c2_frozenlake8_stepchange", "frozenlake8_cyclic_stale",
                                  "cliff_drift", "bridge_stepwise", "frozenlake5_multi_start"])
def test_tiled_gridworld_kernels_equal_precompiled_kernels(name, n):
    """Batches of >= 2^21 envs run the tiled specialised kernel (TMA bulk copies prefetch the planes of the
    next tiles into shared memory) on their full tiles of 256 envs and the one-thread-one-env kernel on the
    remainder (n = 2^21 + 100); a batch whose plane stride is not a multiple of 16 bytes (n odd) cannot
    be streamed and keeps the plain kernel.  All must equal the precompiled lean kernel bit for bit."""
    import torch

    case = CASES[name]
    info = {}
    spec = _run(case, "fp64", False, n=n, steps=16, specialize=1, info=info, keep_last=4)
    lean = _run(case, "fp64", False, n=n, steps=16, specialize=0, keep_last=4)
    assert all(info["specialized"])
    for k, (x, y) in enumerate(zip(spec, lean)):
        for key in x:
            assert torch.equal(x[key], y[key]), f"{name}: {key} differs at step {16 - len(spec) + k}"


def test_row_words_are_bit_packed():
    """The varying int words of a per-env row are packed to the width the batch needs: a C4 CartPole slot
    (opcode word, modulus, on-count, two pool offsets vary) takes 2 int planes + 5 coefficient planes = 28 B,
    not 10 planes; the layout reports it and the algorithmic bytes per env-step follow."""
    from tests import parity_util as pu

    env = pu.gpu_env(CASES["c4_cartpole_rows"], 4096, precision="fp32", want_delta=True, want_obs=False)
    rows = env.row_bytes_per_env
    assert 0 < rows < 80, rows          # two slots; 80 B = one plane per varying word (round 1)
    assert abs(env.bytes_per_step - (66 + 8 + rows)) < 1e-9
    g = pu.gpu_env(CASES["c4_frozenlake8_rows"], 4096, precision="fp64")
    assert 0 < g.row_bytes_per_env < 60, g.row_bytes_per_env      # 60 B unpacked


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("name", SPECIALISING_SLOW)
def test_specialised_general_class_kernels_equal_precompiled_kernels(name, precision):
    """Programs with slow-class slots on native draws: the specialised kernel (level 2, slow class unrolled over
    the slots, no injection code) against the precompiled general kernel -- fp64 bit for bit, fp32 to rounding."""
    import torch

    from ns_gym_b200 import native as nv

    if name not in CASES:
        pytest.skip(f"no case {name}")
    case = CASES[name]
    info = {}
    steps = 40 if precision == "fp64" else 6
    spec = _run(case, precision, False, specialize=1, info=info, steps=steps)
    pre = _run(case, precision, False, specialize=0, steps=steps)
    assert all(info["specialized"]) and all(c == nv.KERNEL_GENERAL for c in info["class"])
    for k, (x, y) in enumerate(zip(spec, pre)):
        for key in x:
            if precision == "fp64" or (k == 0 and key in ("flags", "change", "t", "istate")):
                assert torch.equal(x[key], y[key]), f"{name}: {key} differs at step {k}"
            elif key in ("state", "theta", "reward", "obs", "delta"):
                np.testing.assert_allclose(x[key].cpu().numpy(), y[key].cpu().numpy(), rtol=2e-4, atol=2e-5,
                                           err_msg=f"{name}: {key} step {k}")


@pytest.mark.parametrize("name", ["frozenlake4_memoryless_cyclic", "cliff_random_categorical", "bridge_lipschitz_bounded"])
def test_specialised_general_class_gridworld_kernels_equal_precompiled_kernels(name):
    """Gridworld programs with stochastic schedulers / RandomCategorical / the Lipschitz-bounded wrapper on native
    draws: specialised general-class kernel against the precompiled one, bit for bit."""
    import torch

    from ns_gym_b200 import native as nv

    case = CASES[name]
    info = {}
    spec = _run(case, "fp64", False, specialize=1, info=info, steps=40)
    pre = _run(case, "fp64", False, specialize=0, steps=40)
    assert all(info["specialized"]) and all(c == nv.KERNEL_GENERAL for c in info["class"])
    for k, (x, y) in enumerate(zip(spec, pre)):
        for key in x:
            assert torch.equal(x[key], y[key]), f"{name}: {key} differs at step {k}"


@pytest.mark.parametrize("in_sim_change", [False, True])
@pytest.mark.parametrize("name,precision", [("c1_cartpole_readme", "fp64"), ("cartpole_persistent", "fp64"),
                                            ("c5_bridge_uniform", "fp64"), ("c3_pendulum", "fp64")])
def test_specialised_kernels_on_planning_copies(name, precision, in_sim_change):
    """A planning copy is not a root env: its TimeLimit counts from the copy (plan_elapsed) and, with
    in_sim_change False, its parameters are frozen (skip_updates) -- launch facts the specialised kernels keep
    as run-time values (SpecFix::root = -1).  Steps and fused rollouts of a copy: specialised = precompiled."""
    import torch

    from tests import parity_util as pu

    case = dict(CASES[name])
    case["wrapper"] = dict(case.get("wrapper", {}), in_sim_change=in_sim_change)
    out = {}
    for specialize in (0, 1):
        root = pu.gpu_env(case, 512, precision=precision)
        root.reset(seed=3)
        for _ in range(5):
            root.step_raw(root.action_space.sample() * 0)
        plan = root.get_planning_env(fanout=4, seed=99)
        plan.set_option("specialize", specialize)
        a = plan.action_space.sample() * 0
        for _ in range(6):
            plan.step_raw(a)
        assert plan.last_kernel_specialized == bool(specialize)
        mid = {k: v.clone() for k, v in plan.buffers.items() if v is not None}
        ret, length = plan.rollout(12, gamma=0.97)
        assert plan.last_kernel_specialized == bool(specialize)
        out[specialize] = (mid, {k: v.clone() for k, v in plan.buffers.items() if v is not None}, ret.clone(), length.clone())
    for part in (0, 1):
        for key in out[0][part]:
            assert torch.equal(out[0][part][key], out[1][part][key]), f"{name}: {key} differs ({'steps' if part == 0 else 'rollout'})"
    assert torch.equal(out[0][2], out[1][2]) and torch.equal(out[0][3], out[1][3])
