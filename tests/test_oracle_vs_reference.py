"""Pin the oracle port against the REAL reference, imported verbatim from /root/reference
(this container only; the GPU box has no reference tree, there the golden vectors stand in).

Both sides consume identical pre-drawn uniform / normal tables (oracle/streams.py) and the
same action sequence, through the same next-step-autoreset vector loop.
"""
import numpy as np
import pytest

from oracle import harness, ref_loader, vector
from tests.cases import CASES

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")

N_ENVS = 6


def _run_both(case, n_envs=N_ENVS, seed=11):
    K = case["steps"]
    n_slots = len(case["params"](*_port_ns()))
    actions = harness.draw_actions(case, seed + 1, K, n_envs)
    out = []
    for builder in (harness.reference_envs, harness.port_envs):
        clock, per_env, _, _ = harness.make_streams(seed, n_envs, K + 1, n_slots)
        envs = builder(case, n_envs, per_env)
        out.append(vector.trace(vector.SyncVector(envs, per_env, clock), actions))
    return out


def _port_ns():
    import ns_gym_b200.schedulers as PS
    import ns_gym_b200.update_functions as PU
    return PS, PU


@pytest.mark.parametrize("name", sorted(CASES))
def test_port_matches_reference_bit_for_bit(name):
    ref, port = _run_both(CASES[name])
    assert ref["terminated"].any() or ref["truncated"].any() or "pendulum" in name or name in (
        "cartpole_silent", "c3_acrobot", "acrobot_constraints", "mountaincar_continuous",
        "mountaincar_constraint"), "case never ends an episode"
    for key in ref:
        a, b = ref[key], port[key]
        assert a.shape == b.shape, key
        assert np.array_equal(a, b, equal_nan=True), (
            f"{name}: {key} differs at {np.argwhere(a != b)[:5]}")


# ---- planning copies: get_planning_env() of the reference vs the port -----------------------------
from tests import parity_util as pu  # noqa: E402
from tests.planning_cases import PLAN_CASES  # noqa: E402


@pytest.mark.parametrize("name", sorted(PLAN_CASES))
def test_port_planning_copies_match_reference(name):
    import warnings

    sc = PLAN_CASES[name]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _, ref, tables = pu.oracle_planning_trace(harness.reference_envs, sc, N_ENVS, seed=31)
        _, port, _ = pu.oracle_planning_trace(harness.port_envs, sc, N_ENVS, seed=31, tables=tables)
    for key in ref:
        assert np.array_equal(ref[key], port[key], equal_nan=True), f"{name}: {key} differs"


# ---- time-indexed transition tables: unwrapped.P / transition_matrix vs the port -------------------
@pytest.mark.parametrize("name", pu.TABLE_CASES)
def test_port_transition_tables_match_reference(name):
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = pu.oracle_table_trace(harness.reference_envs, CASES[name], 40)
        port = pu.oracle_table_trace(harness.port_envs, CASES[name], 40)
    for key in ref:
        assert np.array_equal(ref[key], port[key]), f"{name}: {key} differs"
