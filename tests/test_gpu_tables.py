"""GPU: time-indexed transition tables (nsgym_transition_table) against the oracle port and the
golden tables read off the REAL reference's ``unwrapped.P`` / ``Bridge.transition_matrix``."""
import os

import numpy as np
import pytest

from oracle import harness
from tests import parity_util as pu
from tests.cases import CASES

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "tables")


def _gpu_table(case, T):
    env = pu.gpu_env(case, 8, autoreset="none")
    env.reset(seed=1)
    tab = env.transition_table(T=T, env=3)
    return {k: v.cpu().numpy() for k, v in tab.items()}


def _check(ref, got, name):
    assert np.array_equal(ref["next"], got["next"]), f"{name}: next"
    assert np.array_equal(ref["done"].astype(np.uint8), got["done"].astype(np.uint8)), f"{name}: done"
    assert np.array_equal(ref["reward"].astype(np.float32), got["reward"].astype(np.float32)), f"{name}: reward"
    np.testing.assert_allclose(got["prob"], ref["prob"], rtol=1e-12, atol=1e-15, err_msg=f"{name}: prob")


@pytest.mark.parametrize("name", pu.TABLE_CASES)
def test_tables_match_oracle(name):
    T = 60
    ref = pu.oracle_table_trace(harness.port_envs, CASES[name], T)
    _check(ref, _gpu_table(CASES[name], T), name)


@pytest.mark.parametrize("name", pu.TABLE_CASES)
def test_tables_match_golden(name):
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    ref = {k: g[k] for k in g.files}
    _check(ref, _gpu_table(CASES[name], ref["prob"].shape[0]), name)


def test_table_rows_are_distributions():
    tab = _gpu_table(CASES["c5_bridge_split"], 30)
    np.testing.assert_allclose(tab["prob"].sum(-1), 1.0, atol=1e-12)


@pytest.mark.parametrize("name,env", [("c4_frozenlake8_rows", 3), ("c4_frozenlake8_rows", 6), ("het_bridge_split", 2)])
def test_tables_of_a_heterogeneous_batch_follow_the_env_rows(name, env):
    T = 50
    ref = pu.oracle_table_trace(harness.port_envs, CASES[name], T, env=env)
    genv = pu.gpu_env(CASES[name], 8, autoreset="none")
    genv.reset(seed=1)
    got = {k: v.cpu().numpy() for k, v in genv.transition_table(T=T, env=env).items()}
    _check(ref, got, f"{name}[{env}]")
