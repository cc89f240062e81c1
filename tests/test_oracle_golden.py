"""CPU: the oracle port reproduces the golden vectors generated from the REAL reference
(tests/golden/make_golden.py) bit for bit.  Runs everywhere, including the GPU box where the
reference tree is absent."""
import glob
import os

import numpy as np
import pytest

from oracle import harness, streams, vector
from tests.cases import CASES

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


def test_every_case_has_a_golden_file():
    assert set(NAMES) == set(CASES)


@pytest.mark.parametrize("name", NAMES)
def test_port_matches_golden(name):
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    case = CASES[name]
    u, z, actions = g["uniforms"], g["normals"], g["actions"]
    n = u.shape[2]
    clock = streams.Clock()
    per_env = [streams.EnvStreams(u[:, :, i], z[:, :, i], clock) for i in range(n)]
    envs = harness.port_envs(case, n, per_env)
    tr = vector.trace(vector.SyncVector(envs, per_env, clock), actions)
    for key, got in tr.items():
        want = g[key]
        assert np.array_equal(np.asarray(got).astype(want.dtype) if want.dtype.kind in "iu" else got, want,
                              equal_nan=got.dtype.kind == "f"), f"{name}: {key} differs from golden"


# ---- planning copies ------------------------------------------------------------------------------
from tests import parity_util as pu  # noqa: E402
from tests.planning_cases import PLAN_CASES  # noqa: E402

PLAN_NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "planning", "*.npz")))


def test_every_planning_case_has_a_golden_file():
    assert set(PLAN_NAMES) == set(PLAN_CASES)


@pytest.mark.parametrize("name", PLAN_NAMES)
def test_port_planning_copies_match_golden(name):
    import warnings

    g = np.load(os.path.join(GOLDEN, "planning", f"{name}.npz"))
    n = g["uniforms"].shape[2]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _, tr, _ = pu.oracle_planning_trace(harness.port_envs, PLAN_CASES[name], n, 0,
                                            tables=(g["actions"], g["uniforms"], g["normals"]))
    for key, got in tr.items():
        want = g[key]
        assert np.array_equal(np.asarray(got).astype(want.dtype) if want.dtype.kind in "iu" else got, want,
                              equal_nan=got.dtype.kind == "f"), f"{name}: {key} differs from golden"


# ---- transition tables --------------------------------------------------------------------------------
@pytest.mark.parametrize("name", pu.TABLE_CASES)
def test_port_transition_tables_match_golden(name):
    import warnings

    g = np.load(os.path.join(GOLDEN, "tables", f"{name}.npz"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        tr = pu.oracle_table_trace(harness.port_envs, CASES[name], g["prob"].shape[0])
    assert np.array_equal(tr["prob"], g["prob"]) and np.array_equal(tr["next"], g["next"])
    assert np.array_equal(tr["reward"], g["reward"]) and np.array_equal(tr["done"].astype(np.int8), g["done"])
