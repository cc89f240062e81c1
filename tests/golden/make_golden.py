"""Generate golden vectors from the REAL reference (run in the build container only).

    python tests/golden/make_golden.py [case ...]      # no names: every case

For every case of tests/cases.py this imports /root/reference/ns_gym verbatim on top of the
restated gymnasium shim (oracle/ref_loader.py), injects pre-drawn uniform / normal tables
(oracle/streams.py), drives N envs through the next-step-autoreset vector loop and stores
the inputs (actions, tables) and every output (observations, raw fp64 states, theta, rewards,
flags, change masks, deltas) in tests/golden/<case>.npz.  The fixtures travel to the GPU box,
where /root/reference does not exist.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import harness, ref_loader, vector  # noqa: E402
from tests.cases import CASES  # noqa: E402
from tests.parity_util import n_slots_of  # noqa: E402

N_ENVS = 4
SEED = 101


def planning(out_dir, only):
    """tests/golden/planning/<scenario>.npz: traces of the reference's get_planning_env() copies."""
    import warnings

    from tests import parity_util as pu
    from tests.planning_cases import PLAN_CASES

    os.makedirs(os.path.join(out_dir, "planning"), exist_ok=True)
    for name, sc in sorted(PLAN_CASES.items()):
        if only and name not in only:
            continue
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            tr0, tr1, (actions, u, z) = pu.oracle_planning_trace(harness.reference_envs, sc, N_ENVS, SEED)
        tr1 = {k: (v.astype(np.int8) if v.dtype == bool else v) for k, v in tr1.items()}
        np.savez_compressed(os.path.join(out_dir, "planning", f"{name}.npz"), actions=actions, uniforms=u, normals=z,
                            **tr1)
        print(f"planning/{name}: k0={sc['k0']} k1={sc['k1']} ended={int(tr1['terminated'].sum() + tr1['truncated'].sum())}")


def tables(out_dir, only):
    """tests/golden/tables/<case>.npz: unwrapped.P / transition_matrix of the reference at t = 0..T-1."""
    import warnings

    from tests import parity_util as pu

    os.makedirs(os.path.join(out_dir, "tables"), exist_ok=True)
    for name in pu.TABLE_CASES:
        if only and name not in only:
            continue
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            tr = pu.oracle_table_trace(harness.reference_envs, CASES[name], 50)
        np.savez_compressed(os.path.join(out_dir, "tables", f"{name}.npz"), prob=tr["prob"], next=tr["next"],
                            reward=tr["reward"], done=tr["done"].astype(np.int8))
        print(f"tables/{name}: {tr['prob'].shape}")


def main():
    assert ref_loader.available(), "the reference tree is needed to (re)generate golden vectors"
    out_dir = os.path.dirname(os.path.abspath(__file__))
    only = set(sys.argv[1:])
    if "--tables" in only:
        only.discard("--tables")
        return tables(out_dir, only)
    if "--planning" in only:
        only.discard("--planning")
        return planning(out_dir, only)
    for name, case in sorted(CASES.items()):
        if only and name not in only:
            continue
        K = case["steps"]
        actions = harness.draw_actions(case, SEED + 1, K, N_ENVS)
        clock, per_env, u, z = harness.make_streams(SEED, N_ENVS, K + 1, n_slots_of(case))
        envs = harness.reference_envs(case, N_ENVS, per_env)
        tr = vector.trace(vector.SyncVector(envs, per_env, clock), actions)
        tr = {k: (v.astype(np.int8) if v.dtype == bool else v) for k, v in tr.items()}
        np.savez_compressed(os.path.join(out_dir, f"{name}.npz"), actions=actions, uniforms=u, normals=z, **tr)
        print(f"{name}: K={K} N={N_ENVS} ended={int(tr['terminated'].sum() + tr['truncated'].sum())}")


if __name__ == "__main__":
    main()
