"""GPU parity tests proper: the CUDA step path, called through the C ABI, against the oracle
port on the same injected uniform / normal tables and the same actions.

fp64 parity mode: integer states, flags, fire indices and change masks bit-exact; floating
state / theta / delta within 1e-9 relative (north_star; the only non-bit-exact ingredients
are CUDA's sin / cos / exp, <= 2 ulp from glibc's).
"""
import numpy as np
import pytest

from tests.cases import CASES
from tests import parity_util as pu

pytestmark = pytest.mark.gpu

N_ENVS = 48


@pytest.mark.parametrize("name", sorted(CASES))
def test_fp64_matches_oracle(name):
    case = CASES[name]
    ref, actions, u, z = pu.oracle_trace(case, N_ENVS, seed=21)
    got = pu.gpu_trace(case, N_ENVS, actions, u, z, precision="fp64")
    assert not got["_bad_dist"]
    # float32 observations are rounded from fp64 values that may differ in the last bits
    pu.compare(ref, got, float_obs_rtol=1e-6, name=name)


def test_bucketed_heterogeneous_batch_is_a_permutation():
    """bucket=True stores the envs sorted by opcode signature; with the injected tables and actions
    permuted the same way the results are the un-bucketed results, permuted."""
    import ns_gym_b200.schedulers as PS
    import ns_gym_b200.update_functions as PU
    from ns_gym_b200.vector_env import NSVectorEnv

    case, n = CASES["c4_cartpole_rows"], 64
    ref, actions, u, z = pu.oracle_trace(case, n, seed=5)
    kw = dict(precision="fp64", want_delta=True, want_obs=True, **case["wrapper"], **case["make"])
    env = NSVectorEnv.heterogeneous(case["env_id"], [case["params_of"](PS, PU, e) for e in range(n)], bucket=True, **kw)
    order = env.env_order
    assert sorted(order.tolist()) == list(range(n)) and not np.array_equal(order, np.arange(n))
    got = pu.gpu_run(env, actions[:, order], u[:, :, order], z[:, :, order])
    inv = np.argsort(order)
    got = {k: (v if not isinstance(v, np.ndarray) else (v[inv] if k in ("obs0", "raw0") else v[:, inv]))
           for k, v in got.items()}
    pu.compare(ref, got, float_obs_rtol=1e-6, name="bucketed c4_cartpole_rows")
    assert np.array_equal(env.to_caller(env.to_storage(np.arange(n))), np.arange(n))
