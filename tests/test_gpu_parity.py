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
