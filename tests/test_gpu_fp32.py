"""GPU: fp32 FAST mode against the fp64 oracle, with the tolerance stated here.

fp32 mode stores state and theta as float32 and uses MUFU sin / cos and reciprocal-multiply
division in the dynamics, so it is compared with the oracle only while both run the SAME
episode (up to each env's first episode end in the oracle: a one-ulp difference near a
termination threshold can move an episode boundary by a step, after which the trajectories
are different episodes).  Stated tolerance over that window (<= ~60 free-running steps):

    theta (update functions)    rtol 2e-5 + atol 2e-5 (random walks cancel towards 0)
    state / observation         atol 2e-3 + rtol 2e-3   (chaotic amplification included)
    fire flags / change masks   bit-exact (schedulers compare integers, fp64 thresholds)
    integer gridworld paths     not applicable: gridworlds always run fp64 probabilities
"""
import numpy as np
import pytest

from tests import parity_util as pu
from tests.cases import CASES

pytestmark = pytest.mark.gpu

FP32_CASES = ["c1_cartpole_readme", "cartpole_all_params", "cartpole_lists", "cartpole_stochastic",
              "cartpole_constraint", "cartpole_persistent", "c3_acrobot", "c3_mountaincar", "c3_pendulum",
              "mountaincar_continuous", "pendulum_all",
              "c4_cartpole_rows", "het_cartpole_lean"]      # heterogeneous rows: the lean per-env kernel


@pytest.mark.parametrize("name", FP32_CASES)
def test_fp32_fast_mode_tracks_oracle(name):
    case = CASES[name]
    n, steps = 32, min(case["steps"], 60)
    ref, actions, u, z = pu.oracle_trace(case, n, seed=33, steps=steps)
    got = pu.gpu_trace(case, n, actions, u, z, precision="fp32")
    ended = (ref["terminated"] | ref["truncated"] | ref["was_reset"])
    alive = np.cumsum(ended, axis=0) == 0                    # [K, N]: before the first episode end
    # a constraint rejection can flip on a last-bit difference: compare flags where theta agrees
    assert alive[:5].all()
    for k in range(steps):
        m = alive[k]
        if not m.any():
            continue
        assert np.array_equal(ref["gt_change"][k][m], got["gt_change"][k][m]), f"{name}: fire flags differ at step {k}"
        np.testing.assert_allclose(got["theta"][k][m], ref["theta"][k][m], rtol=2e-5, atol=2e-5,
                                   err_msg=f"{name}: theta step {k}")
        np.testing.assert_allclose(got["raw"][k][m], ref["raw"][k][m], rtol=2e-3, atol=2e-3,
                                   err_msg=f"{name}: state step {k}")
        np.testing.assert_allclose(got["reward"][k][m], ref["reward"][k][m], rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(got["raw0"], ref["raw0"], rtol=1e-6, atol=1e-7)
