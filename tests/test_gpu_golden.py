"""GPU: the CUDA path, through the C ABI, against the golden vectors generated from the REAL
reference (tests/golden/make_golden.py), fp64 parity mode, same injected tables."""
import glob
import os

import numpy as np
import pytest

from tests import parity_util as pu
from tests.cases import CASES

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


@pytest.mark.parametrize("name", NAMES)
def test_cuda_matches_golden(name):
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    ref = {k: g[k] for k in g.files if k not in ("actions", "uniforms", "normals")}
    n = g["uniforms"].shape[2]
    got = pu.gpu_trace(CASES[name], n, g["actions"], g["uniforms"], g["normals"], precision="fp64")
    pu.compare(ref, got, float_obs_rtol=1e-6, name=name)
