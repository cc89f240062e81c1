"""GPU: the kernels' W1 (delta_change of a distribution parameter, base.py:192-203 ->
utils.py:55-94) shares one reciprocal refinement per divisor instead of dividing per quotient.
It must equal the plain IEEE-division sum BIT FOR BIT, and the oracle's restatement of scipy's
CDF algorithm, on adversarial inputs: sums one ulp off 1, zeros, tiny / huge weights."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cases(dim, n, seed):
    r = np.random.default_rng([77, dim, seed])
    blocks = []
    m = n // 8
    blocks.append(r.dirichlet(np.ones(dim), m))                                  # generic, sum ~ 1
    d = r.dirichlet(np.ones(dim) * 0.2, m)                                        # peaked
    blocks.append(d)
    x = r.dirichlet(np.ones(dim), m)
    x *= (1.0 + r.integers(-4, 5, (m, 1)) * 2.0 ** -52)                           # sums a few ulps off 1
    blocks.append(x)
    z = r.dirichlet(np.ones(dim), m)
    z[r.random((m, dim)) < 0.3] = 0.0                                             # exact zeros (leading too)
    blocks.append(z)
    blocks.append(r.random((m, dim)) * 10.0 ** r.integers(-300, 300, (m, 1)))     # far from normalised
    blocks.append(r.random((m, dim)) * 10.0 ** r.integers(-320, -290, (m, dim)))  # subnormal neighbourhood
    u = 0.9 * 0.95 ** r.integers(0, 400, m)                                       # UniformDrift trajectories
    blocks.append(np.stack([u] + [(1.0 - u) / (dim - 1)] * (dim - 1), 1))
    s = r.dirichlet(np.ones(dim), n - 7 * m)
    s[r.random(s.shape[0]) < 0.05, 0] = -1e-3                                     # negatives -> NaN ("raises")
    s[r.random(s.shape[0]) < 0.02] = 0.0                                          # zero mass
    blocks.append(s)
    return np.concatenate(blocks, 0)


@pytest.mark.parametrize("dim", [3, 4])
def test_w1_shared_reciprocal_is_bit_exact(dim):
    import torch

    from ns_gym_b200 import native as nv
    from oracle.ns_port import w1_index_distance

    lib = nv.load()
    n = 1 << 21
    u = _cases(dim, n, 0)
    v = _cases(dim, n, 1)[np.random.default_rng(5).permutation(n)]
    dev = torch.device("cuda")
    du = torch.tensor(u.T.copy(), device=dev)
    dv = torch.tensor(v.T.copy(), device=dev)
    out = torch.empty(n, dtype=torch.float64, device=dev)
    ref = torch.empty(n, dtype=torch.float64, device=dev)
    nv.check(lib.nsgym_eval_w1(dim, du.data_ptr(), dv.data_ptr(), out.data_ptr(), ref.data_ptr(), n, None))
    torch.cuda.synchronize()
    a = out.cpu().numpy().view(np.uint64)
    b = ref.cpu().numpy().view(np.uint64)
    both_nan = np.isnan(out.cpu().numpy()) & np.isnan(ref.cpu().numpy())
    diff = (a != b) & ~both_nan
    assert not diff.any(), f"{diff.sum()} of {n} W1 values differ from the IEEE-division sum, first at {np.flatnonzero(diff)[:5]}"
    assert 0 < both_nan.sum() < n // 8
    # against the oracle's scipy restatement on a sample (pure Python: small)
    o = out.cpu().numpy()
    for i in np.random.default_rng(9).choice(n, 3000, replace=False):
        try:
            want = w1_index_distance(u[i], v[i])
        except ValueError:
            assert np.isnan(o[i])
            continue
        if np.isnan(want):
            continue
        assert o[i] == want, (i, u[i], v[i], o[i], want)
