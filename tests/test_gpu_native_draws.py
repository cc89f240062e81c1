"""GPU: the NATIVE Philox draws the throughput kernels consume (parity tests inject tables instead).

`nsgym_eval_draws` returns the draws made by the very device functions the step kernels call
(`Rng<R>::std_normal`, `reset_uniforms`, `sched_uniform`, `dyn_uniform`, `dirichlet_ones`,
`geometric_from_uniform`).  Two kinds of checks, at >= 2^22 samples:

* distributions: N(0, 1) moments and a Kolmogorov-Smirnov test for the fp32 (24-bit MUFU Box-Muller)
  and fp64 normals; uniformity of the 53-bit uniforms; Dirichlet(1, 1, 1) marginals = Beta(1, 2);
  Memoryless inter-fire times vs Geometric(p); slip-outcome frequencies of the step kernels vs p;
  independence across lanes / steps / envs
* value by value against the host restatement in tests/philox_np.py, which is what feeds the oracle
  in test_gpu_native_parity.py
"""
import ctypes as C

import numpy as np
import pytest

from tests import philox_np as PN

pytestmark = pytest.mark.gpu

N = 1 << 22
SEED = 20261018


def _handle(env_id="CartPole-v1", precision="fp32", seed=SEED, offset=0, **kw):
    import ns_gym_b200.schedulers as PS
    import ns_gym_b200.update_functions as PU
    from ns_gym_b200.vector_env import NSVectorEnv

    if "CartPole" in env_id:
        tp = {"gravity": PU.RandomWalk(PS.ContinuousScheduler())}
    else:
        tp = {"P": PU.DistributionNoUpdate(PS.ContinuousScheduler())}
    return NSVectorEnv(env_id, tp, 256, precision=precision, seed=seed, env_id_offset=offset, **kw)


def _draw(env, what, n=N, lane=0, t=0, p=0.0, step=1, planes=1):
    import torch

    from ns_gym_b200 import native as nv

    out = torch.zeros((planes, n), dtype=torch.float64, device=env.device)
    nv.check(env.lib.nsgym_eval_draws(env._h, int(what), int(lane), int(t), float(p), int(step),
                                      C.c_void_p(out.data_ptr()), int(n), env._stream()), "nsgym_eval_draws")
    torch.cuda.synchronize()
    return out.cpu().numpy()


def _check_standard_normal(z, name):
    from scipy import stats

    n = len(z)
    assert np.isfinite(z).all()
    m, v = z.mean(), z.var()
    assert abs(m) < 5.0 / np.sqrt(n), f"{name}: mean {m}"
    assert abs(v - 1.0) < 5.0 * np.sqrt(2.0 / n), f"{name}: variance {v}"
    skew = np.mean(z ** 3)
    kurt = np.mean(z ** 4)
    assert abs(skew) < 5.0 * np.sqrt(15.0 / n), f"{name}: third moment {skew}"
    assert abs(kurt - 3.0) < 5.0 * np.sqrt(96.0 / n), f"{name}: fourth moment {kurt}"
    d, pval = stats.kstest(z, "norm")
    assert pval > 1e-4, f"{name}: KS D={d} p={pval}"
    # tails: P(|z| > 4) = 6.33e-5
    tail = np.mean(np.abs(z) > 4.0)
    assert 0.5 * 6.33e-5 < tail < 1.6 * 6.33e-5, f"{name}: tail mass {tail}"


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_native_standard_normal_distribution(precision):
    from ns_gym_b200 import native as nv

    env = _handle(precision=precision)
    for lane, step in ((0, 1), (1, 1), (3, 7)):
        z = _draw(env, nv.DRAW_NORMAL, lane=lane, step=step)[0]
        _check_standard_normal(z, f"{precision} lane {lane} step {step}")
    # independence: lanes of one block, consecutive steps, neighbouring envs
    a = _draw(env, nv.DRAW_NORMAL, lane=0, step=3)[0]
    b = _draw(env, nv.DRAW_NORMAL, lane=1, step=3)[0]
    c = _draw(env, nv.DRAW_NORMAL, lane=0, step=4)[0]
    lim = 5.0 / np.sqrt(N)
    assert abs(np.corrcoef(a, b)[0, 1]) < lim
    assert abs(np.corrcoef(a, c)[0, 1]) < lim
    assert abs(np.corrcoef(a[:-1], a[1:])[0, 1]) < lim
    assert abs(np.corrcoef(a ** 2, b ** 2)[0, 1]) < lim


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_native_normal_equals_host_restatement(precision):
    """What test_gpu_native_parity.py feeds the oracle is what the device draws: fp64 to rounding of
    log / sqrt / cospi, fp32 to the MUFU approximations (lg2 / sqrt / cos, abs error <= 4e-6 at |z| <= 5.8)."""
    from ns_gym_b200 import native as nv

    env = _handle(precision=precision, offset=1 << 33)          # 64-bit global env ids
    n = 1 << 20
    gids = np.arange(n, dtype=np.uint64) + np.uint64(1 << 33)
    for lane, step in ((0, 1), (1, 2), (2, 5), (7, (1 << 32) + 3)):
        got = _draw(env, nv.DRAW_NORMAL, n=n, lane=lane, step=step)[0]
        want = PN.std_normal(gids, step, lane, SEED, precision)
        if precision == "fp64":
            np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-14)
        else:
            # MUFU lg2 has an ABSOLUTE error of ~2^-22 next to u1 = 1, which the square root magnifies for
            # the smallest radii (|z| < 0.02, one draw in ~10^4): 5e-5 there, 4e-6 for everything else
            err = np.abs(got - want)
            assert err.max() < 5e-5, err.max()
            assert (err > 4e-6).mean() < 2e-4, (err > 4e-6).mean()


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_native_reset_uniforms(precision):
    from scipy import stats

    from ns_gym_b200 import native as nv

    env = _handle(precision=precision)
    u = _draw(env, nv.DRAW_RESET_UNIFORMS, step=0, planes=4)
    gids = np.arange(N, dtype=np.uint64)
    assert np.array_equal(u, PN.reset_uniforms(gids, 0, SEED, precision))      # bit for bit
    assert (u >= 0).all() and (u < 1).all()
    for k in range(4):
        assert stats.kstest(u[k], "uniform").pvalue > 1e-4
        assert abs(u[k].mean() - 0.5) < 5.0 / np.sqrt(12 * N)
    assert abs(np.corrcoef(u[0], u[1])[0, 1]) < 5.0 / np.sqrt(N)
    assert abs(np.corrcoef(u[2], u[3])[0, 1]) < 5.0 / np.sqrt(N)


def test_native_slip_uniform_and_scheduler_uniform():
    from scipy import stats

    from ns_gym_b200 import native as nv

    grid = _handle("FrozenLake-v1", initial_prob_dist=[1, 0, 0])
    gids = np.arange(N, dtype=np.uint64)
    for step in (1, 2, 9):                                   # both halves of a step pair's block
        u = _draw(grid, nv.DRAW_DYN_UNIFORM, step=step)[0]
        assert np.array_equal(u, PN.dyn_uniform(gids, step, SEED))
        assert stats.kstest(u, "uniform").pvalue > 1e-4
    a, b = _draw(grid, nv.DRAW_DYN_UNIFORM, step=2)[0], _draw(grid, nv.DRAW_DYN_UNIFORM, step=3)[0]
    assert abs(np.corrcoef(a, b)[0, 1]) < 5.0 / np.sqrt(N)    # the two halves of one block
    # scheduler uniforms: keyed by the episode time (replay per episode) unless persistent_params
    cart = _handle()
    for lane, t in ((0, 0), (1, 17), (5, 499)):
        u = _draw(cart, nv.DRAW_SCHED_UNIFORM, lane=lane, t=t, step=123)[0]
        assert np.array_equal(u, PN.sched_uniform(gids, 123, lane, SEED, t=t))
        assert np.array_equal(u, _draw(cart, nv.DRAW_SCHED_UNIFORM, lane=lane, t=t, step=77)[0])
        assert stats.kstest(u, "uniform").pvalue > 1e-4
    pers = _handle(persistent_params=True)
    u = _draw(pers, nv.DRAW_SCHED_UNIFORM, lane=2, t=5, step=123)[0]
    assert np.array_equal(u, PN.sched_uniform(gids, 123, 2, SEED))
    assert not np.array_equal(u, _draw(pers, nv.DRAW_SCHED_UNIFORM, lane=2, t=5, step=124)[0])


def test_native_dirichlet_marginals():
    """RandomCategorical (distribution.py:37-38): Dirichlet(1, 1, 1) -> every component ~ Beta(1, 2),
    components sum to 1; CliffWalking's four outcomes: Beta(1, 3)."""
    from scipy import stats

    from ns_gym_b200 import native as nv

    gids = np.arange(1 << 18, dtype=np.uint64)
    for env_id, dim, ipd in (("FrozenLake-v1", 3, [1, 0, 0]), ("CliffWalking-v1", 4, [1, 0, 0, 0])):
        env = _handle(env_id, initial_prob_dist=ipd)
        p = _draw(env, nv.DRAW_DIRICHLET, lane=0, t=0, step=4, planes=dim)
        np.testing.assert_allclose(p.sum(0), 1.0, rtol=0, atol=1e-15)
        assert (p > 0).all()
        for k in range(dim):
            assert stats.kstest(p[k], "beta", args=(1, dim - 1)).pvalue > 1e-4, (env_id, k)
            assert abs(p[k].mean() - 1.0 / dim) < 5.0 * np.sqrt((dim - 1) / (dim * dim * (dim + 1.0)) / N)
        # pairwise correlation of Dirichlet(1,..,1) components: -1 / (dim - 1)
        assert abs(np.corrcoef(p[0], p[1])[0, 1] + 1.0 / (dim - 1)) < 5.0 / np.sqrt(N)
        want = PN.dirichlet_ones(gids, 4, 0, SEED, 0, dim)
        np.testing.assert_allclose(p[:, :len(gids)], want, rtol=1e-12)
        # a redraw (Lipschitz-bounded wrapper) is a different, equally distributed draw
        p2 = _draw(env, nv.DRAW_DIRICHLET, lane=0, t=3, step=4, planes=dim)
        assert abs(np.corrcoef(p[0], p2[0])[0, 1]) < 5.0 / np.sqrt(N)
        np.testing.assert_allclose(p2[:, :len(gids)], PN.dirichlet_ones(gids, 4, 0, SEED, 3, dim), rtol=1e-12)


@pytest.mark.parametrize("p", [0.25, 0.6, 0.02])
def test_native_memoryless_gaps_are_geometric(p):
    """MemorylessScheduler (schedulers.py:110-116): the next fire is t + Geometric(p) on {1, 2, ..}."""
    from scipy import stats

    from ns_gym_b200 import native as nv

    env = _handle()
    g = _draw(env, nv.DRAW_GEOMETRIC, lane=0, t=11, p=p, step=5)[0]
    assert (g >= 1).all() and np.array_equal(g, np.round(g))
    assert abs(g.mean() - 1.0 / p) < 5.0 * np.sqrt((1 - p) / p ** 2 / N)
    kmax = int(np.ceil(np.log(1e-4) / np.log1p(-p)))
    obs = np.bincount(g.astype(np.int64), minlength=kmax + 2)[1:kmax + 1].astype(np.float64)
    obs = np.append(obs, N - obs.sum())
    exp = N * np.append(stats.geom.pmf(np.arange(1, kmax + 1), p), stats.geom.sf(kmax, p))
    chi2 = ((obs - exp) ** 2 / exp).sum()
    assert stats.chi2.sf(chi2, len(obs) - 1) > 1e-4, (p, chi2)


@pytest.mark.parametrize("env_id,ipd,cell", [("FrozenLake-v1", [0.5, 0.3, 0.2], 9), ("ns_gym/Bridge-v0", [0.6, 0.25, 0.15], 20),
                                            ("CliffWalking-v1", [0.4, 0.3, 0.2, 0.1], 14)])
def test_step_kernel_slip_frequencies_follow_p(env_id, ipd, cell):
    """The LEAN gridworld step kernels on native draws: from an interior cell the outcomes of one action
    land on distinct cells, so next-cell frequencies are the slip distribution (chi-square at 2^22 envs)."""
    import torch
    from scipy import stats

    import ns_gym_b200.schedulers as PS
    import ns_gym_b200.update_functions as PU
    from ns_gym_b200 import native as nv
    from ns_gym_b200.vector_env import NSVectorEnv

    tp = {"P": PU.DistributionNoUpdate(PS.ContinuousScheduler(start=10 ** 6))}
    kw = {"map_name": "8x8"} if "Frozen" in env_id else {}
    env = NSVectorEnv(env_id, tp, N, autoreset="none", seed=3, initial_prob_dist=ipd, **kw)
    env.reset()
    ncol = {"FrozenLake-v1": 8, "ns_gym/Bridge-v0": 8, "CliffWalking-v1": 12}[env_id]
    for action in (0, 2):
        env.buffers["state"].fill_(cell)
        env.buffers["t"].fill_(nv.T_TABLE_FRESH if "Bridge" not in env_id else 0)
        if "Bridge" not in env_id:      # the sampling table holds the planes (reset leaves them as created)
            env.buffers["theta"].copy_(torch.tensor(ipd, dtype=torch.float64, device=env.device)[:, None].expand(-1, N))
        env.step_raw(torch.full((N,), action, dtype=torch.int32, device=env.device))
        torch.cuda.synchronize()
        assert env.lib.nsgym_last_kernel_class(env._h) == nv.KERNEL_LEAN_FAST
        nxt = env.buffers["state"].cpu().numpy().astype(np.int64)
        if "Cliff" in env_id:           # UP RIGHT DOWN LEFT
            delta = {0: -ncol, 1: 1, 2: ncol, 3: -1}
        else:                           # LEFT DOWN RIGHT UP
            delta = {0: -1, 1: ncol, 2: 1, 3: -ncol}
        dirs = [action, (action + 1) % 4, (action - 1) % 4, (action + 2) % 4][:len(ipd)]
        obs = np.array([(nxt == cell + delta[b]).sum() for b in dirs], dtype=np.float64)
        assert obs.sum() == N
        exp = N * np.array(ipd)
        chi2 = ((obs - exp) ** 2 / exp).sum()
        assert stats.chi2.sf(chi2, len(ipd) - 1) > 1e-4, (env_id, action, obs / N)


def test_fp32_box_muller_radius_is_finite_and_accurate_for_every_input():
    """All 2^24 values the radius bits can take (u1 = (k + 1) / 2^24): lg2.approx must not return a
    positive value next to u1 = 1 (sqrt of a negative number would put a NaN into theta), and the
    radius stays within the stated error of sqrt(-2 ln u1).  The angle bits are swept separately."""
    from ns_gym_b200 import native as nv

    env = _handle(precision="fp32")
    worst_abs, worst_abs_body, worst_rel_body = 0.0, 0.0, 0.0
    for first in range(0, 1 << 24, 1 << 22):
        r = _draw(env, nv.DRAW_BOX_MULLER_SWEEP, n=1 << 22, t=0, step=first)[0]      # angle 0: cos = 1
        k = np.arange(first, first + (1 << 22), dtype=np.float64)
        want = np.sqrt(-2.0 * np.log((k + 1.0) / 16777216.0))
        assert np.isfinite(r).all() and (r >= 0).all()
        err = np.abs(r - want)
        worst_abs = max(worst_abs, float(err.max()))
        body = want > 0.05                               # all but ~1.2e-3 of the draws
        worst_abs_body = max(worst_abs_body, float(err[body].max()))
        worst_rel_body = max(worst_rel_body, float((err[body] / want[body]).max()))
    import json
    import os
    if os.path.isdir("gpurun_out"):
        with open("gpurun_out/native_draws_report.json", "w") as f:
            json.dump({"fp32_box_muller_radius": {"worst_abs_all_2^24_inputs": worst_abs, "worst_abs_r>0.05": worst_abs_body,
                                                  "worst_rel_r>0.05": worst_rel_body}}, f)
    # MUFU.LG2's absolute error (~2^-22) is magnified by the square root for radii next to 0
    assert worst_abs < 5e-4, worst_abs
    assert worst_abs_body < 4e-6 and worst_rel_body < 4e-5, (worst_abs_body, worst_rel_body)
    # angle sweep at the largest radius (u1 = 2^-24, r = 5.768): cos via MUFU on [0, 2 pi)
    r0 = _draw(env, nv.DRAW_BOX_MULLER_SWEEP, n=1, t=0, step=0)[0][0]
    assert abs(r0 - np.sqrt(-2.0 * np.log(2.0 ** -24))) < 4e-6
    ts = np.unique(np.concatenate([np.arange(0, 1 << 24, 65521), [(1 << 22) - 1, 1 << 22, (1 << 23) - 1, 1 << 23,
                                                                    3 << 22, (1 << 24) - 1]]))
    got = np.array([_draw(env, nv.DRAW_BOX_MULLER_SWEEP, n=1, t=int(t), step=0)[0][0] for t in ts])
    ang = ts.astype(np.float32) * np.float32(6.283185307179586 / 16777216.0)
    worst = float(np.abs(got / r0 - np.cos(ang.astype(np.float64))).max())
    assert worst < 2e-6, worst                           # __cosf on [0, 2 pi)
