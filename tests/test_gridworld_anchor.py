"""The gridworld sampling rule (`categorical_sample`) and the clamped move, held to the reference's
OWN in-tree copies of them (rats-experiments/code/envs/nsfrozenlake_v0.py:61-68, 215-228;
nscliff_v0.py:40-47; nsbridge_v0.py:30-37) through known-answer vectors generated from those files
(tests/golden/make_gridworld_anchor.py).

gymnasium cannot be imported here or on the GPU box; for FrozenLake / CliffWalking this pins the two
pieces of gymnasium arithmetic the NS wrappers lean on (the table semantics
themselves are the reference's own, toy_text.py:426-469, 86-138, and pinned in test_oracle_vs_reference).
CPU: the oracle's restatement and the port's step.  GPU: the CUDA kernels' step, through the C ABI.
"""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "anchors", "gridworld_anchor.npz")


class _Rand:
    def __init__(self, u):
        self.u = u

    def random(self):
        return self.u


def test_oracle_categorical_sample_matches_the_reference_copies():
    from oracle.gym_restated import categorical_sample

    g = np.load(GOLDEN)
    got3 = np.array([categorical_sample(p, _Rand(float(u))) for p, u in zip(g["p3"], g["u"])])
    got4 = np.array([categorical_sample(p, _Rand(float(u))) for p, u in zip(g["p4"], g["u"])])
    assert np.array_equal(got3, g["idx3"])
    assert np.array_equal(got4, g["idx4"])
    # every outcome and the fall-through-to-0 case are exercised
    assert min(np.bincount(g["idx3"], minlength=3)) > 500 and min(np.bincount(g["idx4"], minlength=4)) > 500
    never = np.cumsum(g["p3"], 1)[:, -1] <= g["u"]
    assert never.sum() > 100 and (g["idx3"][never] == 0).all()


def _port_env(env_id, make, ipd):
    import ns_gym_b200.schedulers as PS
    import ns_gym_b200.update_functions as PU
    from oracle.ns_port import NSEnvPort

    tp = {"P": PU.DistributionNoUpdate(PS.ContinuousScheduler(start=10 ** 6))}
    env = NSEnvPort(env_id, tp, initial_prob_dist=ipd, **make)
    env.reset(seed=0)
    return env


@pytest.mark.parametrize("kind", ["frozenlake8", "frozenlake4", "cliff"])
def test_port_grid_step_matches_the_reference_copies(kind):
    """The port's table-driven step (the rule inline + its own move) on the anchor vectors."""
    g = np.load(GOLDEN)
    n = 768
    r = np.random.default_rng(5)
    if kind == "cliff":
        env = _port_env("CliffWalking-v1", {}, [1, 0, 0, 0])
        P, idx, move, n_cells = g["p4"], g["idx4"], g["move_4x12"], 48
        to_inc = [3, 2, 1, 0]                       # UP RIGHT DOWN LEFT -> LEFT DOWN RIGHT UP
        cells = r.integers(0, n_cells, n)
    else:
        size = 8 if kind == "frozenlake8" else 4
        env = _port_env("FrozenLake-v1", {"map_name": f"{size}x{size}"}, [1, 0, 0])
        P, idx, move, n_cells = g["p3"], g["idx3"], g[f"move_{size}x{size}"], size * size
        to_inc = [0, 1, 2, 3]
        desc = env.base.desc.ravel()
        free = np.array([c for c in range(n_cells) if bytes(desc[c]) not in (b"G", b"H")])
        cells = free[r.integers(0, len(free), n)]
    actions = r.integers(0, 4, n)
    pick = r.choice(len(idx), n, replace=False)
    for k in range(n):
        p, u, i = P[pick[k]], float(g["u"][pick[k]]), int(idx[pick[k]])
        env.table_prob = [float(x) for x in p]
        env.base.s = int(cells[k])
        env.base.np_random = _Rand(u)
        a = int(actions[k])
        ns, _, _, info = env._grid_step(a)
        b = [a, (a + 1) % 4, (a - 1) % 4, (a + 2) % 4][i]
        want = int(move[cells[k], to_inc[b]])
        if kind == "cliff" and want // 12 == 3 and 1 <= want % 12 <= 10:
            want = 36                               # toy_text.py:118-121: the cliff sends the agent back to start
        assert ns == want, (k, p, u, i, a, cells[k], ns, want)
        assert info["prob"] == env.table_prob[i]


@pytest.mark.skipif(not os.path.exists("/root/reference/ns_gym"), reason="needs the reference tree (build container)")
def test_anchor_vectors_come_from_the_reference_files():
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_gridworld_anchor", os.path.join(HERE, "golden", "make_gridworld_anchor.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    g = np.load(GOLDEN)
    p3, p4, u = mk.inputs()
    assert np.array_equal(p3, g["p3"]) and np.array_equal(p4, g["p4"]) and np.array_equal(u, g["u"])
    i3, i4, moves = mk.reference_outputs(p3[:1024], p4[:1024], u[:1024])
    assert np.array_equal(i3, g["idx3"][:1024]) and np.array_equal(i4, g["idx4"][:1024])
    for k, v in moves.items():
        assert np.array_equal(v, g[k])


# ---------------------------------------------------------------------------------------------
def _gpu_env(env_id, make, ipd, n, **kw):
    import ns_gym_b200.schedulers as PS
    import ns_gym_b200.update_functions as PU
    from ns_gym_b200.vector_env import NSVectorEnv

    tp = {"P": PU.DistributionNoUpdate(PS.ContinuousScheduler(start=10 ** 6))}
    env = NSVectorEnv(env_id, tp, n, autoreset="none", seed=0, initial_prob_dist=ipd, **make, **kw)
    env.reset(seed=0)
    return env


def _gpu_step(env, cells, probs, actions, u):
    """One kernel step from hand-set cells / table probabilities with the slip uniform injected."""
    import torch

    from ns_gym_b200 import native as nv
    from oracle import streams as S_

    n = env.num_envs
    dev = env.device
    env.buffers["state"].copy_(torch.as_tensor(cells, dtype=torch.int32, device=dev))
    env.buffers["theta"].copy_(torch.as_tensor(np.ascontiguousarray(probs.T), dtype=torch.float64, device=dev))
    # "table rebuilt since reset": the stored planes are what the env samples from
    env.buffers["t"].fill_(nv.T_TABLE_FRESH)
    U = torch.zeros((S_.n_uniform_lanes(1), n), dtype=torch.float64, device=dev)
    U[S_.LANE_DYN] = torch.as_tensor(u, dtype=torch.float64, device=dev)
    env.step_raw(torch.as_tensor(actions, dtype=torch.int32, device=dev), inject_uniform=U)
    torch.cuda.synchronize()
    return env.buffers["state"].cpu().numpy().astype(np.int64)


@pytest.mark.gpu
@pytest.mark.parametrize("size", [8, 4])
def test_kernel_frozenlake_step_matches_the_reference_copies(size):
    g = np.load(GOLDEN)
    n = len(g["u"])
    env = _gpu_env("FrozenLake-v1", {"map_name": f"{size}x{size}"}, [1, 0, 0], n)
    r = np.random.default_rng(11 + size)
    spec = env.program.spec
    terminal = int(spec.hole_mask) | int(spec.goal_mask)
    free = np.array([c for c in range(size * size) if not (terminal >> c) & 1])
    cells = free[r.integers(0, len(free), n)]
    actions = r.integers(0, 4, n)
    got = _gpu_step(env, cells, g["p3"], actions, g["u"])
    dirs = np.stack([actions, (actions + 1) % 4, (actions - 1) % 4], 1)
    want = g[f"move_{size}x{size}"][cells, dirs[np.arange(n), g["idx3"]]]
    assert np.array_equal(got, want), np.argwhere(got != want)[:5].tolist()


@pytest.mark.gpu
def test_kernel_cliffwalking_step_matches_the_reference_copies():
    g = np.load(GOLDEN)
    n = len(g["u"])
    env = _gpu_env("CliffWalking-v1", {}, [1, 0, 0, 0], n)
    r = np.random.default_rng(13)
    cells = r.integers(0, 48, n)
    actions = r.integers(0, 4, n)
    got = _gpu_step(env, cells, g["p4"], actions, g["u"])
    dirs = np.stack([actions, (actions + 1) % 4, (actions - 1) % 4, (actions + 2) % 4], 1)
    b = dirs[np.arange(n), g["idx4"]]
    to_inc = np.array([3, 2, 1, 0])                 # UP RIGHT DOWN LEFT (toy_text.py:74-76) -> LEFT DOWN RIGHT UP
    want = g["move_4x12"][cells, to_inc[b]]
    cliff = (want // 12 == 3) & (want % 12 >= 1) & (want % 12 <= 10)
    want = np.where(cliff, 36, want)
    assert np.array_equal(got, want), np.argwhere(got != want)[:5].tolist()


@pytest.mark.gpu
def test_kernel_bridge_move_matches_the_reference_copy():
    """Bridge samples by np.random.choice (reference-owned, envs/Bridge.py:95-97, pinned elsewhere);
    with P = [1, 0, 0] the step is the move alone: 'out of bounds -> stay' == the copies' clamp."""
    import ns_gym_b200.schedulers as PS
    import ns_gym_b200.update_functions as PU
    import torch

    from ns_gym_b200.vector_env import NSVectorEnv
    from oracle import streams as S_

    g = np.load(GOLDEN)
    n = 40 * 4
    tp = {"P": PU.DistributionNoUpdate(PS.ContinuousScheduler(start=10 ** 6))}
    env = NSVectorEnv("ns_gym/Bridge-v0", tp, n, autoreset="none", seed=0, initial_prob_dist=[1, 0, 0])
    env.reset(seed=0)
    cells, actions = np.repeat(np.arange(40), 4), np.tile(np.arange(4), 40)
    env.buffers["state"].copy_(torch.as_tensor(cells, dtype=torch.int32, device=env.device))
    U = torch.full((S_.n_uniform_lanes(1), n), 0.5, dtype=torch.float64, device=env.device)
    env.step_raw(torch.as_tensor(actions, dtype=torch.int32, device=env.device), inject_uniform=U)
    torch.cuda.synchronize()
    got = env.buffers["state"].cpu().numpy().astype(np.int64)
    assert np.array_equal(got, g["move_5x8"][cells, actions])
