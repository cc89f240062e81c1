"""GPU: explicit resets through the C ABI (nsgym_reset): masked resets touch only the selected envs
(NSWrapper.reset per env, base.py:365-431), persistent parameters survive, cursors rewind."""
import numpy as np
import pytest

from tests.cases import CASES

pytestmark = pytest.mark.gpu


def _env(name, n, **kw):
    from tests import parity_util as pu

    case = dict(CASES[name])
    case["wrapper"] = {**case["wrapper"], **kw.pop("wrapper", {})}
    env = pu.gpu_env(case, n, **kw)
    env.reset(seed=4)
    return env


def _step(env, k):
    for _ in range(k):
        env.step(env.action_space.sample())


@pytest.mark.parametrize("name", ["c1_cartpole_readme", "cartpole_lists", "c3_pendulum", "c2_frozenlake8_drift",
                                  "bridge_stepwise", "c4_cartpole_rows", "het_cartpole_wide"])
def test_masked_reset_touches_only_the_selected_envs(name):
    import torch

    n = 96
    env = _env(name, n, autoreset="none")
    _step(env, 7)
    before = {k: v.clone() for k, v in env.buffers.items() if v is not None}
    mask = torch.zeros(n, dtype=torch.bool, device=env.device)
    mask[::3] = True
    obs, info = env.reset(mask=mask)
    after = env.buffers
    keep = ~mask
    for key in ("state", "t"):
        assert torch.equal(after[key][keep], before[key][keep]), f"{name}: {key} of an unselected env changed"
    assert torch.equal(after["theta"][:, keep], before["theta"][:, keep])
    if after["istate"] is not None:
        assert torch.equal(after["istate"][:, keep], before["istate"][:, keep])
    # selected envs: t = 0, parameters and cursors back to their initial values
    assert int((after["t"][mask] & 0x0FFFFFFF).abs().sum()) == 0
    fresh = _env(name, n, autoreset="none")
    if env.program.is_grid and env.program.env_kind != 7:
        # FrozenLake / CliffWalking: the sampling table stays stale until the next fire (toy_text.py:
        # 365-367, 395-398), the wrapper's transition_prob is back to the initial distribution
        assert torch.equal(env.transition_prob()["P"][:, mask], fresh.transition_prob()["P"][:, mask])
        assert torch.equal(after["theta"][:, mask], before["theta"][:, mask])
    else:
        assert torch.equal(after["theta"][:, mask], fresh.buffers["theta"][:, mask])
    if after["istate"] is not None:
        assert torch.equal(after["istate"][:, mask], fresh.buffers["istate"][:, mask])
    assert bool((after["flags"][mask] == 4).all())                       # NSGYM_FLAG_RESET
    if not env.program.is_grid:                                          # a new initial state was drawn
        assert not torch.equal(after["state"][mask], before["state"][mask])
    # and the batch keeps stepping
    _step(env, 3)
    assert bool(((env.relative_time() == 3) == mask).all())


def test_masked_reset_with_persistent_params_keeps_theta_and_cursors():
    import torch

    n = 64
    env = _env("cartpole_persistent", n, autoreset="none")
    _step(env, 9)
    before = {k: v.clone() for k, v in env.buffers.items() if v is not None}
    mask = torch.arange(n, device=env.device) % 2 == 0
    env.reset(mask=mask)
    assert torch.equal(env.buffers["theta"], before["theta"])            # base.py:392-395
    assert torch.equal(env.buffers["istate"], before["istate"])
    assert int((env.buffers["t"][mask] & 0x0FFFFFFF).abs().sum()) == 0
    assert torch.equal(env.buffers["t"][~mask], before["t"][~mask])


def test_first_reset_must_cover_the_batch():
    import torch

    import ns_gym_b200.schedulers as PS
    import ns_gym_b200.update_functions as PU
    from ns_gym_b200 import native as nv
    from ns_gym_b200.vector_env import NSVectorEnv

    env = NSVectorEnv("CartPole-v1", {"masspole": PU.IncrementUpdate(PS.ContinuousScheduler(), k=0.1)}, 32)
    with pytest.raises(nv.NsgymError):
        env.step_raw(torch.zeros(32, dtype=torch.int32, device=env.device))       # step before reset
    with pytest.raises(nv.NsgymError):
        env.reset(mask=torch.ones(32, dtype=torch.bool, device=env.device))       # first reset needs all envs
    env.reset(seed=1)
    env.step_raw(torch.zeros(32, dtype=torch.int32, device=env.device))


def test_seeded_resets_replay_and_differ_across_seeds():
    import torch

    a, b, c = (_env("c1_cartpole_readme", 256, precision="fp32") for _ in range(3))
    a.reset(seed=10); b.reset(seed=10); c.reset(seed=11)
    assert torch.equal(a.buffers["state"], b.buffers["state"])
    assert not torch.equal(a.buffers["state"], c.buffers["state"])


def test_slot_streams_follow_the_spawn_structure_of_the_reference():
    """NSWrapper._seed_update_fns (base.py:412-421) gives update function j the child
    SeedSequence(seed).spawn(P)[j]: its stream depends on (seed, j) only -- not on how many other
    parameters are bound, nor on what they are.  Here the stream of slot j is Philox block j >> 1,
    half j & 1 under the key `seed`: binding a third parameter, or swapping the rule of another slot,
    leaves the draws of slot j unchanged; reset(seed) replays them, reset() lets them run on
    (base.py:391-393: the generators are transplanted, not re-seeded)."""
    import torch

    import ns_gym_b200.schedulers as S
    import ns_gym_b200.update_functions as U
    from ns_gym_b200.vector_env import NSVectorEnv

    def run(params, seed, steps=6, second_reset=None):
        # (6 steps: no episode ends yet -- a termination resets theta, and when an env terminates depends
        # on every bound parameter through the dynamics)
        env = NSVectorEnv("CartPole-v1", params, 512, precision="fp64", seed=0)
        env.reset(seed=seed)
        a = torch.zeros(512, dtype=torch.int32, device=env.device)
        out = []
        for k in range(steps):
            if second_reset is not None and k == steps // 2:
                env.reset(**second_reset)
            env.step_raw(a)
            assert int((env.buffers["flags"] & 7).max()) == 0, "an episode ended inside the window"
            out.append(env.buffers["theta"].clone())
        return env.keys, out

    rw = lambda: U.RandomWalk(S.ContinuousScheduler(), mu=0.0, sigma=0.01)  # noqa: E731
    k2, two = run({"gravity": rw(), "masscart": rw()}, seed=5)
    k3, three = run({"gravity": rw(), "masscart": rw(), "length": rw()}, seed=5)
    kx, other = run({"gravity": rw(), "masscart": U.IncrementUpdate(S.ContinuousScheduler(), k=0.001)}, seed=5)
    for k in range(6):
        assert torch.equal(two[k][0], three[k][0]) and torch.equal(two[k][1], three[k][1])   # slots 0, 1 unchanged
        assert torch.equal(two[k][0], other[k][0])                                             # slot 0 unchanged
    assert not torch.equal(three[3][2] - 0.5, three[3][1] - 1.0)                               # slot 2 has its own stream
    # another seed: every slot's stream changes
    _, reseeded = run({"gravity": rw(), "masscart": rw()}, seed=6)
    assert not torch.equal(two[0][0], reseeded[0][0])
    # reset(seed) in mid-run replays the stream from its start; reset() continues it
    _, replay = run({"gravity": rw(), "masscart": rw()}, seed=5, second_reset=dict(seed=5))
    _, cont = run({"gravity": rw(), "masscart": rw()}, seed=5, second_reset=dict())
    first_draw = two[0][0] - 9.8
    assert torch.allclose(replay[3][0] - 9.8, first_draw, rtol=0, atol=1e-12)
    assert not torch.allclose(cont[3][0] - 9.8, first_draw, rtol=0, atol=1e-12)


def test_start_cells_of_a_multi_start_map_are_sampled_uniformly():
    """FrozenLake map with three 'S' cells (toy_text.py:314-319 accepts any desc): explicit resets and next-step
    autoresets draw the start cell with categorical_sample over the start cells -- uniform frequencies, only
    start cells, independent of the previous cell; same cells from the specialised kernel."""
    import numpy as np
    import torch

    n = 1 << 16
    cells = {}
    for specialize in (0, 1):
        env = _env("frozenlake5_multi_start", n, precision="fp64")
        env.set_option("specialize", specialize)
        env.reset(seed=4)
        first = env.buffers["state"].clone()
        counts = np.bincount(first.cpu().numpy(), minlength=25)
        assert set(np.nonzero(counts)[0]) == {0, 4, 20}
        assert np.all(np.abs(counts[[0, 4, 20]] / n - 1 / 3) < 0.01), counts[[0, 4, 20]]
        a = torch.ones(n, dtype=torch.int32, device=env.device)
        seen = []
        for _ in range(40):
            env.step_raw(a)
            was_reset = (env.buffers["flags"] & 4) != 0
            seen.append(env.buffers["state"][was_reset].clone())
        again = torch.cat(seen).cpu().numpy()
        c2 = np.bincount(again, minlength=25)
        assert len(again) > n and set(np.nonzero(c2)[0]) == {0, 4, 20}
        assert np.all(np.abs(c2[[0, 4, 20]] / len(again) - 1 / 3) < 0.01), c2[[0, 4, 20]]
        cells[specialize] = (first, again)
    assert torch.equal(cells[0][0], cells[1][0]) and np.array_equal(cells[0][1], cells[1][1])
