"""GPU: the reference's own wrapper tests (tests/test_step_reset.py, test_classic_control_wrapper.py,
test_gridworld_wrappers.py of scope-lab-vu/ns_gym), re-asked of the batched drop-in through the same
plugin surface: same constructors, same keyword arguments, same observation / info keys, same
assertions -- values are tensors with one entry per env instead of Python scalars.
Reference line numbers in each test refer to /root/reference/tests/."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CLASSIC_CONTROL_ENV_IDS = ["CartPole-v1", "Acrobot-v1", "MountainCar-v0", "MountainCarContinuous-v0", "Pendulum-v1"]
GRIDWORLD_ENV_IDS = ["CliffWalking-v1", "FrozenLake-v1"]
OBS_KEYS = ["state", "env_change", "delta_change", "relative_time"]
N = 6


def _api():
    import ns_gym_b200 as nsb
    from ns_gym_b200.base import TUNABLE_PARAMS, Reward
    from ns_gym_b200.schedulers import ContinuousScheduler, PeriodicScheduler
    from ns_gym_b200.update_functions import DistributionIncrementUpdate, IncrementUpdate, RandomWalk
    from ns_gym_b200.wrappers import (NSBridgeWrapper, NSClassicControlWrapper, NSCliffWalkingWrapper,
                                      NSFrozenLakeWrapper)
    return locals()


@pytest.fixture
def api():
    return _api()


@pytest.fixture
def cc_params(api):                                        # test_step_reset.py:35-45
    fn = api["IncrementUpdate"](api["ContinuousScheduler"](), k=0.1)
    dec_fn = api["IncrementUpdate"](api["ContinuousScheduler"](), k=-0.1)
    return {"CartPole-v1": {"masspole": fn, "gravity": fn}, "Acrobot-v1": {"LINK_LENGTH_1": fn, "LINK_MASS_2": fn},
            "MountainCar-v0": {"gravity": dec_fn, "force": fn}, "MountainCarContinuous-v0": {"power": fn},
            "Pendulum-v1": {"m": fn, "g": fn}}


@pytest.fixture
def gw(api):                                               # test_step_reset.py:48-63
    fn = api["DistributionIncrementUpdate"](api["ContinuousScheduler"](), k=-0.1)
    return {"CliffWalking-v1": (api["NSCliffWalkingWrapper"], {"P": fn}, dict(initial_prob_dist=[1, 0, 0, 0])),
            "FrozenLake-v1": (api["NSFrozenLakeWrapper"], {"P": fn}, dict(initial_prob_dist=[1, 0, 0]))}


def _sample(env):
    a = env.action_space.sample()                 # one action per env, as ns_env.action_space.sample()
    assert env.action_space.contains(a) and a.shape[0] == env.num_envs
    return a


def _make_cc(api, env_id, params, **kw):
    return api["NSClassicControlWrapper"](api["nsb"].make(env_id, num_envs=N), params, precision="fp64",
                                          autoreset="none", **kw)


@pytest.mark.parametrize("env_id", CLASSIC_CONTROL_ENV_IDS)
def test_reset_restores_all_params_classic_control(api, cc_params, env_id):      # test_step_reset.py:70-96
    ns_env = _make_cc(api, env_id, cc_params[env_id])
    ns_env.reset(seed=42)
    defaults = ns_env.get_default_params()
    for _ in range(10):
        ns_env.step(_sample(ns_env))
    assert any(not np.allclose(v.cpu().numpy(), defaults[p]) for p, v in ns_env.theta().items())
    ns_env.reset(seed=42)
    for p, v in ns_env.theta().items():
        assert np.allclose(v.cpu().numpy(), defaults[p]), f"Param '{p}' not restored after reset"


@pytest.mark.parametrize("env_id", GRIDWORLD_ENV_IDS)
def test_reset_restores_all_params_gridworld(api, gw, env_id):                   # test_step_reset.py:99-117
    cls, params, kw = gw[env_id]
    ns_env = cls(api["nsb"].make(env_id, num_envs=N), params, autoreset="none", **kw)
    ns_env.reset(seed=42)
    initial_tp = ns_env.transition_prob()["P"].clone()
    for _ in range(5):
        ns_env.step(_sample(ns_env))
    assert not np.array_equal(ns_env.transition_prob()["P"].cpu().numpy(), initial_tp.cpu().numpy())
    ns_env.reset(seed=42)
    assert np.array_equal(ns_env.transition_prob()["P"].cpu().numpy(), initial_tp.cpu().numpy())


@pytest.mark.parametrize("env_id", CLASSIC_CONTROL_ENV_IDS)
def test_multiple_resets_classic_control(api, cc_params, env_id):                # test_step_reset.py:124-154
    ns_env = _make_cc(api, env_id, cc_params[env_id])
    for cycle in range(3):
        obs, info = ns_env.reset(seed=42)
        assert isinstance(obs, dict) and all(k in obs for k in OBS_KEYS)
        assert int(obs["relative_time"].abs().sum()) == 0, f"t != 0 after reset cycle {cycle}"
        for p in cc_params[env_id]:
            assert int(obs["env_change"][p].sum()) == 0 and float(obs["delta_change"][p].abs().sum()) == 0.0
        for p, v in ns_env.theta().items():
            assert np.allclose(v.cpu().numpy(), ns_env.get_default_params()[p])
        for _ in range(5):
            ns_env.step(_sample(ns_env))
    assert isinstance(info["Ground Truth Env Change"], dict) and isinstance(info["Ground Truth Delta Change"], dict)


@pytest.mark.parametrize("env_id", CLASSIC_CONTROL_ENV_IDS)
def test_step_increments_t_classic_control(api, cc_params, env_id):              # test_step_reset.py:160-172
    ns_env = _make_cc(api, env_id, cc_params[env_id])
    ns_env.reset(seed=42)
    for step_num in range(1, 6):
        obs, _, done, trunc, _ = ns_env.step(_sample(ns_env))
        assert (obs["relative_time"] == step_num).all()
        assert (ns_env.t == step_num).all(), f"Expected t={step_num}"


@pytest.mark.parametrize("env_id", GRIDWORLD_ENV_IDS)
def test_step_increments_t_gridworld(api, gw, env_id):                           # test_step_reset.py:175-186
    cls, params, kw = gw[env_id]
    ns_env = cls(api["nsb"].make(env_id, num_envs=N), params, autoreset="none", **kw)
    ns_env.reset(seed=42)
    for step_num in range(1, 6):
        obs, *_ = ns_env.step(_sample(ns_env))
        assert (obs["relative_time"] == step_num).all()


def test_step_updates_params_by_known_amount(api):                               # test_step_reset.py:193-215
    import torch

    k = 0.5
    fn = api["IncrementUpdate"](api["ContinuousScheduler"](start=0), k=k)
    ns_env = _make_cc(api, "CartPole-v1", {"masspole": fn})
    ns_env.reset(seed=42)
    s0 = ns_env.buffers["state"].clone()
    ns_env.step(torch.zeros(N, dtype=torch.int32))
    assert np.allclose(ns_env.theta()["masspole"].cpu().numpy(), 0.1 + k)
    # the dependency resolver ran: the step used total_mass = masspole + masscart and
    # polemass_length = masspole * length of the UPDATED masspole (classic_control.py:424-444)
    x, x_dot, th, th_dot = (s0[:, i].cpu().numpy() for i in range(4))
    masspole, masscart, length, g, tau, force = 0.1 + k, 1.0, 0.5, 9.8, 0.02, -10.0
    total_mass, pml = masspole + masscart, masspole * length
    temp = (force + pml * th_dot ** 2 * np.sin(th)) / total_mass
    thacc = (g * np.sin(th) - np.cos(th) * temp) / (length * (4.0 / 3.0 - masspole * np.cos(th) ** 2 / total_mass))
    xacc = temp - pml * thacc * np.cos(th) / total_mass
    want = np.stack([x + tau * x_dot, x_dot + tau * xacc, th + tau * th_dot, th_dot + tau * thacc], 1)
    assert np.allclose(ns_env.buffers["state"].cpu().numpy(), want, rtol=1e-12, atol=1e-14)


def test_step_notification_matrix(api, cc_params):                               # test_step_reset.py:222-285
    import torch

    a = torch.zeros(N, dtype=torch.int32)
    params = cc_params["CartPole-v1"]
    env = _make_cc(api, "CartPole-v1", params, change_notification=False, delta_change_notification=False)
    env.reset(seed=42)
    obs, *_ = env.step(a)
    for p in params:
        assert int(obs["env_change"][p].sum()) == 0 and float(obs["delta_change"][p].abs().sum()) == 0.0
    env = _make_cc(api, "CartPole-v1", params, change_notification=True, delta_change_notification=False)
    env.reset(seed=42)
    obs, _, _, _, info = env.step(a)
    assert any(int(obs["env_change"][p].sum()) for p in params), "env_change should reflect actual changes"
    for p in params:
        assert float(obs["delta_change"][p].abs().sum()) == 0.0
    env = _make_cc(api, "CartPole-v1", params, change_notification=True, delta_change_notification=True)
    env.reset(seed=42)
    obs, _, _, _, info = env.step(a)
    for p in params:
        assert (obs["env_change"][p] == 1).all() and np.allclose(obs["delta_change"][p].cpu().numpy(), 0.1)
        assert (info["Ground Truth Env Change"][p] == 1).all()
    with pytest.raises(AssertionError):                                          # base.py:252-255
        _make_cc(api, "CartPole-v1", params, change_notification=False, delta_change_notification=True)


def test_ground_truth_is_always_in_info(api, cc_params):                         # test_step_reset.py:358-372
    import torch

    env = _make_cc(api, "CartPole-v1", cc_params["CartPole-v1"], want_delta=True)
    env.reset(seed=42)
    obs, _, _, _, info = env.step(torch.zeros(N, dtype=torch.int32))
    for p in cc_params["CartPole-v1"]:
        assert int(obs["env_change"][p].sum()) == 0
        assert (info["Ground Truth Env Change"][p] == 1).all()
        assert np.allclose(info["Ground Truth Delta Change"][p].cpu().numpy(), 0.1)


def test_constraint_checker_prevents_invalid_values(api):                        # test_step_reset.py:379-397
    import torch

    fn = api["IncrementUpdate"](api["ContinuousScheduler"](start=0), k=-100.0)
    ns_env = _make_cc(api, "CartPole-v1", {"masscart": fn}, change_notification=True)
    ns_env.reset(seed=42)
    obs, _, _, _, info = ns_env.step(torch.zeros(N, dtype=torch.int32))
    assert (ns_env.theta()["masscart"] > 0).all(), "Constraint checker failed"
    assert np.allclose(ns_env.theta()["masscart"].cpu().numpy(), 1.0)
    assert int(info["Ground Truth Env Change"]["masscart"].sum()) == 0          # rejected: flag 0


def test_persistent_params_preserves_values_classic_control(api):               # test_step_reset.py:436-461
    fn = api["IncrementUpdate"](api["ContinuousScheduler"](), k=0.5)
    ns_env = _make_cc(api, "CartPole-v1", {"masspole": fn}, persistent_params=True)
    ns_env.reset(seed=42)
    for _ in range(5):
        ns_env.step(_sample(ns_env))
    mutated = ns_env.theta()["masspole"].clone()
    assert not np.allclose(mutated.cpu().numpy(), 0.1)
    obs, _ = ns_env.reset(seed=42)
    assert int(obs["relative_time"].abs().sum()) == 0, "Time should still reset to 0"
    assert np.array_equal(ns_env.theta()["masspole"].cpu().numpy(), mutated.cpu().numpy())


def test_persistent_params_rng_continuity_classic_control(api):                 # test_step_reset.py:464-487
    import torch

    fn = api["RandomWalk"](api["ContinuousScheduler"](), mu=0, sigma=0.01, seed=42)
    ns_env = _make_cc(api, "CartPole-v1", {"masspole": fn}, persistent_params=True)
    ns_env.reset(seed=0)
    a = torch.zeros(N, dtype=torch.int32)
    traj_a, traj_b = [], []
    for _ in range(5):
        ns_env.step(a)
        traj_a.append(ns_env.theta()["masspole"].clone())
    ns_env.reset()
    for _ in range(5):
        ns_env.step(a)
        traj_b.append(ns_env.theta()["masspole"].clone())
    assert not all(torch.equal(x, y) for x, y in zip(traj_a, traj_b))


def test_freeze_mutes_notifications_not_parameters(api, cc_params):             # base.py:443-451 vs classic_control.py:70-97
    import torch

    env = _make_cc(api, "CartPole-v1", cc_params["CartPole-v1"], change_notification=True)
    env.reset(seed=1)
    env.freeze()
    obs, _, _, _, info = env.step(torch.zeros(N, dtype=torch.int32))
    assert all(int(obs["env_change"][p].sum()) == 0 for p in obs["env_change"])
    assert np.allclose(env.theta()["masspole"].cpu().numpy(), 0.2)             # theta moved all the same
    env.unfreeze()
    obs, *_ = env.step(torch.zeros(N, dtype=torch.int32))
    assert all((obs["env_change"][p] == 1).all() for p in obs["env_change"])
    with pytest.raises(TypeError):
        env.freeze("yes")


def test_reward_dataclass_when_not_scalar(api, cc_params):                       # base.py:33-47, 343-352
    import torch

    env = _make_cc(api, "CartPole-v1", cc_params["CartPole-v1"], change_notification=True, scalar_reward=False)
    env.reset(seed=1)
    obs, reward, *_ = env.step(torch.zeros(N, dtype=torch.int32))
    assert isinstance(reward, api["Reward"])
    assert (reward.reward == 1).all() and (reward.relative_time == 1).all()
    assert set(reward.env_change) == set(cc_params["CartPole-v1"])


def test_unknown_parameter_and_wrong_env_are_rejected(api):                      # base.py:257-261, classic_control.py:36-38
    fn = api["IncrementUpdate"](api["ContinuousScheduler"](), k=0.1)
    with pytest.raises(AssertionError):
        _make_cc(api, "CartPole-v1", {"not_a_parameter": fn})
    with pytest.raises(AssertionError):
        api["NSClassicControlWrapper"](api["nsb"].make("FrozenLake-v1", num_envs=N), {"P": fn})
    with pytest.raises(AssertionError):                                          # toy_text.py:329-334
        api["NSFrozenLakeWrapper"](api["nsb"].make("FrozenLake-v1", num_envs=N),
                                   {"P": api["DistributionIncrementUpdate"](api["ContinuousScheduler"](), k=0.1)},
                                   initial_prob_dist=[0.5, 0.2, 0.2])


def test_probability_mass_is_conserved(api):                                     # test_gridworld_wrappers.py:222-229
    fn = api["DistributionIncrementUpdate"](api["ContinuousScheduler"](), k=-0.07)
    env = api["NSFrozenLakeWrapper"](api["nsb"].make("FrozenLake-v1", num_envs=N), {"P": fn},
                                     initial_prob_dist=[1, 0, 0], autoreset="none")
    env.reset(seed=3)
    for _ in range(8):
        env.step(_sample(env))
        assert np.allclose(env.transition_prob()["P"].sum(0).cpu().numpy(), 1.0)
