"""Shared parity cases.  Each case builds its ``tunable_params`` from a (schedulers,
update_functions) namespace pair, so the same text runs on the reference's classes, on the
oracle port and on the ns_gym_b200 descriptions compiled for the GPU.

C1..C5 are the BASELINE.json configs; the rest widen opcode coverage.
"""
from __future__ import annotations

MAP8 = {"map_name": "8x8"}


import numpy as np


def _c(env_id, params, wrapper=None, make=None, steps=60, fp32_rtol=None):
    return dict(env_id=env_id, params=params, wrapper=wrapper or {}, make=make or {}, steps=steps)


def _het(env_id, params_of, wrapper=None, make=None, steps=60):
    """Heterogeneous case (BASELINE config C4): ``params_of(S, U, e)`` builds the tunable_params
    of env e; ``params`` (env 0) keeps the homogeneous helpers working."""
    return dict(env_id=env_id, params=lambda S, U: params_of(S, U, 0), params_of=params_of,
                wrapper=wrapper or {}, make=make or {}, steps=steps)


def params_for(case, S, U, e):
    """tunable_params of env ``e`` of ``case`` (every env alike unless the case is heterogeneous)."""
    return case["params_of"](S, U, e) if "params_of" in case else case["params"](S, U)


# ---- C4: per-env opcode rows drawn with default_rng(4) --------------------------
def _c4_scheduler(S, r):
    k = int(r.integers(0, 5))
    if k == 0:
        return S.ContinuousScheduler()
    if k == 1:
        return S.PeriodicScheduler(int(r.integers(2, 8)))
    if k == 2:
        cycle = int(r.integers(2, 7))
        return S.BurstScheduler(int(r.integers(1, cycle + 1)), cycle)
    if k == 3:
        a, b = int(r.integers(0, 10)), int(r.integers(12, 40))
        return S.WindowScheduler([(a, a + int(r.integers(1, 6))), (b, b + int(r.integers(1, 9)))])
    return S.DiscreteScheduler({int(x) for x in r.integers(0, 60, size=int(r.integers(1, 9)))})


def _c4_scalar(S, U, r, y0, scale):
    """one of {Increment, Decrement, Trend, Geometric, LinearInterp, RandomWalk}, coefficients per env"""
    sch = _c4_scheduler(S, r)
    k = int(r.integers(0, 6))
    if k == 0:
        return U.IncrementUpdate(sch, k=float(r.uniform(0.1, 1.0)) * scale)
    if k == 1:
        return U.DecrementUpdate(sch, k=float(r.uniform(0.01, 0.2)) * scale)
    if k == 2:
        return U.DeterministicTrend(sch, slope=float(r.uniform(-0.01, 0.02)) * scale)
    if k == 3:
        return U.GeometricProgression(sch, r=float(r.uniform(0.98, 1.03)))
    if k == 4:
        return U.LinearInterpolation(sch, y0 * float(r.uniform(0.8, 1.0)), y0 * float(r.uniform(1.0, 1.5)),
                                     T=int(r.integers(20, 200)))
    return U.RandomWalk(sch, mu=float(r.uniform(-0.02, 0.02)) * scale, sigma=float(r.uniform(0.0, 0.3)) * scale)


def _het_cartpole_lean(S, U, e):
    """rows of the fast class only (any precision): the lean heterogeneous kernel"""
    r = np.random.default_rng([7, e])

    def one(scale):
        sch = _c4_scheduler(S, r)
        k = int(r.integers(0, 5))
        if k == 0:
            return U.IncrementUpdate(sch, k=float(r.uniform(0.1, 1.0)) * scale)
        if k == 1:
            return U.DecrementUpdate(sch, k=float(r.uniform(0.01, 0.2)) * scale)
        if k == 2:
            return U.DeterministicTrend(sch, slope=float(r.uniform(-0.01, 0.02)) * scale)
        if k == 3:
            return U.GeometricProgression(sch, r=float(r.uniform(0.98, 1.03)))
        return U.RandomWalkWithDrift(sch, alpha=0.01 * scale, mu=0.0, sigma=float(r.uniform(0.0, 0.3)) * scale)

    return {"force_mag": one(1.0), "length": one(0.02), "gravity": one(0.5)}


def _c4_cartpole(S, U, e):
    r = np.random.default_rng([4, e])
    return {"masspole": _c4_scalar(S, U, r, 0.1, 0.01), "gravity": _c4_scalar(S, U, r, 9.8, 0.5)}


def _c4_frozenlake(S, U, e):
    r = np.random.default_rng([4, 1 << 20, e])
    sch = _c4_scheduler(S, r)
    k = int(r.integers(0, 6))
    if k == 0:
        fn = U.DistributionDecrementUpdate(sch, k=float(r.uniform(0.01, 0.08)))
    elif k == 1:
        fn = U.DistributionIncrementUpdate(sch, k=float(r.uniform(0.0, 0.05)))
    elif k == 2:
        fn = U.UniformDrift(sch, rate=float(r.uniform(0.01, 0.2)))
    elif k == 3:
        tgt = r.dirichlet([2.0, 1.0, 1.0])
        fn = U.TargetReversion(sch, target=[float(x) for x in tgt], theta=float(r.uniform(0.05, 0.4)))
    elif k == 4:
        end = r.dirichlet([1.0, 1.0, 1.0])
        fn = U.DistributionLinearInterpolation(sch, [1.0, 0.0, 0.0], [float(x) for x in end],
                                               T=int(r.integers(10, 80)))
    else:
        vals = [[float(x) for x in r.dirichlet([3.0, 1.0, 1.0])] for _ in range(int(r.integers(1, 4)))]
        fn = U.DistributionStepWiseUpdate(sch, vals)
    return {"P": fn}


def _het_cartpole_wide(S, U, e):
    """every scalar opcode and every scheduler, mixed per env (cursors, stochastic schedulers, slow rules)"""
    r = np.random.default_rng([5, e])
    scheds = [
        lambda: S.ContinuousScheduler(start=int(r.integers(0, 4)), end=int(r.integers(20, 70))),
        lambda: S.PeriodicScheduler(int(r.integers(1, 5))),
        lambda: S.RandomScheduler(probability=float(r.uniform(0.2, 0.9)), seed=e),
        lambda: S.DecayingProbabilityScheduler(float(r.uniform(0.5, 1.0)), float(r.uniform(0.01, 0.1)), seed=e),
        lambda: S.MemorylessScheduler(p=float(r.uniform(0.2, 0.6)), seed=e),
        lambda: S.BurstScheduler(2, 5),
    ]

    def sch():
        return scheds[int(r.integers(0, len(scheds)))]()

    def upd(y0, scale, allow_cursor):
        k = int(r.integers(0, 9 if allow_cursor else 7))
        if k == 0:
            return U.OrnsteinUhlenbeck(sch(), theta=float(r.uniform(0.05, 0.3)), mu=y0, sigma=float(r.uniform(0, 0.05)) * scale)
        if k == 1:
            return U.BoundedRandomWalk(sch(), mu=0.0, sigma=0.2 * scale, lo=0.8 * y0, hi=1.2 * y0)
        if k == 2:
            return U.SigmoidTransition(sch(), a=y0, b=y0 * 1.3, k=float(r.uniform(0.1, 1.0)), t0=int(r.integers(5, 30)))
        if k == 3:
            return U.PolynomialTrend(sch(), coeffs=[1e-3 * scale, float(r.uniform(-1, 1)) * 1e-5 * scale])
        if k == 4:
            return U.OscillatingUpdate(sch(), delta=float(r.uniform(0.01, 0.1)) * scale)
        if k == 5:
            return U.ExponentialDecay(sch(), decay_rate=float(r.uniform(1e-4, 1e-3)))
        if k == 6:
            return U.RandomWalkWithDriftAndTrend(sch(), alpha=0.001 * scale, mu=0.0, sigma=0.05 * scale, slope=1e-4 * scale)
        s2 = S.PeriodicScheduler(int(r.integers(1, 4)))       # cursor updates: deterministic scheduler
        vals = [y0 * float(r.uniform(0.7, 1.4)) for _ in range(int(r.integers(1, 5)))]
        return U.StepWiseUpdate(s2, vals) if k == 7 else U.CyclicUpdate(s2, vals)

    return {"length": upd(0.5, 0.5, True), "force_mag": upd(10.0, 1.0, False), "masscart": upd(1.0, 1.0, True)}


def _het_bridge(S, U, e):
    r = np.random.default_rng([6, e])
    left = U.UniformDrift(S.PeriodicScheduler(int(r.integers(1, 4))), rate=float(r.uniform(0.01, 0.1)))
    if int(r.integers(0, 2)):
        right = U.DistributionDecrementUpdate(S.ContinuousScheduler(), k=float(r.uniform(0.005, 0.03)))
    else:
        right = U.DistributionCyclicUpdate(S.BurstScheduler(1, 3), [[0.8, 0.1, 0.1], [0.5, 0.25, 0.25]])
    return {"P_left": left, "P_right": right}


CASES = {
    # ---- BASELINE configs --------------------------------------------------------------
    "c1_cartpole_readme": _c(
        "CartPole-v1",
        lambda S, U: {
            "masspole": U.IncrementUpdate(S.ContinuousScheduler(), k=0.1),
            "gravity": U.RandomWalk(S.PeriodicScheduler(period=3)),
        },
        wrapper=dict(change_notification=True), steps=120),
    "c2_frozenlake8_stepchange": _c(
        "FrozenLake-v1",
        lambda S, U: {"P": U.DistributionStepWiseUpdate(S.DiscreteScheduler({12}), [[0.0, 0.5, 0.5]])},
        wrapper=dict(initial_prob_dist=[1, 0, 0], change_notification=True,
                     delta_change_notification=True),
        make=dict(MAP8, max_episode_steps=200), steps=260),
    "c2_frozenlake8_drift": _c(
        "FrozenLake-v1",
        lambda S, U: {"P": U.DistributionDecrementUpdate(S.ContinuousScheduler(), k=0.05)},
        wrapper=dict(initial_prob_dist=[1, 0, 0], change_notification=True,
                     delta_change_notification=True),
        make=dict(MAP8, max_episode_steps=200), steps=120),
    "c3_acrobot": _c(
        "Acrobot-v1",
        lambda S, U: {
            "LINK_MASS_2": U.GeometricProgression(S.ContinuousScheduler(), r=1.001),
            "LINK_LENGTH_1": U.IncrementUpdate(S.PeriodicScheduler(5), k=0.01),
        },
        wrapper=dict(change_notification=True, delta_change_notification=True), steps=40),
    "c3_mountaincar": _c(
        "MountainCar-v0",
        lambda S, U: {
            "gravity": U.LinearInterpolation(S.ContinuousScheduler(), 0.0025, 0.0035, T=200),
            "force": U.IncrementUpdate(S.ContinuousScheduler(), k=1e-5),
        },
        wrapper=dict(change_notification=True, delta_change_notification=True), steps=230),
    "c3_pendulum": _c(
        "Pendulum-v1",
        lambda S, U: {
            "m": U.IncrementUpdate(S.ContinuousScheduler(), k=0.01),
            "g": U.OscillatingUpdate(S.ContinuousScheduler(), delta=0.1),
        },
        wrapper=dict(change_notification=True, delta_change_notification=True), steps=220),
    "c5_bridge_uniform": _c(
        "ns_gym/Bridge-v0",
        lambda S, U: {"P": U.UniformDrift(S.ContinuousScheduler(), rate=0.05)},
        wrapper=dict(initial_prob_dist=[0.9, 0.05, 0.05], change_notification=True,
                     delta_change_notification=True), steps=130),
    "c5_bridge_split": _c(
        "ns_gym/Bridge-v0",
        lambda S, U: {
            "P_left": U.UniformDrift(S.ContinuousScheduler(), rate=0.05),
            "P_right": U.DistributionDecrementUpdate(S.PeriodicScheduler(2), k=0.03),
        },
        wrapper=dict(initial_prob_dist=([0.9, 0.05, 0.05], [1.0, 0.0, 0.0]),
                     change_notification=True, delta_change_notification=True), steps=130),
    # ---- CartPole: every scalar opcode / scheduler ----------------------------------------
    "cartpole_all_params": _c(
        "CartPole-v1",
        lambda S, U: {
            "gravity": U.DeterministicTrend(S.BurstScheduler(2, 3), slope=0.01),
            "masscart": U.PolynomialTrend(S.WindowScheduler([(2, 4), (10, 12)]), coeffs=[1e-3, 2e-4, -1e-5]),
            "masspole": U.GeometricProgression(S.ContinuousScheduler(start=3, end=40), r=1.01),
            "force_mag": U.ExponentialDecay(S.PeriodicScheduler(4), decay_rate=0.001),
            "tau": U.SigmoidTransition(S.ContinuousScheduler(), a=0.02, b=0.03, k=0.5, t0=10),
            "length": U.LinearInterpolation(S.DiscreteScheduler({1, 5, 9, 30}), 0.5, 0.8, T=25),
        },
        wrapper=dict(change_notification=True, delta_change_notification=True), steps=80),
    "cartpole_lists": _c(
        "CartPole-v1",
        lambda S, U: {
            "gravity": U.StepWiseUpdate(S.PeriodicScheduler(2), [9.0, 10.5, 12.0]),
            "length": U.CyclicUpdate(S.ContinuousScheduler(), [0.4, 0.5, 0.6, 0.55]),
            "masscart": U.NoUpdate(S.PeriodicScheduler(3)),
            "force_mag": U.OscillatingUpdate(S.ContinuousScheduler(), delta=0.5),
            "masspole": U.DecrementUpdate(S.ContinuousScheduler(), k=0.004),
        },
        wrapper=dict(change_notification=True, delta_change_notification=True), steps=90),
    "cartpole_stochastic": _c(
        "CartPole-v1",
        lambda S, U: {
            "gravity": U.RandomWalkWithDrift(S.ContinuousScheduler(), alpha=0.01, mu=0.0, sigma=0.2),
            "force_mag": U.RandomWalkWithDriftAndTrend(S.PeriodicScheduler(2), alpha=-0.01, mu=0.1, sigma=0.3, slope=0.002),
            "masspole": U.OrnsteinUhlenbeck(S.ContinuousScheduler(), theta=0.1, mu=0.2, sigma=0.01),
            "length": U.OrnsteinUhlenbeck(S.ContinuousScheduler(), theta=0.05, mu=0.7),
            "masscart": U.BoundedRandomWalk(S.ContinuousScheduler(), mu=0.0, sigma=0.2, lo=0.8, hi=1.2),
        },
        wrapper=dict(change_notification=True, delta_change_notification=True), steps=80),
    "cartpole_stochastic_scheds": _c(
        "CartPole-v1",
        lambda S, U: {
            "gravity": U.IncrementUpdate(S.RandomScheduler(probability=0.3, seed=5), k=0.1),
            "force_mag": U.IncrementUpdate(S.DecayingProbabilityScheduler(0.9, 0.05, seed=6), k=0.2),
            "length": U.IncrementUpdate(S.MemorylessScheduler(p=0.25, seed=7), k=0.01),
            "masscart": U.RandomWalk(S.RandomScheduler(probability=0.5, start=2, end=30, seed=8), mu=0, sigma=0.01),
        },
        wrapper=dict(change_notification=True, delta_change_notification=True), steps=80),
    "cartpole_constraint": _c(
        "CartPole-v1",
        lambda S, U: {
            "masspole": U.DecrementUpdate(S.ContinuousScheduler(), k=0.03),
            "gravity": U.DecrementUpdate(S.PeriodicScheduler(2), k=4.0),
            "length": U.StepWiseUpdate(S.PeriodicScheduler(3), [-1.0, 0.6, 0.0, 0.7]),
        },
        wrapper=dict(change_notification=True, delta_change_notification=True), steps=40),
    "cartpole_silent": _c(
        "CartPole-v1",
        lambda S, U: {"masspole": U.IncrementUpdate(S.ContinuousScheduler(), k=0.1)},
        wrapper=dict(), steps=30),
    "cartpole_persistent": _c(
        "CartPole-v1",
        lambda S, U: {
            "masspole": U.IncrementUpdate(S.ContinuousScheduler(), k=0.001),
            "gravity": U.CyclicUpdate(S.PeriodicScheduler(2), [9.0, 9.8, 10.5]),
            "force_mag": U.RandomWalk(S.ContinuousScheduler(), mu=0, sigma=0.05),
        },
        wrapper=dict(change_notification=True, delta_change_notification=True,
                     persistent_params=True), steps=120),
    "cartpole_custom_sched": _c(
        "CartPole-v1",
        lambda S, U: {
            "gravity": U.IncrementUpdate(S.CustomScheduler(lambda t: t % 7 in (1, 2)), k=0.1),
        },
        wrapper=dict(change_notification=True), steps=60),
    "cartpole_memoryless_lists": _c(
        # MemorylessScheduler driving StepWise / Cyclic updates (schedulers.py:92-116 with
        # single_param.py:202-223, 388-408): next-fire time and list cursor are both per-slot state
        "CartPole-v1",
        lambda S, U: {
            "gravity": U.StepWiseUpdate(S.MemorylessScheduler(p=0.3, seed=3), [9.0, 10.5, 12.0, 8.5, 9.9]),
            "length": U.CyclicUpdate(S.MemorylessScheduler(p=0.5, seed=4), [0.4, 0.5, 0.6, 0.55]),
            "masspole": U.IncrementUpdate(S.MemorylessScheduler(p=0.2, seed=5), k=0.01),
        },
        wrapper=dict(change_notification=True, delta_change_notification=True), steps=90),
    # ---- Acrobot / MountainCar / Pendulum extras ----------------------------------------------
    "acrobot_constraints": _c(
        "Acrobot-v1",
        lambda S, U: {
            "LINK_LENGTH_1": U.DecrementUpdate(S.ContinuousScheduler(), k=0.06),
            "LINK_COM_POS_1": U.IncrementUpdate(S.PeriodicScheduler(3), k=0.05),
            "LINK_LENGTH_2": U.DecrementUpdate(S.ContinuousScheduler(), k=0.2),
            "LINK_COM_POS_2": U.IncrementUpdate(S.ContinuousScheduler(), k=0.07),
            "LINK_MASS_1": U.DecrementUpdate(S.ContinuousScheduler(), k=0.15),
            "dt": U.IncrementUpdate(S.PeriodicScheduler(4), k=0.01),
            "LINK_MOI": U.GeometricProgression(S.ContinuousScheduler(), r=0.99),
        },
        wrapper=dict(change_notification=True, delta_change_notification=True), steps=30),
    "mountaincar_constraint": _c(
        "MountainCar-v0",
        lambda S, U: {
            "gravity": U.DecrementUpdate(S.ContinuousScheduler(), k=0.0004),
            "force": U.RandomWalk(S.ContinuousScheduler(), mu=0, sigma=0.0008),
        },
        wrapper=dict(change_notification=True, delta_change_notification=True), steps=60),
    "mountaincar_continuous": _c(
        "MountainCarContinuous-v0",
        lambda S, U: {"power": U.IncrementUpdate(S.ContinuousScheduler(), k=1e-5)},
        wrapper=dict(change_notification=True, delta_change_notification=True), steps=80),
    "pendulum_all": _c(
        "Pendulum-v1",
        lambda S, U: {
            "m": U.GeometricProgression(S.ContinuousScheduler(), r=0.999),
            "l": U.LinearInterpolation(S.ContinuousScheduler(), 1.0, 1.5, T=100),
            "dt": U.DecrementUpdate(S.PeriodicScheduler(10), k=0.02),
            "g": U.RandomWalk(S.PeriodicScheduler(2), mu=0, sigma=3.0),
        },
        wrapper=dict(change_notification=True, delta_change_notification=True), steps=60),
    # ---- gridworld extras ---------------------------------------------------------------------------
    "frozenlake4_ops": _c(
        "FrozenLake-v1",
        lambda S, U: {"P": U.TargetReversion(S.PeriodicScheduler(2), target=[0.4, 0.3, 0.3], theta=0.2)},
        wrapper=dict(initial_prob_dist=[1, 0, 0], change_notification=True,
                     delta_change_notification=True,
                     modified_rewards={"H": -1, "G": 1, "F": 0, "S": 0}), steps=150),
    "frozenlake8_lerp": _c(
        "FrozenLake-v1",
        lambda S, U: {"P": U.DistributionLinearInterpolation(
            S.ContinuousScheduler(), [1.0, 0.0, 0.0], [0.2, 0.4, 0.4], T=30)},
        wrapper=dict(initial_prob_dist=[1, 0, 0], change_notification=True,
                     delta_change_notification=True), make=dict(MAP8), steps=150),
    "frozenlake8_cyclic_stale": _c(
        # scheduler does not fire at t=0: exposes the stale-table-after-reset behaviour
        "FrozenLake-v1",
        lambda S, U: {"P": U.DistributionCyclicUpdate(
            S.ContinuousScheduler(start=3), [[0.5, 0.25, 0.25], [0.8, 0.1, 0.1], [0.0, 0.5, 0.5]])},
        wrapper=dict(initial_prob_dist=[1, 0, 0], change_notification=True,
                     delta_change_notification=True), make=dict(MAP8), steps=200),
    "frozenlake4_increment": _c(
        "FrozenLake-v1",
        lambda S, U: {"P": U.DistributionIncrementUpdate(S.BurstScheduler(3, 2), k=0.07)},
        wrapper=dict(initial_prob_dist=[0.4, 0.3, 0.3], change_notification=True,
                     delta_change_notification=True), steps=120),
    "frozenlake4_noupdate": _c(
        "FrozenLake-v1",
        lambda S, U: {"P": U.DistributionNoUpdate(S.PeriodicScheduler(2))},
        wrapper=dict(initial_prob_dist=[0.6, 0.2, 0.2], change_notification=True,
                     delta_change_notification=True), steps=60),
    "frozenlake4_memoryless_cyclic": _c(
        "FrozenLake-v1",
        lambda S, U: {"P": U.DistributionCyclicUpdate(
            S.MemorylessScheduler(p=0.4, seed=9), [[0.5, 0.25, 0.25], [0.8, 0.1, 0.1], [0.2, 0.4, 0.4]])},
        wrapper=dict(initial_prob_dist=[1, 0, 0], change_notification=True,
                     delta_change_notification=True), steps=120),
    "frozenlake12_custom_map": _c(
        # a 12x12 desc (144 cells): the wrapper accepts any map (toy_text.py:314-319)
        "FrozenLake-v1",
        lambda S, U: {"P": U.DistributionDecrementUpdate(S.PeriodicScheduler(2), k=0.04)},
        wrapper=dict(initial_prob_dist=[0.8, 0.1, 0.1], change_notification=True, delta_change_notification=True),
        make=dict(desc=["SFFFFFFFFFFF", "FFFHFFFFFHFF", "FFFFFFFFFFFF", "FHFFFFHFFFFF", "FFFFFFFFFFHF", "FFFFHFFFFFFF",
                        "FFFFFFFFFFFF", "FFHFFFFFHFFF", "FFFFFFFFFFFF", "FFFFFHFFFFFF", "FHFFFFFFFHFF", "FFFFFFFFFFFG"],
                  max_episode_steps=60), steps=150),
    "frozenlake5_multi_start": _c(
        # three 'S' cells: every reset samples the start cell (categorical_sample over initial_state_distrib);
        # short episodes so that each env resets several times
        "FrozenLake-v1",
        lambda S, U: {"P": U.DistributionDecrementUpdate(S.ContinuousScheduler(), k=0.03)},
        wrapper=dict(initial_prob_dist=[0.9, 0.05, 0.05], change_notification=True, delta_change_notification=True),
        make=dict(desc=["SFFFS", "FHFHF", "FFFFF", "HFFFH", "SFFFG"], max_episode_steps=9), steps=90),
    "bridge_stepwise": _c(
        "ns_gym/Bridge-v0",
        lambda S, U: {"P": U.DistributionStepWiseUpdate(
            S.PeriodicScheduler(4), [[0.8, 0.1, 0.1], [0.6, 0.2, 0.2], [0.34, 0.33, 0.33]])},
        wrapper=dict(initial_prob_dist=[1, 0, 0], change_notification=True,
                     delta_change_notification=True), steps=130),
    "cliff_drift": _c(
        "CliffWalking-v1",
        lambda S, U: {"P": U.DistributionDecrementUpdate(S.ContinuousScheduler(), k=0.02)},
        wrapper=dict(initial_prob_dist=[1, 0, 0, 0], change_notification=True,
                     delta_change_notification=True), make=dict(max_episode_steps=50), steps=120),
    "cliff_terminal": _c(
        "CliffWalking-v1",
        lambda S, U: {"P": U.UniformDrift(S.PeriodicScheduler(3), rate=0.1)},
        wrapper=dict(initial_prob_dist=[0.7, 0.1, 0.1, 0.1], change_notification=True,
                     delta_change_notification=True, terminal_cliff=True), steps=120),
    # ---- stochastic distribution rules ----------------------------------------
    "frozenlake4_random_categorical": _c(
        "FrozenLake-v1",
        lambda S, U: {"P": U.RandomCategorical(S.PeriodicScheduler(3))},
        wrapper=dict(initial_prob_dist=[0.8, 0.1, 0.1], change_notification=True,
                     delta_change_notification=True), steps=120),
    "cliff_random_categorical": _c(
        "CliffWalking-v1",
        lambda S, U: {"P": U.RandomCategorical(S.BurstScheduler(1, 4))},
        wrapper=dict(initial_prob_dist=[0.7, 0.1, 0.1, 0.1], change_notification=True,
                     delta_change_notification=True), make=dict(max_episode_steps=40), steps=100),
    "bridge_lipschitz_bounded": _c(
        "ns_gym/Bridge-v0",
        lambda S, U: {"P": U.LCBoundedDistrubutionUpdate(S.ContinuousScheduler(), L=0.7)},
        wrapper=dict(initial_prob_dist=[0.4, 0.3, 0.3], change_notification=True,
                     delta_change_notification=True), steps=100),
    # ---- C4: heterogeneous batches (per-env rows) ---------------------------------------------------
    "c4_cartpole_rows": _het(
        "CartPole-v1", _c4_cartpole,
        wrapper=dict(change_notification=True, delta_change_notification=True), steps=90),
    "c4_frozenlake8_rows": _het(
        "FrozenLake-v1", _c4_frozenlake,
        wrapper=dict(initial_prob_dist=[1, 0, 0], change_notification=True, delta_change_notification=True),
        make=dict(MAP8, max_episode_steps=200), steps=150),
    "het_cartpole_lean": _het(
        "CartPole-v1", _het_cartpole_lean,
        wrapper=dict(change_notification=True, delta_change_notification=True), steps=80),
    "het_cartpole_wide": _het(
        "CartPole-v1", _het_cartpole_wide,
        wrapper=dict(change_notification=True, delta_change_notification=True), steps=80),
    "het_bridge_split": _het(
        "ns_gym/Bridge-v0", _het_bridge,
        wrapper=dict(initial_prob_dist=([0.9, 0.05, 0.05], [1.0, 0.0, 0.0]),
                     change_notification=True, delta_change_notification=True), steps=100),
}
