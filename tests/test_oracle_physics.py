"""CPU: physics-consistency properties of the restated base-env dynamics (oracle/gym_restated.py).

gymnasium cannot be imported, so Acrobot / Pendulum have no numeric pin.  What
can be checked independently is that the restated equations of motion are the Lagrangian dynamics
they claim to be: with zero torque the total energy of the Acrobot ("book" variant) and of the
Pendulum is conserved in the limit of small time steps, for random link parameters -- a sign or
coefficient slip in `_dsdt` / the Pendulum update breaks it at first order."""
import numpy as np
import pytest

from oracle.gym_restated import AcrobotEnv, PendulumEnv


def _acrobot_energy(env, s):
    th1, th2, w1, w2 = s
    m1, m2, l1 = env.LINK_MASS_1, env.LINK_MASS_2, env.LINK_LENGTH_1
    lc1, lc2, I1, I2, g = env.LINK_COM_POS_1, env.LINK_COM_POS_2, env.LINK_MOI, env.LINK_MOI, 9.8
    kin = (0.5 * (I1 + m1 * lc1 ** 2) * w1 ** 2 + 0.5 * I2 * (w1 + w2) ** 2
           + 0.5 * m2 * (l1 ** 2 * w1 ** 2 + lc2 ** 2 * (w1 + w2) ** 2 + 2 * l1 * lc2 * w1 * (w1 + w2) * np.cos(th2)))
    pot = -m1 * g * lc1 * np.cos(th1) - m2 * g * (l1 * np.cos(th1) + lc2 * np.cos(th1 + th2))
    return kin + pot


@pytest.mark.parametrize("seed", range(4))
def test_acrobot_book_dynamics_conserve_energy_without_torque(seed):
    r = np.random.default_rng([11, seed])
    env = AcrobotEnv()
    env.LINK_MASS_1, env.LINK_MASS_2 = r.uniform(0.5, 2.0, 2)
    env.LINK_LENGTH_1 = r.uniform(0.7, 1.5)
    env.LINK_COM_POS_1, env.LINK_COM_POS_2 = r.uniform(0.3, 0.6, 2)
    env.LINK_MOI = r.uniform(0.5, 1.5)
    env.dt = 1e-3                                       # RK4: energy error O(dt^4) per unit time
    env.reset(seed=seed)
    env.state = np.array([r.uniform(-1, 1), r.uniform(-1, 1), r.uniform(-0.5, 0.5), r.uniform(-0.5, 0.5)])
    e0 = _acrobot_energy(env, env.state)
    for _ in range(1500):
        env.step(1)                                     # AVAIL_TORQUE[1] = 0
        assert abs(env.state[2]) < 4 * np.pi and abs(env.state[3]) < 9 * np.pi   # velocity clips not hit
    assert abs(_acrobot_energy(env, env.state) - e0) < 1e-9 * max(1.0, abs(e0))


@pytest.mark.parametrize("seed", range(4))
def test_pendulum_dynamics_conserve_energy_without_torque(seed):
    r = np.random.default_rng([12, seed])
    env = PendulumEnv(g=float(r.uniform(5, 12)))
    env.m, env.l = float(r.uniform(0.5, 2)), float(r.uniform(0.5, 2))
    env.reset(seed=seed)
    th, w = float(r.uniform(2.0, 3.0)), float(r.uniform(-0.5, 0.5))          # swinging, |thdot| stays < 8

    def energy(th, w):                                  # rod about its end: I = m l^2 / 3; theta = 0 is upright
        return 0.5 * (env.m * env.l ** 2 / 3) * w ** 2 + env.m * env.g * (env.l / 2) * np.cos(th)

    drift = []
    for dt in (2e-3, 1e-3):                             # semi-implicit Euler: drift is first order in dt
        env.dt = dt
        env.state = np.array([th, w])
        e0 = energy(th, w)
        worst = 0.0
        for _ in range(int(round(1.0 / dt))):
            env.step(np.array([0.0]))
            worst = max(worst, abs(energy(*env.state) - e0))
        drift.append(worst / abs(e0))
    assert drift[0] < 2e-2 and drift[1] < 0.6 * drift[0], drift
