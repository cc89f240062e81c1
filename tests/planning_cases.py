"""Planning-copy scenarios: run a case for ``k0`` steps, take
``get_planning_env()`` of every env, keep stepping the COPIES for ``k1`` steps.

``wrapper`` overrides the case's wrapper keywords: the agent is *told* about the parameters
(``delta_change_notification`` -> the copy starts from the current theta) or not (the copy
starts from the initial theta), and the copy's parameters either keep evolving
(``in_sim_change``) or are frozen.  ``gpu=False`` marks the one reference quirk the kernel does
not reproduce (CartPole derived masses left stale, classic_control.py:131-135)."""
from tests.cases import CASES

TOLD = dict(change_notification=True, delta_change_notification=True)
UNTOLD = dict(change_notification=False, delta_change_notification=False)


def _p(case, wrapper, k0, k1, gpu=True, params=None):
    c = dict(CASES[case])
    c["wrapper"] = {**c["wrapper"], **wrapper}
    if params is not None:
        c["params"] = params
        c.pop("params_of", None)
    return dict(case=c, k0=k0, k1=k1, gpu=gpu)


PLAN_CASES = {
    "cartpole_told_frozen": _p("c1_cartpole_readme", {**TOLD, "in_sim_change": False}, 12, 25),
    "cartpole_told_evolving": _p("cartpole_lists", {**TOLD, "in_sim_change": True}, 7, 30),
    "cartpole_untold_frozen": _p(
        "c1_cartpole_readme", {**UNTOLD, "in_sim_change": False}, 10, 20,
        params=lambda S, U: {"gravity": U.IncrementUpdate(S.ContinuousScheduler(), k=0.3),
                             "force_mag": U.RandomWalk(S.PeriodicScheduler(2), mu=0.0, sigma=0.5)}),
    "cartpole_untold_stale_masses": _p("c1_cartpole_readme", {**UNTOLD, "in_sim_change": False}, 10, 20, gpu=False),
    "cartpole_stochastic_evolving": _p("cartpole_stochastic_scheds", {**TOLD, "in_sim_change": True}, 9, 30),
    "acrobot_told_frozen": _p("c3_acrobot", {**TOLD, "in_sim_change": False}, 8, 15),
    "pendulum_untold_evolving": _p("c3_pendulum", {**UNTOLD, "in_sim_change": True}, 6, 25),
    "mountaincar_told_frozen": _p("c3_mountaincar", {**TOLD, "in_sim_change": False}, 15, 25),
    "frozenlake8_told_frozen": _p("c2_frozenlake8_drift", {**TOLD, "in_sim_change": False}, 6, 120),
    "frozenlake8_untold_frozen": _p("c2_frozenlake8_drift", {**UNTOLD, "in_sim_change": False}, 6, 40),
    "bridge_told_evolving": _p("c5_bridge_split", {**TOLD, "in_sim_change": True}, 5, 40),
    "bridge_untold_frozen": _p("c5_bridge_uniform", {**UNTOLD, "in_sim_change": False}, 5, 40),
    "cliff_told_frozen": _p("cliff_drift", {**TOLD, "in_sim_change": False}, 9, 60),
    "het_cartpole_told_evolving": _p("c4_cartpole_rows", {**TOLD, "in_sim_change": True}, 10, 30),
}
