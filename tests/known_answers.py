"""Known-answer vectors transcribed from the reference's own unit tests
(/root/reference/tests/test_update_functions.py, test_schedulers.py, test_bridge.py).

UPDATE_KA: (id, builder(S, U) -> update fn, [(param, t), ...] applied in sequence feeding each
result to the next call when ``chain`` is True, expected [(param, flag), ...], reference line).
SCHED_KA:  (id, builder(S) -> scheduler, [t...], expected fire pattern, reference line).
Time arguments are integers (the device clock is an int32), so the two reference vectors that
use t = pi / 2 are not representable and are left to the oracle-vs-reference test.
"""
import math

NEVER = dict(start=999999, end=999999)          # test_update_functions.py:76-78

UPDATE_KA = [
    ("noupdate", lambda S, U: U.NoUpdate(S.ContinuousScheduler()), [(5.0, 1)], [(5.0, 1)], False, "tuf:88-93"),
    ("increment", lambda S, U: U.IncrementUpdate(S.ContinuousScheduler(), k=2.0), [(10.0, 0)], [(12.0, 1)], False, "tuf:108-113"),
    ("increment_neg", lambda S, U: U.IncrementUpdate(S.ContinuousScheduler(), k=-3.0), [(10.0, 0)], [(7.0, 1)], False, "tuf:115-119"),
    ("increment_never", lambda S, U: U.IncrementUpdate(S.ContinuousScheduler(**NEVER), k=5.0), [(10.0, 0)], [(10.0, 0)], False, "tuf:121-126"),
    ("increment_chain", lambda S, U: U.IncrementUpdate(S.ContinuousScheduler(), k=1.0),
     [(0.0, t) for t in range(5)], [(float(t + 1), 1) for t in range(5)], True, "tuf:128-133"),
    ("decrement", lambda S, U: U.DecrementUpdate(S.ContinuousScheduler(), k=2.0), [(10.0, 0)], [(8.0, 1)], False, "tuf:139-143"),
    ("trend_t1", lambda S, U: U.DeterministicTrend(S.ContinuousScheduler(), slope=2.0), [(10.0, 1)], [(12.0, 1)], False, "tuf:155-159"),
    ("trend_t0", lambda S, U: U.DeterministicTrend(S.ContinuousScheduler(), slope=5.0), [(10.0, 0)], [(10.0, 1)], False, "tuf:161-164"),
    ("trend_neg", lambda S, U: U.DeterministicTrend(S.ContinuousScheduler(), slope=-1.0), [(10.0, 3)], [(7.0, 1)], False, "tuf:166-169"),
    ("rw_drift_sigma0", lambda S, U: U.RandomWalkWithDrift(S.ContinuousScheduler(), alpha=2.0, mu=0, sigma=0, seed=42),
     [(10.0, 0)], [(12.0, 1)], False, "tuf:215-217"),
    ("rw_drift_trend_sigma0", lambda S, U: U.RandomWalkWithDriftAndTrend(S.ContinuousScheduler(), alpha=1.0, mu=0, sigma=0, slope=2.0, seed=42),
     [(10.0, 3)], [(17.0, 1)], False, "tuf:238-243"),
    ("stepwise", lambda S, U: U.StepWiseUpdate(S.ContinuousScheduler(), param_list=[20.0, 30.0, 40.0]),
     [(10.0, 0), (0, 1), (0, 2)], [(20.0, 1), (30.0, 1), (40.0, 1)], True, "tuf:250-258"),
    ("stepwise_empty", lambda S, U: U.StepWiseUpdate(S.ContinuousScheduler(), param_list=[]), [(10.0, 0)], [(10.0, 1)], False, "tuf:261-263"),
    ("stepwise_never", lambda S, U: U.StepWiseUpdate(S.ContinuousScheduler(**NEVER), param_list=[99.0]), [(10.0, 0)], [(10.0, 0)], False, "tuf:266-269"),
    ("oscillating_t0", lambda S, U: U.OscillatingUpdate(S.ContinuousScheduler(), delta=1.0), [(10.0, 0)], [(10.0, 1)], False, "tuf:276-278"),
    ("expdecay_t0", lambda S, U: U.ExponentialDecay(S.ContinuousScheduler(), decay_rate=0.5), [(10.0, 0)], [(10.0, 1)], False, "tuf:297-299"),
    ("expdecay_t1", lambda S, U: U.ExponentialDecay(S.ContinuousScheduler(), decay_rate=1.0), [(10.0, 1)], [(10.0 * math.exp(-1.0), 1)], False, "tuf:302-304"),
    ("geometric", lambda S, U: U.GeometricProgression(S.ContinuousScheduler(), r=2.0), [(5.0, 0)], [(10.0, 1)], False, "tuf:316-319"),
    ("geometric_half", lambda S, U: U.GeometricProgression(S.ContinuousScheduler(), r=0.5), [(10.0, 0)], [(5.0, 1)], False, "tuf:322-324"),
    ("geometric_chain", lambda S, U: U.GeometricProgression(S.ContinuousScheduler(), r=3.0),
     [(1.0, 0), (0, 1), (0, 2)], [(3.0, 1), (9.0, 1), (27.0, 1)], True, "tuf:327-331"),
    ("ou_above", lambda S, U: U.OrnsteinUhlenbeck(S.ContinuousScheduler(), theta=0.5, mu=10.0, sigma=0, seed=42), [(20.0, 0)], [(15.0, 1)], False, "tuf:339-345"),
    ("ou_below", lambda S, U: U.OrnsteinUhlenbeck(S.ContinuousScheduler(), theta=0.5, mu=10.0, sigma=0, seed=42), [(0.0, 0)], [(5.0, 1)], False, "tuf:349-354"),
    ("ou_eq", lambda S, U: U.OrnsteinUhlenbeck(S.ContinuousScheduler(), theta=0.5, mu=10.0, sigma=0, seed=42), [(10.0, 0)], [(10.0, 1)], False, "tuf:358-362"),
    ("sigmoid_mid", lambda S, U: U.SigmoidTransition(S.ContinuousScheduler(), a=0.0, b=10.0, k=1.0, t0=50), [(0.0, 50)], [(5.0, 1)], False, "tuf:415-419"),
    ("cyclic", lambda S, U: U.CyclicUpdate(S.ContinuousScheduler(), value_list=[10.0, 20.0, 30.0]),
     [(0.0, 0), (0, 1), (0, 2)], [(10.0, 1), (20.0, 1), (30.0, 1)], True, "tuf:447-456"),
    ("cyclic_wrap", lambda S, U: U.CyclicUpdate(S.ContinuousScheduler(), value_list=[1.0, 2.0]),
     [(0.0, t) for t in range(6)], [(v, 1) for v in (1.0, 2.0, 1.0, 2.0, 1.0, 2.0)], True, "tuf:460-466"),
    ("brw_hi", lambda S, U: U.BoundedRandomWalk(S.ContinuousScheduler(), mu=100.0, sigma=0, lo=0.0, hi=15.0, seed=42), [(10.0, 0)], [(15.0, 1)], False, "tuf:497-501"),
    ("brw_lo", lambda S, U: U.BoundedRandomWalk(S.ContinuousScheduler(), mu=-100.0, sigma=0, lo=5.0, hi=20.0, seed=42), [(10.0, 0)], [(5.0, 1)], False, "tuf:505-509"),
    ("poly_linear", lambda S, U: U.PolynomialTrend(S.ContinuousScheduler(), coeffs=[2.0]), [(10.0, 3)], [(16.0, 1)], False, "tuf:536-540"),
    ("poly_quadratic", lambda S, U: U.PolynomialTrend(S.ContinuousScheduler(), coeffs=[0, 1.0]), [(10.0, 3)], [(19.0, 1)], False, "tuf:544-547"),
    ("poly_cubic", lambda S, U: U.PolynomialTrend(S.ContinuousScheduler(), coeffs=[0, 0, 0.5]), [(0.0, 2)], [(4.0, 1)], False, "tuf:551-554"),
    ("poly_mixed", lambda S, U: U.PolynomialTrend(S.ContinuousScheduler(), coeffs=[1.0, -0.5]), [(10.0, 4)], [(6.0, 1)], False, "tuf:558-561"),
    ("poly_t0", lambda S, U: U.PolynomialTrend(S.ContinuousScheduler(), coeffs=[5.0, 3.0]), [(10.0, 0)], [(10.0, 1)], False, "tuf:564-567"),
    ("lerp_start", lambda S, U: U.LinearInterpolation(S.ContinuousScheduler(), start_val=0.0, end_val=10.0, T=100), [(0.0, 0)], [(0.0, 1)], False, "tuf:580-585"),
    ("lerp_end", lambda S, U: U.LinearInterpolation(S.ContinuousScheduler(), start_val=0.0, end_val=10.0, T=100), [(0.0, 100)], [(10.0, 1)], False, "tuf:588-592"),
    ("lerp_mid", lambda S, U: U.LinearInterpolation(S.ContinuousScheduler(), start_val=0.0, end_val=10.0, T=100), [(0.0, 50)], [(5.0, 1)], False, "tuf:595-599"),
    ("lerp_clamp", lambda S, U: U.LinearInterpolation(S.ContinuousScheduler(), start_val=0.0, end_val=10.0, T=100), [(0.0, 200)], [(10.0, 1)], False, "tuf:603-607"),
    ("periodic_integration", lambda S, U: U.IncrementUpdate(S.PeriodicScheduler(period=2), k=1.0),
     [(0.0, t) for t in range(6)], [(1.0, 1), (1.0, 0), (2.0, 1), (2.0, 0), (3.0, 1), (3.0, 0)], True, "tuf:1005-1013"),
]

# distributions: expected (list, flag); comparisons use allclose like the reference's tests
DIST_KA = [
    ("d_noupdate", lambda S, U: U.DistributionNoUpdate(S.ContinuousScheduler()), [([0.5, 0.3, 0.2], 0)], [([0.5, 0.3, 0.2], 1)], False, "tuf:633-636"),
    ("d_increment", lambda S, U: U.DistributionIncrementUpdate(S.ContinuousScheduler(), k=0.1), [([0.5, 0.25, 0.25], 0)], [([0.6, 0.2, 0.2], 1)], False, "tuf:673-678"),
    ("d_increment_clamp", lambda S, U: U.DistributionIncrementUpdate(S.ContinuousScheduler(), k=0.9), [([0.5, 0.25, 0.25], 0)], [([1.0, 0.0, 0.0], 1)], False, "tuf:681-685"),
    ("d_decrement", lambda S, U: U.DistributionDecrementUpdate(S.ContinuousScheduler(), k=0.1), [([0.5, 0.25, 0.25], 0)], [([0.4, 0.3, 0.3], 1)], False, "tuf:701-706"),
    ("d_decrement_clamp", lambda S, U: U.DistributionDecrementUpdate(S.ContinuousScheduler(), k=0.9), [([0.5, 0.25, 0.25], 0)], [([0.0, 0.5, 0.5], 1)], False, "tuf:709-713"),
    ("d_stepwise", lambda S, U: U.DistributionStepWiseUpdate(S.ContinuousScheduler(), update_values=[[0.6, 0.2, 0.2], [0.3, 0.4, 0.3]]),
     [([0.5, 0.25, 0.25], 0), (None, 1)], [([0.6, 0.2, 0.2], 1), ([0.3, 0.4, 0.3], 1)], True, "tuf:730-737"),
    ("d_stepwise_empty", lambda S, U: U.DistributionStepWiseUpdate(S.ContinuousScheduler(), update_values=[]), [([0.5, 0.25, 0.25], 0)], [([0.5, 0.25, 0.25], 1)], False, "tuf:740-743"),
    ("d_uniform_full", lambda S, U: U.UniformDrift(S.ContinuousScheduler(), rate=1.0), [([0.7, 0.2, 0.1], 0)], [([1 / 3, 1 / 3, 1 / 3], 1)], False, "tuf:751-756"),
    ("d_uniform_zero", lambda S, U: U.UniformDrift(S.ContinuousScheduler(), rate=0.0), [([0.7, 0.2, 0.1], 0)], [([0.7, 0.2, 0.1], 1)], False, "tuf:760-763"),
    ("d_uniform_half", lambda S, U: U.UniformDrift(S.ContinuousScheduler(), rate=0.5), [([1.0, 0.0, 0.0], 0)], [([2 / 3, 1 / 6, 1 / 6], 1)], False, "tuf:767-772"),
    ("d_uniform_never", lambda S, U: U.UniformDrift(S.ContinuousScheduler(**NEVER), rate=0.5), [([0.8, 0.1, 0.1], 0)], [([0.8, 0.1, 0.1], 0)], False, "tuf:789-793"),
    ("d_target_full", lambda S, U: U.TargetReversion(S.ContinuousScheduler(), target=[0.2, 0.4, 0.4], theta=1.0), [([0.8, 0.1, 0.1], 0)], [([0.2, 0.4, 0.4], 1)], False, "tuf:802-806"),
    ("d_target_zero", lambda S, U: U.TargetReversion(S.ContinuousScheduler(), target=[0.2, 0.4, 0.4], theta=0.0), [([0.8, 0.1, 0.1], 0)], [([0.8, 0.1, 0.1], 1)], False, "tuf:811-814"),
    ("d_target_half", lambda S, U: U.TargetReversion(S.ContinuousScheduler(), target=[0.2, 0.4, 0.4], theta=0.5), [([0.8, 0.1, 0.1], 0)], [([0.5, 0.25, 0.25], 1)], False, "tuf:819-823"),
    ("d_lerp_start", lambda S, U: U.DistributionLinearInterpolation(S.ContinuousScheduler(), [1.0, 0.0, 0.0], [0.0, 0.5, 0.5], T=100), [([0.0, 0.0, 1.0], 0)], [([1.0, 0.0, 0.0], 1)], False, "tuf:854-859"),
    ("d_lerp_end", lambda S, U: U.DistributionLinearInterpolation(S.ContinuousScheduler(), [1.0, 0.0, 0.0], [0.0, 0.5, 0.5], T=100), [([0.0, 0.0, 1.0], 100)], [([0.0, 0.5, 0.5], 1)], False, "tuf:864-868"),
    ("d_cyclic", lambda S, U: U.DistributionCyclicUpdate(S.ContinuousScheduler(), [[0.5, 0.25, 0.25], [0.2, 0.4, 0.4]]),
     [([1.0, 0.0, 0.0], 0), (None, 1), (None, 2)], [([0.5, 0.25, 0.25], 1), ([0.2, 0.4, 0.4], 1), ([0.5, 0.25, 0.25], 1)], True, "tuf:900-955"),
]

SCHED_KA = [
    ("continuous_range", lambda S: S.ContinuousScheduler(start=3, end=7), list(range(10)),
     [False] * 3 + [True] * 5 + [False] * 2, "tsch:59-65"),
    ("continuous_point", lambda S: S.ContinuousScheduler(start=5, end=5), [4, 5, 6], [False, True, False], "tsch:67-72"),
    ("periodic3", lambda S: S.PeriodicScheduler(period=3), list(range(7)), [True, False, False, True, False, False, True], "tsch:146-177"),
    ("periodic_gated", lambda S: S.PeriodicScheduler(period=2, start=3, end=8), list(range(11)),
     [False, False, False, False, True, False, True, False, True, False, False], "tsch:164-170"),
    ("discrete", lambda S: S.DiscreteScheduler({2, 5, 9}), list(range(11)), [t in (2, 5, 9) for t in range(11)], "tsch:89-140"),
    ("burst_2_3", lambda S: S.BurstScheduler(2, 3), list(range(10)), [True, True, False, False, False] * 2, "tsch:394-401"),
    ("window", lambda S: S.WindowScheduler([(2, 4), (8, 9)]), list(range(12)),
     [2 <= t <= 4 or 8 <= t <= 9 for t in range(12)], "tsch:524-589"),
    ("window_gated", lambda S: S.WindowScheduler([(0, 10)], start=3, end=6), list(range(12)), [3 <= t <= 6 for t in range(12)], "tsch:560-589"),
    ("random_p0", lambda S: S.RandomScheduler(probability=0.0, seed=1), list(range(8)), [False] * 8, "tsch:221-229"),
    ("random_p1", lambda S: S.RandomScheduler(probability=1.0, seed=1), list(range(8)), [True] * 8, "tsch:221-229"),
    ("custom_even", lambda S: S.CustomScheduler(lambda t: t % 2 == 0), list(range(6)), [True, False] * 3, "tsch:235-284"),
]
