import time, torch, sys
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import ns_gym_b200 as nsb
from ns_gym_b200.schedulers import ContinuousScheduler, PeriodicScheduler
from ns_gym_b200.update_functions import IncrementUpdate, RandomWalk
from ns_gym_b200.wrappers import NSClassicControlWrapper
for n in (1 << 24, 1 << 16, 1 << 8):
    env = NSClassicControlWrapper(nsb.make("CartPole-v1", num_envs=n),
        {"masspole": IncrementUpdate(ContinuousScheduler(), k=0.1), "gravity": RandomWalk(PeriodicScheduler(period=3))},
        change_notification=True)
    env.reset(seed=0)
    a = env.action_space.sample()
    for name, fn in (("step_raw", lambda: env.step_raw(a)), ("step", lambda: env.step(a))):
        for _ in range(5): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        K = 100
        for _ in range(K): fn()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / K
        print(f"n={n} {name}: {dt*1e6:.1f} us/step  {n/dt:.3e} steps/s")
