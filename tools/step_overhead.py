"""Wall-clock cost of one step() / step_raw() call at several batch sizes (GPU box): where the launch-bound
regime starts and what the host side of a launch costs (lowering of the spec, ctypes, Python packaging).

    python tools/step_overhead.py
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ns_gym_b200 as nsb  # noqa: E402
from ns_gym_b200.schedulers import ContinuousScheduler, PeriodicScheduler  # noqa: E402
from ns_gym_b200.update_functions import IncrementUpdate, RandomWalk  # noqa: E402
from ns_gym_b200.wrappers import NSClassicControlWrapper  # noqa: E402

for n in (1 << 24, 1 << 20, 1 << 16, 1 << 12, 1 << 8):
    env = NSClassicControlWrapper(
        nsb.make("CartPole-v1", num_envs=n),
        {"masspole": IncrementUpdate(ContinuousScheduler(), k=0.1), "gravity": RandomWalk(PeriodicScheduler(period=3))},
        change_notification=True)
    env.reset(seed=0)
    a = env.action_space.sample()
    for name, fn in (("step_raw", lambda: env.step_raw(a)), ("step", lambda: env.step(a))):
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        K = 300
        t0 = time.perf_counter()
        for _ in range(K):
            fn()
        t_issue = (time.perf_counter() - t0) / K          # host time to ISSUE a step (asynchronous)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / K
        print(f"n={n:9d} {name:8s}: {dt * 1e6:7.1f} us/step ({n / dt:.3e} steps/s), host issue {t_issue * 1e6:6.1f} us, "
              f"specialised={env.last_kernel_specialized}")
