"""Throughput of arbitrary tests/cases.py programs with and without run-time specialisation (GPU box).

    python tools/measure_cases.py [--log2-envs 22] [--precision fp32] case ...
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    from tests import parity_util as pu
    from tests.cases import CASES

    ap = argparse.ArgumentParser()
    ap.add_argument("cases", nargs="+")
    ap.add_argument("--log2-envs", type=int, default=22)
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--steps", type=int, default=100)
    a = ap.parse_args()
    n = 1 << a.log2_envs
    for name in a.cases:
        row = [name]
        for specialize in (0, 1):
            env = pu.gpu_env(CASES[name], n, precision=a.precision, want_delta=False, want_obs=False)
            env.set_option("specialize", specialize)
            env.reset(seed=0)
            if env.action_space_n is None:
                act = torch.rand(n, device=env.device, dtype=env.real) * 2 - 1
            else:
                act = torch.randint(0, env.action_space_n, (n,), device=env.device, dtype=torch.int32)
            for _ in range(10):
                env.step_raw(act)
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(a.steps):
                env.step_raw(act)
            t1.record()
            torch.cuda.synchronize()
            us = t0.elapsed_time(t1) * 1e3 / a.steps
            row.append(f"{'specialised' if env.last_kernel_specialized else 'precompiled'} {us:7.1f} us "
                       f"{n / us * 1e6:.3e} steps/s ({env.bytes_per_step:.0f} B: {env.bytes_per_step * n / us / 1e3 / 6552:.2f})")
            del env
            torch.cuda.empty_cache()
        print(" | ".join(row), flush=True)


if __name__ == "__main__":
    main()
