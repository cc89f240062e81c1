"""Speed of the oracle port relative to the REAL reference (build container only: the reference is
imported from /root/reference on top of the restated gymnasium shim, oracle/ref_loader.py).

    python tools/port_vs_reference.py            # writes profiles/port_vs_reference.json

Both run the same next-step-autoreset vector loop (oracle/vector.py) over 32 envs on ONE core with
the same actions; bench.py copies the entry of its workload into cpu_baseline.port_vs_reference, so
the CPU arm of the bench (the port -- the reference cannot run on the GPU box) can be converted
into what the reference itself would have done."""
import json
import os
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import harness, ref_loader, vector  # noqa: E402
from tests.cases import CASES  # noqa: E402

CASE_NAMES = ["c1_cartpole_readme", "c2_frozenlake8_stepchange", "c3_acrobot", "c3_mountaincar", "c3_pendulum",
              "c4_cartpole_rows", "c5_bridge_uniform", "c5_bridge_split"]
N_ENVS, N_STEPS = 32, 400


def rate(builder, case):
    import numpy as np

    envs = builder(case, N_ENVS)
    vec = vector.SyncVector(envs, None, None)
    actions = harness.draw_actions(case, 1, N_STEPS, N_ENVS)
    vec.reset(k=0)
    for k in range(20):
        vec.step(actions[k], k=k + 1)
    t0 = time.perf_counter()
    for k in range(20, N_STEPS):
        vec.step(actions[k], k=k + 1)
    return N_ENVS * (N_STEPS - 20) / (time.perf_counter() - t0)


def main():
    assert ref_loader.available(), "needs /root/reference"
    out = {"_how": f"{N_ENVS} envs x {N_STEPS - 20} timed vector steps, one core, same actions; "
                   "ratio = port steps/s / reference steps/s (> 1: the port is faster, the bench's CPU arm flatters the CPU)"}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for name in CASE_NAMES:
            case = CASES[name]
            ref = rate(harness.reference_envs, case)
            port = rate(harness.port_envs, case)
            out[name] = {"reference_steps_per_s_1core": ref, "port_steps_per_s_1core": port, "ratio": port / ref}
            print(name, out[name])
    with open(os.path.join(ROOT, "profiles", "port_vs_reference.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
