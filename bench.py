#!/usr/bin/env python
"""bench.py -- NS env-steps/s of the hot path on B200 (BASELINE.json metric), one JSON line.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port)

A *step* is one pass of the hot path over one batch: one kernel launch that advances every
env of the batch by one NS step (scheduler fire test -> theta advance -> dynamics ->
termination / truncation -> next-step autoreset).

Workload (config.workload): the C1-shaped NS-CartPole throughput variant of SURVEY 8(d) -- the
configuration BASELINE.json's target is quoted on: CartPole-v1, masspole IncrementUpdate(k=0.1)
/ ContinuousScheduler + gravity RandomWalk / PeriodicScheduler(3), change_notification, fp32
fast mode, native Philox draws, next-step autoreset, 2^24 envs per GPU (1.1 GB of SoA state:
far larger than the 126 MB L2, so every step streams from HBM).

value   whole-job env-steps/s with actions and state resident in HBM (CUDA events, max over ranks)
e2e     the same metric through the C-ABI host call nsgym_step_host: actions copied from pinned
        host memory and observation + reward + flags copied back, every step, inside the timed region
roofline  algorithmic bytes per launch / CUDA-event launch duration vs MEASURED_PEAKS.json hbm_gbs
cpu_baseline  the oracle port (the reference's algorithm on the host cores), bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ns_env_steps_per_sec"
UNIT = "env-steps/s"

WORKLOADS = {
    # name: (case builder key, env count per GPU, precision)
    "c1_cartpole": dict(case="c1_cartpole_readme", log2_envs=24, precision="fp32"),
    "c1_cartpole_fp64": dict(case="c1_cartpole_readme", log2_envs=23, precision="fp64"),
    "c2_frozenlake8": dict(case="c2_frozenlake8_stepchange", log2_envs=20, precision="fp64"),
    "c2_frozenlake8_16m": dict(case="c2_frozenlake8_stepchange", log2_envs=24, precision="fp64"),
    "c3_acrobot": dict(case="c3_acrobot", log2_envs=22, precision="fp32"),
    "c3_acrobot_fp64": dict(case="c3_acrobot", log2_envs=22, precision="fp64"),
    "c3_mountaincar": dict(case="c3_mountaincar", log2_envs=22, precision="fp32"),
    "c3_pendulum": dict(case="c3_pendulum", log2_envs=22, precision="fp32"),
    "c3_mountaincar_fp64": dict(case="c3_mountaincar", log2_envs=22, precision="fp64"),
    "c3_pendulum_fp64": dict(case="c3_pendulum", log2_envs=22, precision="fp64"),
    "c5_bridge": dict(case="c5_bridge_uniform", log2_envs=24, precision="fp64"),
    # C4: heterogeneous batch, per-env opcode rows, half CartPole (fp32) + half FrozenLake 8x8
    # (envs bucketed by opcode signature within the shard; coefficients stay per env)
    "c4_hetero": dict(case="c4_cartpole_rows", log2_envs=23, precision="fp32", hetero=True),
    # K-step fused rollouts (state + theta in registers, device-side uniform-random policy)
    "c5_bridge_rollout8": dict(case="c5_bridge_uniform", log2_envs=24, precision="fp64", rollout_k=8),
    "c5_bridge_rollout32": dict(case="c5_bridge_uniform", log2_envs=24, precision="fp64", rollout_k=32),
    "c5_bridge_rollout100": dict(case="c5_bridge_uniform", log2_envs=26, precision="fp64", rollout_k=100),
    "c5_bridge_split_rollout32": dict(case="c5_bridge_split", log2_envs=24, precision="fp64", rollout_k=32),
    "c1_cartpole_rollout32": dict(case="c1_cartpole_readme", log2_envs=24, precision="fp32", rollout_k=32),
    "c3_acrobot_rollout32": dict(case="c3_acrobot", log2_envs=22, precision="fp32", rollout_k=32),
    "c3_acrobot_fp64_rollout32": dict(case="c3_acrobot", log2_envs=22, precision="fp64", rollout_k=32),
}


# ----------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks DURING the timed region")
# ----------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, period_ms: int = 20):
        self.samples = []
        self.proc = None
        self.index = index
        self.period_ms = period_ms
        self.t_mark = [None, None]

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 f"-lms", str(self.period_ms), "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        threading.Thread(target=self._pump, daemon=True).start()

    def _pump(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def mark(self, which):
        self.t_mark[which] = time.perf_counter()

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        rows = []
        for ts, line in self.samples:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                rows.append((ts, float(parts[0]), float(parts[1]), parts[3:7]))
            except ValueError:
                continue
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        t0, t1 = self.t_mark
        inside = [r for r in rows if t0 is not None and t1 is not None and t0 <= r[0] <= t1]
        use = inside if len(inside) >= 3 else rows
        clocks = sorted(r[1] for r in use)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in use for i, v in enumerate(r[3]) if v.lower().startswith("active")})
        return {"sm_mhz": clocks[len(clocks) // 2], "sm_max_mhz": use[0][2], "reasons": reasons,
                "samples": len(use), "samples_in_timed_region": len(inside)}


# ----------------------------------------------------------------------------------------
# CPU side: the reference's algorithm (oracle port) on the host cores
# ----------------------------------------------------------------------------------------
def make_port_envs(case_name, n):
    """Module-level (picklable) factory: n oracle-port envs of a tests/cases.py case."""
    from oracle import harness
    from tests.cases import CASES

    return harness.port_envs(CASES[case_name], n)


def draw_cpu_actions(case_name, rng, n):
    from tests.cases import CASES

    env_id = CASES[case_name]["env_id"]
    if "Pendulum" in env_id:
        return rng.uniform(-2, 2, size=(n, 1))
    if "MountainCarContinuous" in env_id:
        return rng.uniform(-1, 1, size=(n, 1))
    n_act = 2 if "CartPole" in env_id else 3 if ("Acrobot" in env_id or "MountainCar" in env_id) else 4
    return rng.integers(0, n_act, size=n)


def _cpu_make_envs(case_name):
    import functools

    from tests.cases import CASES

    return functools.partial(make_port_envs, case_name), CASES[case_name]


def _cpu_actions(case_name):
    import functools

    return functools.partial(draw_cpu_actions, case_name)


def cpu_baseline(case_name, n_envs_per_proc=64, n_steps=8000, n_procs=None):
    """env-steps/s of the oracle port with one Sync vector loop per host core."""
    from oracle import vector

    make, case = _cpu_make_envs(case_name)
    draw = _cpu_actions(case_name)
    cores = n_procs or os.cpu_count() or 1
    # bounded sample: ~10 s of CPU work per core (the port runs ~8e4 env-steps/s per core)
    sync = vector.run_sync(make, n_envs_per_proc, max(n_steps // 8, 50), draw)
    par = vector.run_parallel(make, n_envs_per_proc, n_steps, draw, cores) if cores > 1 else sync
    return {
        "value": par, "unit": UNIT, "cores": cores, "kind": "port",
        "sync_1core": sync,
        "sample": (f"oracle port of the reference NS wrappers ({case['env_id']}, {case_name}), next-step-autoreset "
                   f"vector loop, {n_envs_per_proc} envs x {n_steps} steps per process, {cores} processes"),
    }


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The reference is
    pure Python over gymnasium; neither gymnasium nor /root/reference exists on the GPU box, so
    the timed code is the oracle port (pinned bit-for-bit against the reference in the build
    container, tests/test_oracle_vs_reference.py), on every host core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    per_step = 64
    make, case = _cpu_make_envs(wl["case"])
    draw = _cpu_actions(wl["case"])
    from oracle import vector

    cores = os.cpu_count() or 1
    steps = max(args.steps, 1)
    # each bench "step" here = one vector step of `per_step` envs per process; bounded so the
    # whole run ends within a few minutes
    n_steps = min(max(steps, 1) * 25, 10000)         # 400 steps -> 10000 vector steps ~ 10 s per core
    warm = vector.run_sync(make, per_step, max(args.warmup, 3), draw)
    del warm
    t0 = time.perf_counter()
    value = vector.run_parallel(make, per_step, n_steps, draw, cores) if cores > 1 else \
        vector.run_sync(make, per_step, n_steps, draw)
    wall = time.perf_counter() - t0
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / n_steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": args.workload, "case": wl["case"], "envs_per_process": per_step,
                   "vector_steps": n_steps, "processes": cores},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{per_step} envs x {n_steps} vector steps per process, {cores} processes"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------
# GPU side
# ----------------------------------------------------------------------------------------
def synth_rows(kind, template, n, seed):
    """Per-env rows of the C4 batch, drawn vectorised with numpy (update function from
    {Increment, Decrement, Trend, Geometric, LinearInterp, RandomWalk} -- gridworld: {Decrement,
    Increment, UniformDrift, TargetReversion, LinearInterp, StepWise} -- x scheduler from
    {Continuous, Periodic(2..7), Burst, Window, Discrete}, per-env coefficients).  Returns
    (rows[n, n_slots], pool_f, pool_i, bitmap): window lists, bitmaps and distribution tables come
    from shared pools of 1024 entries each."""
    import numpy as np

    from ns_gym_b200 import native as nv
    from ns_gym_b200.compile import rows_dtype

    r = np.random.default_rng([4, seed])
    P = template.spec.n_slots
    rows = np.zeros((n, P), dtype=rows_dtype())
    n_pool = 1024
    # shared pools: windows (4 ints), bitmaps (2 words = 64 event bits), distributions
    wa = r.integers(0, 10, n_pool); wb = r.integers(12, 40, n_pool)
    pool_i = np.stack([wa, wa + r.integers(1, 6, n_pool), wb, wb + r.integers(1, 9, n_pool)], 1).astype(np.int32)
    bitmap = (r.integers(0, 1 << 32, (n_pool, 2), dtype=np.uint64) & r.integers(0, 1 << 32, (n_pool, 2), dtype=np.uint64)
              & r.integers(0, 1 << 32, (n_pool, 2), dtype=np.uint64)).astype(np.uint32)
    end_d = r.dirichlet([1.0, 1.0, 1.0], n_pool)
    start_d = np.tile(np.array([1.0, 0.0, 0.0]), (n_pool, 1))
    lerp_pool = np.concatenate([start_d, end_d - start_d], 1)                 # 6 doubles per entry
    step_pool = r.dirichlet([3.0, 1.0, 1.0], (n_pool, 3)).reshape(n_pool, 9)  # up to 3 distributions per entry
    pool_f = np.concatenate([lerp_pool.reshape(-1), step_pool.reshape(-1)])
    step_off = lerp_pool.size
    for j in range(P):
        key = template.spec.slots[j]
        col = rows[:, j]
        col["theta_index"] = key.theta_index
        col["constraint"] = key.constraint
        col["partner_slot"] = key.partner_slot
        col["partner_index"] = key.partner_index
        col["istate_plane"] = -1
        col["start"] = 0
        col["end"] = nv.INT32_MAX
        sk = r.integers(0, 5, n)
        si = np.zeros((n, 4), dtype=np.int32)
        sched = np.zeros(n, dtype=np.int32)
        m = sk == 1; sched[m] = nv.SCHED_PERIODIC; si[m, 0] = r.integers(2, 8, m.sum())
        m = sk == 2; sched[m] = nv.SCHED_BURST; cyc = r.integers(2, 7, m.sum()); si[m, 1] = cyc
        si[m, 0] = 1 + (r.integers(0, 1 << 16, m.sum()) % cyc)
        m = sk == 3; sched[m] = nv.SCHED_WINDOW; si[m, 0] = 4 * r.integers(0, n_pool, m.sum()); si[m, 1] = 2
        m = sk == 4; sched[m] = nv.SCHED_BITMAP; si[m, 0] = 2 * r.integers(0, n_pool, m.sum()); si[m, 1] = 64
        col["sched_op"] = sched
        col["si"] = si
        uk = r.integers(0, 6, n)
        uf = np.zeros((n, 6))
        ui = np.zeros((n, 4), dtype=np.int32)
        upd = np.zeros(n, dtype=np.int32)
        if kind == "grid":
            m = uk == 0; upd[m] = nv.UPD_D_DEC; uf[m, 0] = r.uniform(0.01, 0.08, m.sum())
            m = uk == 1; upd[m] = nv.UPD_D_INC; uf[m, 0] = r.uniform(0.0, 0.05, m.sum())
            m = uk == 2; upd[m] = nv.UPD_D_UNIFORM; rate = r.uniform(0.01, 0.2, m.sum())
            uf[m, 0] = 1 - rate; uf[m, 1] = rate * (1.0 / 3)
            m = uk == 3; upd[m] = nv.UPD_D_TARGET; uf[m, 0] = r.uniform(0.05, 0.4, m.sum())
            uf[m, 1:4] = r.dirichlet([2.0, 1.0, 1.0], m.sum())
            m = uk == 4; upd[m] = nv.UPD_D_LERP; uf[m, 0] = r.integers(10, 80, m.sum()); ui[m, 0] = 6 * r.integers(0, n_pool, m.sum())
            m = uk == 5; upd[m] = nv.UPD_D_STEPWISE; ui[m, 0] = step_off + 9 * r.integers(0, n_pool, m.sum())
            ui[m, 1] = r.integers(1, 4, m.sum())
            ist = col["istate_plane"]; ist[m] = 0; col["istate_plane"] = ist
        else:
            y0 = float(template.spec.theta_init[key.theta_index][0])
            scale = 0.05 * abs(y0) if y0 else 0.01
            m = uk == 0; upd[m] = nv.UPD_ADD; uf[m, 0] = r.uniform(0.1, 1.0, m.sum()) * scale
            m = uk == 1; upd[m] = nv.UPD_ADD; uf[m, 0] = -r.uniform(0.01, 0.2, m.sum()) * scale
            m = uk == 2; upd[m] = nv.UPD_ADD_T; uf[m, 0] = r.uniform(-0.01, 0.02, m.sum()) * scale
            m = uk == 3; upd[m] = nv.UPD_MUL; uf[m, 0] = r.uniform(0.98, 1.03, m.sum())
            m = uk == 4; upd[m] = nv.UPD_LERP; s0 = y0 * r.uniform(0.8, 1.0, m.sum())
            uf[m, 0] = s0; uf[m, 1] = y0 * r.uniform(1.0, 1.5, m.sum()) - s0; uf[m, 2] = r.integers(20, 200, m.sum())
            m = uk == 5; upd[m] = nv.UPD_RW; uf[m, 1] = r.uniform(-0.02, 0.02, m.sum()) * scale
            uf[m, 2] = r.uniform(0.0, 0.3, m.sum()) * scale
        col["upd_op"] = upd
        col["uf"] = uf
        col["ui"] = ui
        rows[:, j] = col
    return rows, pool_f, pool_i.reshape(-1), bitmap.reshape(-1)


def build_hetero(n_envs, rank, seed=0):
    """C4: n_envs/2 NS-CartPole (fp32) + n_envs/2 NS-FrozenLake 8x8, every env with its own row."""
    import ns_gym_b200.schedulers as PS
    import ns_gym_b200.update_functions as PU
    from ns_gym_b200.vector_env import MixedVectorEnv, NSVectorEnv
    from tests.cases import CASES

    half = n_envs // 2
    shards = []
    for k, (name, kind, precision) in enumerate((("c4_cartpole_rows", "classic", "fp32"),
                                                 ("c4_frozenlake8_rows", "grid", "fp64"))):
        case = CASES[name]
        tp = case["params"](PS, PU)
        from ns_gym_b200.compile import compile_program

        # key-set template: which parameters are bound, their constraints and theta_init
        tmpl = compile_program(case["env_id"], tp, half, precision=precision, **case["wrapper"], **case["make"])
        rows, pool_f, pool_i, bitmap = synth_rows(kind, tmpl, half, seed * 1000 + rank * 2 + k)
        env = NSVectorEnv(case["env_id"], tp, half, precision=precision, autoreset="next_step", seed=seed,
                          env_id_offset=rank * n_envs + k * half, rows=rows,
                          pools=(pool_f, pool_i, bitmap), bucket=True, **case["wrapper"], **case["make"])
        shards.append(env)
    return MixedVectorEnv(shards), CASES["c4_cartpole_rows"]


def build_env(workload, n_envs, rank, seed=0):
    import ns_gym_b200.schedulers as PS
    import ns_gym_b200.update_functions as PU
    from ns_gym_b200.vector_env import NSVectorEnv
    from tests.cases import CASES

    wl = WORKLOADS[workload]
    if wl.get("hetero"):
        return build_hetero(n_envs, rank, seed)
    case = CASES[wl["case"]]
    env = NSVectorEnv(case["env_id"], case["params"](PS, PU), n_envs, precision=wl["precision"],
                      autoreset="next_step", seed=seed, env_id_offset=rank * n_envs,
                      **case.get("wrapper", {}), **case.get("make", {}))
    return env, case


def random_actions(env, gen_seed):
    import torch
    from ns_gym_b200.compile import BOX_ACTION, N_ACTIONS

    g = torch.Generator(device=env.device)
    g.manual_seed(gen_seed)
    kind = env.program.env_kind
    if kind in BOX_ACTION:
        lo, hi = BOX_ACTION[kind]
        return (torch.rand(env.num_envs, generator=g, device=env.device, dtype=env.real) * (hi - lo) + lo)
    return torch.randint(0, N_ACTIONS[kind], (env.num_envs,), generator=g, device=env.device, dtype=torch.int32)


def time_steps(shards, actions, steps, warmup, dist=None, rollout_k=0, policy=None):
    """W untimed + K timed steps (one launch per shard per step), CUDA events on the launching
    stream; returns seconds."""
    import torch

    if rollout_k:
        env = shards[0]
        ret = torch.zeros(env.num_envs, dtype=torch.float32, device=env.device)
        length = torch.zeros(env.num_envs, dtype=torch.int32, device=env.device)

        pol = None
        if policy in ("linear", "linear_per_env"):      # device-side linear / tabular policy, random weights
            g = torch.Generator(device=env.device)
            g.manual_seed(4242)
            shape = env.policy_shape(policy == "linear_per_env")
            if len(env.policy_shape()) == 1:
                pol = torch.randint(0, 4, shape, generator=g, device=env.device, dtype=torch.uint8)
            else:
                pol = torch.randn(shape, generator=g, device=env.device, dtype=torch.float32)

        def launch():
            env.rollout(rollout_k, 1.0, ret, length, policy=pol)
    else:
        if len(shards) > 1:           # C4: the shards of a mixed batch are stepped by one host call
            from ns_gym_b200.vector_env import MixedVectorEnv

            mixed = MixedVectorEnv(shards)

            def launch():
                mixed.step_raw(actions)
        else:
            def launch():
                shards[0].step_raw(actions[0])
    return _time_launches(launch, steps, warmup, dist)


def _time_launches(launch, steps, warmup, dist=None, segments=5):
    """W untimed + EXACTLY `steps` timed launches between two CUDA events on the launching stream
    (barrier + synchronize on both sides).  Intermediate events split the run into `segments`
    repetitions (median of 5); returns (total seconds, [seconds per launch of every segment])."""
    import torch

    for _ in range(warmup):
        launch()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()
    segments = max(1, min(segments, steps))
    cuts = [round(k * steps / segments) for k in range(segments + 1)]
    ev = [torch.cuda.Event(enable_timing=True) for _ in cuts]
    ev[0].record()
    nxt = 1
    for k in range(steps):
        launch()
        if k + 1 == cuts[nxt]:
            ev[nxt].record()
            nxt += 1
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    per_launch = [ev[k].elapsed_time(ev[k + 1]) * 1e-3 / (cuts[k + 1] - cuts[k]) for k in range(segments)]
    return ev[0].elapsed_time(ev[-1]) * 1e-3, per_launch


def load_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        with open(peaks_path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic():
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            return json.load(f)
    return {}


# BASELINE.json configs next to the headline workload: timed briefly in every default run so that the
# driver's bench line carries them (roofline.workloads)
TABLE_WORKLOADS = ["c1_cartpole_fp64", "c2_frozenlake8_16m", "c3_acrobot", "c3_acrobot_fp64", "c3_mountaincar",
                   "c3_mountaincar_fp64", "c3_pendulum", "c3_pendulum_fp64", "c4_hetero", "c5_bridge",
                   "c5_bridge_rollout32", "c1_cartpole_rollout32"]


def quick_measure(workload, rank=0, seed=0, launches=30, warmup=5):
    """{steps_per_s, us_per_launch, frac_algorithmic, frac_physical, ...} of one workload on this GPU:
    `launches` timed launches (CUDA events) at the workload's own batch size."""
    import torch

    wl = WORKLOADS[workload]
    n_envs = 1 << wl["log2_envs"]
    env, _case = build_env(workload, n_envs, rank, seed=seed)
    shards = list(getattr(env, "shards", [env]))
    env.reset(seed=seed)
    actions = [random_actions(s, 1234 + rank + 17 * k) for k, s in enumerate(shards)]
    rollout_k = int(wl.get("rollout_k", 0))
    secs, per = time_steps(shards, actions, launches, warmup, None, rollout_k, "random")
    med = sorted(per)[len(per) // 2]
    peak, _ = load_peak()
    traffic = load_traffic().get(workload)
    out = {"steps_per_s": n_envs * max(rollout_k, 1) / med, "us_per_launch": med * 1e6 / len(shards),
           "envs": n_envs, "launches_per_step": len(shards)}
    if not rollout_k:
        out["frac_algorithmic"] = env.bytes_per_step * n_envs / med / 1e9 / peak
        out["frac_physical"] = (traffic / med / 1e9 / peak) if traffic else None
    del env, shards, actions
    torch.cuda.empty_cache()
    return out


def pcie_ceiling(device, dist=None, mbytes=256, reps=6):
    """What this box gives plain pinned cudaMemcpyAsync traffic, all ranks at once: D2H and H2D copies
    of `mbytes` MiB running concurrently on two streams (torch copy_ = one cudaMemcpyAsync each).
    Returns GB/s per GPU (d2h, h2d) -- the ceiling of the e2e number, which moves 22 B out and 4 B in
    per CartPole env-step."""
    import torch

    n = mbytes << 20
    d_out = torch.empty(n, dtype=torch.uint8, device=device)
    d_in = torch.empty(n // 4, dtype=torch.uint8, device=device)
    h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_in = torch.empty(n // 4, dtype=torch.uint8, pin_memory=True)
    s1, s2 = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)
    best = None
    for rep in range(reps):
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            h_out.copy_(d_out, non_blocking=True)
        with torch.cuda.stream(s2):
            d_in.copy_(h_in, non_blocking=True)
        s1.synchronize()
        s2.synchronize()
        dt = time.perf_counter() - t0
        if rep and (best is None or dt < best):
            best = dt
    return n / best / 1e9, (n // 4) / best / 1e9


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank's threads (and with them the first-touch placement of its pinned host buffers) to
    the CPU set NVML reports as local to its GPU.  Returns a short description for the bench line."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (word >> b) & 1}
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            return f"bound to {len(use)} CPUs local to GPU {local_rank} (NVML affinity)"
        return f"NVML affinity of GPU {local_rank} = all {len(allowed)} allowed CPUs (single NUMA domain visible)"
    except Exception as e:          # no NVML / not permitted: run unbound
        return f"unbound ({type(e).__name__})"


def run_gpu(args):
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa_note = bind_to_gpu_numa_node(local)      # before any pinned allocation: first touch decides the node
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_mod
    from ns_gym_b200 import native

    native.load()   # no CUDA library -> loud failure, never a fallback
    wl = WORKLOADS[args.workload]
    n_envs = 1 << (args.log2_envs or wl["log2_envs"])
    env, case = build_env(args.workload, n_envs, rank, seed=args.seed)
    shards = list(getattr(env, "shards", [env]))      # C4: one shard (= one kernel launch) per env kind
    dev = shards[0].device
    if args.general_kernels:        # the general (all rule classes, injection-capable) instantiations
        for s in shards:
            s.set_option("general_kernels", 1)
    if args.no_specialize:          # the precompiled lean kernels instead of the program-specialised ones
        for s in shards:
            s.set_option("specialize", 0)
    env.reset(seed=args.seed)
    actions = [random_actions(s, 1234 + rank + 17 * k) for k, s in enumerate(shards)]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- device-resident throughput ----
    launches0 = env.launch_count
    sampler.mark(0)
    rollout_k = int(wl.get("rollout_k", 0))
    per_launch = max(rollout_k, 1)                     # env-steps each env advances per launch
    secs, seg = time_steps(shards, actions, args.steps, max(args.warmup, 3), dist, rollout_k, args.rollout_policy)
    launches = env.launch_count - launches0 - max(args.warmup, 3) * len(shards)
    t = torch.tensor([secs], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    secs_max = float(t.item())
    total_steps = world * n_envs * args.steps * per_launch
    value = total_steps / secs_max
    # ---- roofline for the (only) kernel of the step ----
    peak, peak_src = load_peak()
    launch_s = secs / args.steps                       # this rank's average step duration
    seg_sorted = sorted(seg)
    launch_med = seg_sorted[len(seg_sorted) // 2]      # median of the 5 repetitions
    bytes_per_launch_env = env.bytes_per_step
    if rollout_k:
        # K fused steps move state / theta / t once per launch and write the return + length
        # accumulators: (2 S w + 2 P w + 8) + 12 per env per launch
        bytes_per_launch_env = env.bytes_per_step - (shards[0].buffers["action"].element_size() + 4 + 1 + 1) + 12 + 6
    achieved = bytes_per_launch_env * n_envs / launch_s / 1e9
    traffic = load_traffic().get(args.workload)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_unit": "DRAM bytes per step (ncu --set full, profiles/traffic.json)",
                "peak_source": peak_src,
                "algorithmic_bytes_per_env_step": bytes_per_launch_env / per_launch,
                "kernel_us_per_launch": launch_s * 1e6 / len(shards),
                "kernel_us_per_launch_median_of_5": launch_med * 1e6 / len(shards),
                "kernel_us_per_launch_repetitions": [x * 1e6 / len(shards) for x in seg]}
    if traffic and not rollout_k:
        # the measured DRAM rate next to the algorithmic one
        roofline["dram_gbs_from_traffic"] = traffic / launch_s / 1e9
        roofline["frac_physical"] = traffic / launch_s / 1e9 / peak
        if traffic < 0.9 * bytes_per_launch_env * n_envs:
            # gridworld kernels write a distribution back only when its update fired, so with a rare
            # scheduler (C2) the traffic is well below SURVEY 8(d)'s figure (which counts theta read + write
            # every step) and the algorithmic rate can exceed the copy peak: frac is the PHYSICAL rate there
            roofline["frac_algorithmic"] = roofline["frac"]
            roofline["achieved_algorithmic"] = achieved
            roofline["achieved"] = roofline["dram_gbs_from_traffic"]
            roofline["frac"] = roofline["frac_physical"]
            roofline["note"] = ("DRAM traffic is below the algorithmic bytes (unchanged theta planes are not written "
                                "back): achieved / frac are the physical DRAM rate (ncu traffic / CUDA-event time); "
                                "*_algorithmic keep SURVEY 8(d)'s formula")
    if roofline["frac"] > 1.0 and "note" not in roofline:
        roofline["note"] = ("frac > 1: `peak` is the rate of a COPY kernel on this pool (MEASURED_PEAKS.json), not the DRAM "
                            "ceiling -- ncu puts this kernel at 78.6 % of the theoretical 8.18 TB/s "
                            "(profiles/r2_full_c1_cartpole_spec.txt); physical DRAM traffic is "
                            f"{(traffic or 0) / max(n_envs, 1):.1f} B per env-step against {bytes_per_launch_env:.0f} B "
                            "algorithmic (frac_physical)")
    if len(shards) > 1:
        roofline["note"] = ("heterogeneous batch: one launch per env kind per step; bytes include the per-env row "
                            "words read each step: " +
                            ", ".join(f"{s.program.env_id} {s.bytes_per_step:.0f} B/env-step (rows {s.row_bytes_per_env:.0f} B)"
                                      for s in shards))
    if rollout_k:
        roofline["note"] = (f"fused {rollout_k}-step rollout: bytes are amortised over K steps, the kernel is "
                            "FP32/FP64-pipe + issue bound, not HBM bound (see profiles/)")
    # ---- end to end through the C-ABI host call ----
    host_io = [s.make_host_io() for s in shards]
    for (h_act, _), a in zip(host_io, actions):
        h_act.copy_(a.cpu())

    def e2e_step():
        for s, (h_act, h_out) in zip(shards, host_io):
            s.step_host(h_act, h_out, n_chunks=args.chunks)

    e2e_steps = max(min(args.steps, args.e2e_steps), 1)
    for _ in range(3):
        e2e_step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_secs = time.perf_counter() - t0
    sampler.mark(1)            # clocks are sampled over both timed regions (device-resident and end-to-end)
    t = torch.tensor([e2e_secs], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n_envs * e2e_steps / float(t.item())
    h2d = d2h = 0
    for s, (h_act, h_out) in zip(shards, host_io):
        a_, b_ = s.host_bytes_per_step(h_act, h_out)
        h2d, d2h = h2d + a_, d2h + b_
    # what the box gives plain pinned copies, all ranks at once: the ceiling of the e2e number
    ceil_d2h, ceil_h2d = pcie_ceiling(dev, dist)
    cvec = torch.tensor([ceil_d2h, ceil_h2d], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(cvec, op=dist.ReduceOp.SUM)
    # ---- metric reduction over NCCL (the only collective on this path) ----
    specialised = all(getattr(s, "last_kernel_specialized", False) for s in shards)
    stats = torch.stack([sum(s.buffers["reward"].double().sum() for s in shards),
                         sum(((s.buffers["flags"] & 3) != 0).double().sum() for s in shards),
                         torch.tensor(float(n_envs), dtype=torch.float64, device=dev)])
    if dist is not None:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    sampler.stop()
    # ---- the other BASELINE configs, briefly (rank 0, single GPU): roofline.workloads ----
    table = None
    if world == 1 and not args.no_table and args.workload == "c1_cartpole":
        del host_io, actions, shards, env
        torch.cuda.empty_cache()
        table = {args.workload: {"steps_per_s": value, "us_per_launch": launch_med * 1e6,
                                 "frac_algorithmic": achieved / peak,
                                 "frac_physical": (traffic / launch_med / 1e9 / peak) if traffic else None,
                                 "envs": n_envs, "launches_per_step": 1}}
        for name in TABLE_WORKLOADS:
            table[name] = quick_measure(name, rank, args.seed)
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline(wl["case"])
            cpu["port_vs_reference"] = port_vs_reference(wl["case"])
        if table is not None:
            roofline["workloads"] = table
        e2e_gbs = e2e_value * (h2d + d2h) / n_envs / 1e9      # aggregate: h2d / d2h are this rank's bytes per step
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * secs_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if wl["precision"] == "fp32" else "f64", "data": "synthetic",
            "config": {
                "workload": args.workload, "case": wl["case"],
                "env_id": case["env_id"] if not wl.get("hetero") else "CartPole-v1+FrozenLake-v1",
                "envs_per_gpu": n_envs, "global_envs": world * n_envs, "precision": wl["precision"],
                "autoreset": "next_step", "rng": "philox4x32-10 (native)", "rollout_k": rollout_k, "rollout_policy": args.rollout_policy if rollout_k else None,
                "kernels": "general" if args.general_kernels else (
                    "program-specialised (NVRTC at first launch, nsgym_jit.cu)"
                    if specialised
                    else "precompiled lean where the program allows"),
                "parallelism": f"env-shard x{world}, no data-path collective",
                "l2_policy": f"working set {bytes_per_launch_env * n_envs / 1e6:.0f} MB per GPU >> 126 MB L2 "
                             "(inputs larger than L2, no flush needed)",
                "launches_per_step": 1 if not wl.get("hetero") else 2,
                "timing": "CUDA events around exactly `steps` launches; 5 repetitions by intermediate events",
            },
            "roofline": roofline,
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "chunks": args.chunks,
                    "api": "nsgym_step_host (C ABI, pinned host buffers, H2D actions + D2H obs/reward/flags/change)",
                    "gbs": e2e_gbs,
                    "pcie_ceiling_gbs": {"d2h": float(cvec[0]), "h2d": float(cvec[1]),
                                         "how": "plain pinned cudaMemcpyAsync, D2H + H2D (4:1 bytes) concurrently on "
                                                "two streams, all ranks at once, summed over ranks"},
                    "frac_of_pcie_ceiling": e2e_gbs / float(cvec[0] + cvec[1]),
                    "host_binding": numa_note},
            "gpu_launches": launches,
            "clocks": sampler.summary(),
            "batch_stats": {"mean_reward_last_step": float(stats[0] / stats[2]),
                            "ended_fraction_last_step": float(stats[1] / stats[2])},
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def port_vs_reference(case_name):
    """Speed of the oracle port relative to the REAL reference on the same case, measured in the build
    container (tools/port_vs_reference.py; the reference cannot run on the GPU box)."""
    path = os.path.join(ROOT, "profiles", "port_vs_reference.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f).get(case_name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="c1_cartpole", choices=sorted(WORKLOADS))
    ap.add_argument("--log2-envs", type=int, default=None, dest="log2_envs")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=30, dest="e2e_steps")
    ap.add_argument("--chunks", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-table", action="store_true", dest="no_table",
                    help="skip the brief timing of the other BASELINE configs (roofline.workloads)")
    ap.add_argument("--rollout-policy", default="random", choices=["random", "linear", "linear_per_env"],
                    dest="rollout_policy", help="device-side policy of the *_rollout* workloads")
    ap.add_argument("--no-specialize", action="store_true", dest="no_specialize",
                    help="keep the precompiled lean kernels (no run-time specialisation of the step kernel)")
    ap.add_argument("--general-kernels", action="store_true", dest="general_kernels",
                    help="launch the general kernel instantiations instead of the lean ones (kernel experiments)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
