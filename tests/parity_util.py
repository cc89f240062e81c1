"""Helpers shared by the parity tests: run a case on the oracle port / the CUDA path and
compare traces.  A trace is the dict produced by ``oracle.vector.trace`` (arrays [K, N, ...])."""
from __future__ import annotations

import numpy as np

from oracle import harness, vector

INT_KEYS = ("terminated", "truncated", "was_reset", "relative_time", "env_change", "gt_change")
# stated tolerances (north_star): deterministic paths 1e-9 relative in fp64 mode
FP64_RTOL, FP64_ATOL = 1e-9, 1e-12


def n_slots_of(case):
    import ns_gym_b200.schedulers as PS
    import ns_gym_b200.update_functions as PU
    return len(case["params"](PS, PU))


def oracle_trace(case, n_envs, seed, steps=None, actions=None):
    K = steps or case["steps"]
    if actions is None:
        actions = harness.draw_actions(case, seed + 1, K, n_envs)
    clock, per_env, u, z = harness.make_streams(seed, n_envs, K + 1, n_slots_of(case))
    envs = harness.port_envs(case, n_envs, per_env)
    tr = vector.trace(vector.SyncVector(envs, per_env, clock), actions)
    return tr, actions, u, z


def oracle_trace_streams(case, clock, per_env, actions):
    """Trace of the oracle port driven by prepared per-env stream objects (tests/philox_np.py:
    the kernels' native draws instead of numpy-drawn tables)."""
    envs = harness.port_envs(case, len(per_env), per_env)
    return vector.trace(vector.SyncVector(envs, per_env, clock), actions)


def gpu_env(case, n_envs, precision="fp64", autoreset="next_step", **extra):
    """The case's batch on the CUDA path (heterogeneous cases: one tunable_params dict per env)."""
    import ns_gym_b200.schedulers as PS
    import ns_gym_b200.update_functions as PU
    from ns_gym_b200.vector_env import NSVectorEnv

    kw = dict(precision=precision, autoreset=autoreset, want_delta=True, want_obs=True,
              **case.get("wrapper", {}), **case.get("make", {}))
    kw.update(extra)
    if "params_of" in case:       # heterogeneous batch (nsgym_create_rows)
        return NSVectorEnv.heterogeneous(case["env_id"], [case["params_of"](PS, PU, e) for e in range(n_envs)], **kw)
    return NSVectorEnv(case["env_id"], case["params"](PS, PU), n_envs, **kw)


def gpu_run(env, actions, u, z, first_row=1, do_reset=True):
    """Drive ``env`` with the injected tables (row 0: reset draws, row first_row + k: step k) and
    record the same trace as ``oracle.vector.trace``."""
    import torch

    from ns_gym_b200 import native as nv

    n_envs = env.num_envs
    dev = env.device
    K = len(actions)
    if u is None:       # native Philox draws: nothing injected (the lean kernels where the program allows)
        U = Z = [None] * (first_row + K)
    else:
        U = torch.as_tensor(u, dtype=torch.float64, device=dev).contiguous()
        Z = torch.as_tensor(z, dtype=torch.float64, device=dev).contiguous()
    keys = env.keys

    def raw_state():
        s = env.buffers["state"]
        return s.double().cpu().numpy() if not env.program.is_grid else s.cpu().numpy().astype(np.int64)

    def obs_state():
        o = env.observation()
        return o.cpu().numpy() if not env.program.is_grid else o.cpu().numpy().astype(np.int64)

    rec = {}
    if do_reset:
        env.reset(inject_uniform=U[0])
        rec = {"obs0": obs_state(), "raw0": raw_state()}
    lists = {k: [] for k in ("obs", "raw", "theta", "reward", "terminated", "truncated", "was_reset",
                             "relative_time", "env_change", "delta_change", "gt_change", "gt_delta")}
    bad = False
    for k in range(K):
        a = torch.as_tensor(np.asarray(actions[k]).reshape(n_envs))
        obs, reward, term, trunc, info = env.step(a, inject_uniform=U[first_row + k], inject_normal=Z[first_row + k])
        bad |= bool((env.buffers["flags"] & nv.FLAG_BAD_DIST).any().item())
        lists["obs"].append(obs_state())
        lists["raw"].append(raw_state())
        th = env.theta()
        cols = []
        for key in keys:
            v = th[key].double().cpu().numpy()
            cols.append(v.T if v.ndim == 2 else v[:, None])
        lists["theta"].append(np.concatenate(cols, axis=1) if cols else np.zeros((n_envs, 0)))
        lists["reward"].append(reward.double().cpu().numpy())
        lists["terminated"].append(term.cpu().numpy())
        lists["truncated"].append(trunc.cpu().numpy())
        lists["was_reset"].append(info["was_reset"].cpu().numpy())
        lists["relative_time"].append(obs["relative_time"].cpu().numpy().astype(np.int64))
        lists["env_change"].append(np.stack([obs["env_change"][q].cpu().numpy().astype(np.int64) for q in keys], 1))
        lists["delta_change"].append(np.stack([obs["delta_change"][q].double().cpu().numpy() for q in keys], 1))
        lists["gt_change"].append(np.stack([info["Ground Truth Env Change"][q].cpu().numpy().astype(np.int64) for q in keys], 1))
        lists["gt_delta"].append(np.stack([info["Ground Truth Delta Change"][q].double().cpu().numpy() for q in keys], 1))
    for k2, v in lists.items():
        rec[k2] = np.stack(v)
    rec["_bad_dist"] = bad
    rec["_kernel_class"] = int(env.lib.nsgym_last_kernel_class(env._h))
    rec["_specialized"] = bool(env.lib.nsgym_last_kernel_specialized(env._h))
    return rec


def gpu_trace(case, n_envs, actions, u, z, precision="fp64", autoreset="next_step"):
    """Same trace from the CUDA path with the same injected tables."""
    return gpu_run(gpu_env(case, n_envs, precision, autoreset), actions, u, z)


# ---- time-indexed transition tables ----
TABLE_CASES = ["c2_frozenlake8_drift", "c2_frozenlake8_stepchange", "frozenlake4_ops", "frozenlake8_cyclic_stale",
               "cliff_terminal", "cliff_drift", "c5_bridge_uniform", "c5_bridge_split", "bridge_stepwise"]


def table_of(env):
    """Current transition table of one env (reference wrapper or port) as arrays [S, A, D]."""
    if hasattr(env, "transition_table"):
        return env.transition_table()
    base = env.unwrapped
    P = base.transition_matrix if type(base).__name__ == "Bridge" else base.P
    S, D = len(P), max(len(P[s][a]) for s in P for a in P[s])
    out = {"prob": np.zeros((S, 4, D)), "next": np.zeros((S, 4, D), dtype=np.int64),
           "reward": np.zeros((S, 4, D)), "done": np.zeros((S, 4, D), dtype=bool)}
    for s in range(S):
        for a in range(4):
            for k, (p, ns, r, d) in enumerate(P[s][a]):
                out["prob"][s, a, k], out["next"][s, a, k], out["reward"][s, a, k], out["done"][s, a, k] = p, ns, r, d
    return out


def oracle_table_trace(builder, case, T, seed=3, env=0):
    """Tables in force at NS times 0..T-1 of ONE env (index ``env`` of the batch: heterogeneous cases
    give every env its own rules) stepped from a reset (stepping past episode ends: the parameters
    keep evolving with t): prob [T, S, A, D] + the time-invariant part."""
    n = env + 1
    actions = harness.draw_actions(case, seed + 1, T, n)
    clock, per_env, _, _ = harness.make_streams(seed, n, T + 1, n_slots_of(case))
    envs = builder(case, n, per_env)
    vec = vector.SyncVector(envs, per_env, clock, autoreset=False)
    vec.reset(k=0)
    probs, last = [], None
    for t in range(T):
        vec.step(actions[t], k=t + 1)
        last = table_of(envs[env])
        probs.append(last["prob"])
    return {"prob": np.stack(probs), "next": last["next"], "reward": last["reward"], "done": last["done"]}


# ---- planning copies (tests/planning_cases.py) ----
def oracle_planning_trace(builder, sc, n_envs, seed, tables=None):
    """(trace of the roots over k0 steps, trace of their planning copies over k1 steps, inputs)."""
    case, k0, k1 = sc["case"], sc["k0"], sc["k1"]
    if tables is None:
        actions = harness.draw_actions(case, seed + 1, k0 + k1, n_envs)
        u, z = harness.S_.draw_tables(seed, n_envs, k0 + k1 + 1, n_slots_of(case))
    else:
        actions, u, z = tables
    clock = harness.S_.Clock()
    per_env = [harness.S_.EnvStreams(u[:, :, i], z[:, :, i], clock) for i in range(n_envs)]
    envs = builder(case, n_envs, per_env)
    tr0 = vector.trace(vector.SyncVector(envs, per_env, clock), actions[:k0])
    plans = harness.planning_envs(envs, per_env)
    tr1 = vector.trace(vector.SyncVector(plans, per_env, clock, autoreset=False), actions[k0:],
                       first_row=k0 + 1, do_reset=False)
    return tr0, tr1, (actions, u, z)


def gpu_planning_trace(sc, n_envs, actions, u, z, precision="fp64"):
    case, k0 = sc["case"], sc["k0"]
    env = gpu_env(case, n_envs, precision)
    tr0 = gpu_run(env, actions[:k0], u, z)
    plan = env.get_planning_env()
    tr1 = gpu_run(plan, actions[k0:], u, z, first_row=k0 + 1, do_reset=False)
    return tr0, tr1


def compare(ref, got, rtol=FP64_RTOL, atol=FP64_ATOL, float_obs_rtol=None, name=""):
    """Integer / flag arrays bit-exact; float arrays within (rtol, atol)."""
    for key in ref:
        if key.startswith("_"):
            continue
        assert key in got, f"{name}: the trace under test lacks '{key}'"
        a, b = np.asarray(ref[key]), np.asarray(got[key])
        assert a.shape == b.shape, f"{name}: {key} shape {a.shape} vs {b.shape}"
        if key in INT_KEYS or a.dtype.kind in "iub":
            assert np.array_equal(a.astype(np.int64), b.astype(np.int64)), (
                f"{name}: {key} differs at {np.argwhere(a.astype(np.int64) != b.astype(np.int64))[:5].tolist()}")
        else:
            # observations and rewards leave the kernel as float32 (gymnasium's obs dtype)
            r = rtol if not (float_obs_rtol and key in ("obs", "obs0", "reward")) else float_obs_rtol
            ok = np.isclose(a, b, rtol=r, atol=atol, equal_nan=True)
            assert ok.all(), (
                f"{name}: {key} off at {np.argwhere(~ok)[:5].tolist()} "
                f"ref={a[~ok][:3]} got={b[~ok][:3]} max_abs={np.nanmax(np.abs(a - b))}")
