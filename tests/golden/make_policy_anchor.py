"""Behavioural anchors for the gymnasium dynamics that have no reference-held vector: the reference's
OWN pre-trained agents (run in the build container only).

    python tests/golden/make_policy_anchor.py

gymnasium is absent from /root/reference (third-party, uv.lock:958-959), so the restated Acrobot /
MountainCar / Pendulum dynamics (oracle/gym_restated.py) cannot be diffed against it, and no reference
test holds a post-step state.  The reference does ship agents its authors trained on the real
gymnasium environments:

    ns_gym/evaluate/evaluation_model_weights/AcrobotEnv/{PPO/ppo_acrobot_default.zip, A2C/a2c_acrobot.zip}
    ns_gym/evaluate/evaluation_model_weights/PendulumEnv/{PPO/ppo_pendulum.zip, DDPG/PendulumEnv.zip}
(Pendulum's A2C agent and the DDQN nets under benchmark_algorithms/DDQN/DDQN_models do not solve their tasks
and are left out; MountainCar has no usable in-tree agent.)

A policy fitted to the true dynamics only keeps solving the task on a restatement that has the same
dynamics: a wrong sign, gain, clip or reward in the restatement shows up as a collapsed return.  This
script converts the policy networks (weights only: greedy / deterministic action) to one .npz and
records the returns they reach on the oracle's restated environments; tests/test_policy_anchor.py
holds the oracle (CPU) and the CUDA kernels (GPU) to those returns.  Corroboration, not a bit-level
pin -- a6 / a7 stay labelled accordingly in DESIGN.md.
"""
import io
import json
import os
import sys
import zipfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference/ns_gym"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "anchors", "reference_policies.npz")

DDQN = {}      # the shipped DDQN nets (benchmark_algorithms/DDQN/DDQN_models) do not solve their tasks on either side: not used
SB3 = {"acrobot_ppo": ("Acrobot-v1", "evaluate/evaluation_model_weights/AcrobotEnv/PPO/ppo_acrobot_default.zip",
                       "mlp_extractor.policy_net", "action_net", "tanh"),
       "acrobot_a2c": ("Acrobot-v1", "evaluate/evaluation_model_weights/AcrobotEnv/A2C/a2c_acrobot.zip",
                       "mlp_extractor.policy_net", "action_net", "tanh"),
       "pendulum_ppo": ("Pendulum-v1", "evaluate/evaluation_model_weights/PendulumEnv/PPO/ppo_pendulum.zip",
                        "mlp_extractor.policy_net", "action_net", "tanh"),
       "pendulum_ddpg": ("Pendulum-v1", "evaluate/evaluation_model_weights/PendulumEnv/DDPG/PendulumEnv.zip",
                         "actor.mu", None, "relu")}


def load_policies():
    import torch

    out = {}
    for name, (env_id, fn) in DDQN.items():        # DDQN.py:119-165: Linear / ReLU stack, greedy = argmax
        sd = torch.load(os.path.join(REF, "benchmark_algorithms/DDQN/DDQN_models", fn), map_location="cpu", weights_only=False)
        layers = [(sd[f"layers.{k}.weight"].numpy(), sd[f"layers.{k}.bias"].numpy()) for k in (0, 2, 4)]
        out[name] = dict(env_id=env_id, act="relu", head="argmax", layers=layers)
    for name, (env_id, fn, trunk, head, act) in SB3.items():
        z = zipfile.ZipFile(os.path.join(REF, fn))
        sd = torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu", weights_only=False)
        if head:        # PPO MlpPolicy: tanh trunk, linear action head, deterministic = argmax of the logits
            layers = [(sd[f"{trunk}.{k}.weight"].numpy(), sd[f"{trunk}.{k}.bias"].numpy()) for k in (0, 2)]
            layers.append((sd[f"{head}.weight"].numpy(), sd[f"{head}.bias"].numpy()))
            # discrete: deterministic = argmax of the logits; Box (Pendulum): the Gaussian's mean, clipped by the env
            out[name] = dict(env_id=env_id, act=act, head="argmax" if "Acrobot" in env_id else "mean", layers=layers)
        else:           # TD3/DDPG actor: ReLU trunk, tanh output in [-1, 1] scaled to the action range (+-2)
            layers = [(sd[f"{trunk}.{k}.weight"].numpy(), sd[f"{trunk}.{k}.bias"].numpy()) for k in (0, 2, 4)]
            out[name] = dict(env_id=env_id, act=act, head="tanh2", layers=layers)
    return out


def act(pol, obs):
    """Deterministic action of a converted policy for a batch of float32 observations [N, O]."""
    x = np.asarray(obs, dtype=np.float32)
    for k, (w, b) in enumerate(pol["layers"]):
        x = x @ w.T + b
        if k + 1 < len(pol["layers"]):
            x = np.maximum(x, 0) if pol["act"] == "relu" else np.tanh(x)
    if pol["head"] == "argmax":
        return np.argmax(x, axis=1)
    if pol["head"] == "mean":
        return np.clip(x, -2.0, 2.0)
    return 2.0 * np.tanh(x)


def oracle_returns(pol, episodes=30, seed=7):
    from oracle import gym_restated as G

    rets = []
    env = G.make(pol["env_id"])
    for ep in range(episodes):
        obs, _ = env.reset(seed=seed + ep)
        total, done = 0.0, False
        while not done:
            a = act(pol, np.asarray(obs)[None])[0]
            obs, r, term, trunc, _ = env.step(int(a) if pol["head"] == "argmax" else np.asarray(a, dtype=np.float32))
            total += float(r)
            done = term or trunc
        rets.append(total)
    return np.array(rets)


def save(pols, returns):
    flat = {}
    meta = {}
    for name, p in pols.items():
        for k, (w, b) in enumerate(p["layers"]):
            flat[f"{name}.w{k}"], flat[f"{name}.b{k}"] = w.astype(np.float32), b.astype(np.float32)
        meta[name] = dict(env_id=p["env_id"], act=p["act"], head=p["head"], n_layers=len(p["layers"]),
                          oracle_mean_return=float(returns[name].mean()), oracle_min_return=float(returns[name].min()),
                          oracle_returns=[float(x) for x in returns[name]])
    flat["_meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **flat)


def load_saved(path=OUT):
    z = np.load(path)
    meta = json.loads(bytes(z["_meta"]).decode())
    pols = {}
    for name, m in meta.items():
        pols[name] = dict(m, layers=[(z[f"{name}.w{k}"], z[f"{name}.b{k}"]) for k in range(m["n_layers"])])
    return pols


if __name__ == "__main__":
    pols = load_policies()
    rets = {}
    for name, p in pols.items():
        rets[name] = oracle_returns(p)
        print(f"{name:18s} {p['env_id']:16s} mean {rets[name].mean():9.2f}  min {rets[name].min():9.2f}  max {rets[name].max():9.2f}")
    save(pols, rets)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
