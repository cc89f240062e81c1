"""Batched non-stationary env on one GPU: the host-side mirror of the reference's wrapper
stack (``NSWrapper.step / reset``, ``ns_gym/base.py:296-410``) over the C ABI.

One ``NSVectorEnv`` owns N envs of one kind (one shard).  All per-env data lives in torch
tensors on the device (SoA, ``include/nsgym_b200.h``); one call to ``step`` is one kernel
launch.  ``step`` returns gymnasium-vector style batched results with the reference's
observation dict ``{"state", "env_change", "delta_change", "relative_time"}`` and the same
notification gating (``base.py:323-341``); ``step_raw`` is the same launch without building
the Python result (results stay in ``env.buffers``).
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Optional

import numpy as np
import torch

from . import native as nv
from .base import Reward
from .compile import BOX_ACTION, N_ACTIONS, CompiledProgram, compile_program, compile_rows


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _mix64(*words: int) -> int:
    """splitmix64-style hash of a few 64-bit words (planning-copy keys)."""
    h = 0x9E3779B97F4A7C15
    for w in words:
        h = (h ^ (w & (2**64 - 1))) & (2**64 - 1)
        h = (h * 0xBF58476D1CE4E5B9) & (2**64 - 1)
        h ^= h >> 31
        h = (h * 0x94D049BB133111EB) & (2**64 - 1)
        h ^= h >> 29
    return h


class _LazyInfo(dict):
    """``info`` dict whose expensive entries (callables) are evaluated on first access."""

    def __getitem__(self, key):
        v = dict.__getitem__(self, key)
        if callable(v):
            v = v()
            dict.__setitem__(self, key, v)
        return v

    def get(self, key, default=None):
        return self[key] if key in self else default


class NSVectorEnv:
    def __init__(self, env_id: str, tunable_params: dict, num_envs: int, *,
                 change_notification: bool = False, delta_change_notification: bool = False,
                 in_sim_change: bool = False, scalar_reward: bool = True,
                 persistent_params: bool = False, precision: str = "fp32",
                 autoreset: str = "next_step", seed: int = 0, env_id_offset: int = 0,
                 device: Any = None, want_obs: Optional[bool] = None, want_delta: Optional[bool] = None,
                 rows=None, pools=None, bucket: bool = False, **env_kwargs):
        if delta_change_notification:                         # base.py:252-255
            assert change_notification, (
                "If change_notification is True, delta_change_notification must be True")
        if not torch.cuda.is_available():
            raise nv.NsgymError("ns_gym_b200 needs a CUDA device: there is no CPU execution path")
        # what a planning copy is rebuilt from (get_planning_env)
        self._ctor = dict(env_id=env_id, tunable_params=tunable_params, num_envs=int(num_envs),
                          change_notification=change_notification,
                          delta_change_notification=delta_change_notification, in_sim_change=in_sim_change,
                          scalar_reward=scalar_reward, persistent_params=persistent_params, precision=precision,
                          seed=seed, device=device, want_obs=want_obs, want_delta=want_delta, rows=rows,
                          pools=pools, env_kwargs=dict(env_kwargs))
        self.env_order = None      # bucketed heterogeneous batches: storage position -> caller's env index
        self.env_id_offset = int(env_id_offset)
        self._n_copies = 0         # planning copies taken so far (mixed into their Philox key)
        self.lib = nv.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.num_envs = int(num_envs)
        compile_kw = dict(precision=precision, autoreset=autoreset, seed=seed, env_id_offset=env_id_offset,
                          persistent_params=persistent_params, **env_kwargs)
        self.rows = None
        if rows is not None:
            # heterogeneous batch: `tunable_params` is either one dict per env, or -- with `rows` a
            # prepared NsgymSlot array [num_envs, n_slots] -- the key-set template of the batch
            if rows is True:
                self.program, self.rows = compile_rows(env_id, tunable_params, **compile_kw)
                pools = self.program.pool_lists
                self._ctor["tunable_params"] = tunable_params[0]      # key-set template for planning copies
            else:
                self.program = compile_program(env_id, tunable_params, num_envs, **compile_kw)
                self.rows = np.ascontiguousarray(rows)
                if self.rows.shape != (self.num_envs, len(tunable_params)) or self.rows.dtype.itemsize != C.sizeof(nv.NsgymSlot):
                    raise ValueError("rows must be an NsgymSlot array of shape [num_envs, n_slots]")
                for j in range(len(tunable_params)):
                    if (self.rows["istate_plane"][:, j] >= 0).any() and self.program.spec.slots[j].istate_plane < 0:
                        planes = 1 + max((self.program.spec.slots[q].istate_plane for q in range(len(tunable_params))),
                                         default=-1)
                        self.program.spec.slots[j].istate_plane = planes
                        sel = self.rows["istate_plane"][:, j] >= 0
                        self.rows["istate_plane"][sel, j] = planes
            if len(self.rows) != self.num_envs:
                raise ValueError(f"{len(self.rows)} rows for {self.num_envs} envs")
            self._ctor["pools"] = pools
            if bucket:
                # bucket the envs by opcode signature: warps of the heterogeneous kernel then
                # hold envs that take the same branches.  Storage position k holds the caller's env
                # env_order[k]; actions / results are in storage order.
                from .compile import row_signature
                self.env_order = np.argsort(row_signature(self.rows), kind="stable")
                self.rows = np.ascontiguousarray(self.rows[self.env_order])
            if pools is not None:
                # prepared rows index their own pools (pool_f doubles, pool_i int32, bitmap uint32)
                from .compile import _Pools, _attach_pools
                pl = _Pools()
                pl.f, pl.i, pl.bits = list(map(float, pools[0])), list(map(int, pools[1])), list(map(int, pools[2]))
                self.program._keepalive.clear()
                _attach_pools(self.program, pl)
        else:
            self.program: CompiledProgram = compile_program(env_id, tunable_params, num_envs, **compile_kw)
        self.keys = list(self.program.keys)
        self.change_notification = change_notification
        self.delta_change_notification = delta_change_notification
        self.in_sim_change = in_sim_change
        self.scalar_reward = scalar_reward
        self.persistent_params = persistent_params
        self.frozen = False
        self.is_sim_env = False
        self.has_reset = False
        self.precision = precision if not self.program.is_grid else "fp64"
        self.real = torch.float64 if self.program.spec.precision == nv.F64 else torch.float32

        with torch.cuda.device(self.device):
            h = C.c_void_p()
            if self.rows is not None:
                nv.check(self.lib.nsgym_create_rows(C.byref(self.program.spec),
                                                    C.c_void_p(self.rows.ctypes.data), C.byref(h)),
                         "nsgym_create_rows")
            else:
                nv.check(self.lib.nsgym_create(C.byref(self.program.spec), C.byref(h)), "nsgym_create")
            self._h = h
            kind = self.program.env_kind
            if want_obs is None:
                # fp32 CartPole / MountainCar: the observation IS the stored state
                want_obs = (not self.program.is_grid) and (
                    self.real == torch.float64 or self.program.obs_words != self.program.state_words)
            if want_delta is None:
                want_delta = delta_change_notification
            lay = nv.NsgymLayout()
            nv.check(self.lib.nsgym_layout(self._h, int(want_delta), int(want_obs), C.byref(lay)), "nsgym_layout")
            self.layout = lay
            self.bytes_per_step = float(lay.bytes_per_step)
            self.row_bytes_per_env = float(lay.row_bytes_per_env)
            n, dev = self.num_envs, self.device
            b: dict[str, Optional[torch.Tensor]] = {}
            if self.program.is_grid:
                b["state"] = torch.zeros(n, dtype=torch.int32, device=dev)
            else:
                b["state"] = torch.zeros((n, lay.state_words), dtype=self.real, device=dev)
            b["theta"] = (torch.zeros((lay.theta_planes, n), dtype=self.real, device=dev)
                          if lay.theta_planes else None)
            b["t"] = torch.zeros(n, dtype=torch.int32, device=dev)
            b["istate"] = torch.zeros((lay.n_istate, n), dtype=torch.int32, device=dev) if lay.n_istate else None
            b["action"] = torch.zeros(n, dtype=self.real if kind in BOX_ACTION else torch.int32, device=dev)
            b["reward"] = torch.zeros(n, dtype=torch.float32, device=dev)
            b["flags"] = torch.zeros(n, dtype=torch.uint8, device=dev)
            b["change"] = torch.zeros(n, dtype=torch.uint8, device=dev)
            b["delta"] = (torch.zeros((len(self.keys), n), dtype=self.real, device=dev)
                          if (want_delta and self.keys) else None)
            b["obs"] = (torch.zeros((n, lay.obs_words), dtype=torch.float32, device=dev)
                        if lay.obs else None)
            self.buffers = b
            cb = nv.NsgymBuffers(
                d_state=b["state"].data_ptr(), d_theta=0 if b["theta"] is None else b["theta"].data_ptr(),
                d_t=b["t"].data_ptr(), d_istate=0 if b["istate"] is None else b["istate"].data_ptr(),
                d_action=b["action"].data_ptr(), d_reward=b["reward"].data_ptr(),
                d_flags=b["flags"].data_ptr(), d_change=b["change"].data_ptr(),
                d_delta=0 if b["delta"] is None else b["delta"].data_ptr(),
                d_obs=0 if b["obs"] is None else b["obs"].data_ptr())
            nv.check(self.lib.nsgym_bind(self._h, C.byref(cb)), "nsgym_bind")
        self._zeros_u8 = torch.zeros(n, dtype=torch.uint8, device=self.device)
        self._zeros_real = torch.zeros(n, dtype=self.real, device=self.device)
        # result packaging (nsgym_unpack): written by ONE launch per step and handed out as they are --
        # the tensors a step returns are valid until the next step of this batch
        self._terminated = torch.zeros(n, dtype=torch.bool, device=self.device)
        self._truncated = torch.zeros(n, dtype=torch.bool, device=self.device)
        self._was_reset = torch.zeros(n, dtype=torch.bool, device=self.device)
        self._rel_time = torch.zeros(n, dtype=torch.int32, device=self.device)
        self._env_change = torch.zeros((max(len(self.keys), 1), n), dtype=torch.uint8, device=self.device)

    def to_storage(self, x):
        """Reorder a per-env array / tensor from the caller's env order to storage order."""
        return x if self.env_order is None else x[torch.as_tensor(self.env_order, device=x.device)
                                                  if torch.is_tensor(x) else self.env_order]

    def to_caller(self, x):
        """Reorder a per-env array / tensor (leading env axis) from storage order back to the caller's."""
        if self.env_order is None:
            return x
        inv = np.empty_like(self.env_order)
        inv[self.env_order] = np.arange(len(self.env_order))
        return x[torch.as_tensor(inv, device=x.device) if torch.is_tensor(x) else inv]

    @classmethod
    def heterogeneous(cls, env_id: str, params_per_env, **kwargs):
        """BASELINE config C4: one ``tunable_params`` dict PER ENV (same parameter names, per-env
        schedulers / update functions / coefficients).  ``bucket=True`` stores the envs sorted by
        opcode signature (``env_order``; see ``to_storage`` / ``to_caller``)."""
        params_per_env = list(params_per_env)
        return cls(env_id, params_per_env, len(params_per_env), rows=True, **kwargs)

    # ------------------------------------------------------------------------------------
    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and getattr(self, "lib", None) is not None:
            try:
                self.lib.nsgym_destroy(h)
            except Exception:
                pass
            self._h = None

    close = __del__

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def skip_updates(self) -> bool:
        """Planning env with ``in_sim_change`` False: theta is frozen (classic_control.py:70-75)."""
        return self.is_sim_env and not self.in_sim_change

    @property
    def launch_count(self) -> int:
        return int(self.lib.nsgym_launch_count(self._h))

    @property
    def action_space_n(self):
        return N_ACTIONS.get(self.program.env_kind)

    @property
    def action_space(self):
        """``Discrete(n)`` or ``Box(low, high)`` of the base env; ``sample()`` draws one action per env."""
        from .spaces import Box, Discrete
        kind = self.program.env_kind
        if kind in BOX_ACTION:
            lo, hi = BOX_ACTION[kind]
            return Box(lo, hi, (1,), self.num_envs, self.device, self.real)
        return Discrete(N_ACTIONS[kind], self.num_envs, self.device)

    single_action_space = action_space

    @property
    def t(self):
        """``NSWrapper.t`` per env (``base.py:314``)."""
        return self.relative_time()

    # ------------------------------------------------------------------------------------
    def reset(self, *, seed: Optional[int] = None, options=None, mask: Optional[torch.Tensor] = None,
              inject_uniform: Optional[torch.Tensor] = None):
        """``NSWrapper.reset`` for the batch (``base.py:365-410``).  ``seed`` re-keys the
        Philox streams (every env's stream is keyed by (seed, global env id))."""
        with torch.cuda.device(self.device):
            if seed is not None:
                # reset(seed=s) is reproducible (base.py:386-388, 412-421 reseed every generator from
                # s): new key, and the Philox step counter restarts, so the same seed replays the
                # same draws whatever ran before
                self.lib.nsgym_set_seed(self._h, int(seed) & (2**64 - 1))
                if mask is None:
                    self.lib.nsgym_set_step_index(self._h, 0)
            m = None
            if mask is not None:
                m = mask.to(device=self.device, dtype=torch.uint8).contiguous()
            nv.check(self.lib.nsgym_reset(self._h, _ptr(m), _ptr(inject_uniform), self._stream()), "nsgym_reset")
        self.has_reset = True
        n_keys = len(self.keys)
        zeros_c = {k: self._zeros_u8 for k in self.keys}
        zeros_d = {k: self._zeros_real for k in self.keys}
        obs = {"state": self.observation(), "env_change": zeros_c, "delta_change": zeros_d,
               "relative_time": self.relative_time()}
        info = {"Ground Truth Env Change": dict(zeros_c), "Ground Truth Delta Change": dict(zeros_d)}
        del n_keys
        return obs, info

    def step_raw(self, actions: Optional[torch.Tensor] = None, inject_uniform=None, inject_normal=None):
        """One kernel launch; results stay in ``self.buffers``."""
        nv.check(self.lib.nsgym_step(self._h, _ptr(actions), _ptr(inject_uniform), _ptr(inject_normal),
                                     int(self.skip_updates), self._stream()), "nsgym_step")

    def step(self, actions, inject_uniform=None, inject_normal=None):
        """``<wrapper>.step`` for the batch -> ``(obs, reward, terminated, truncated, info)``."""
        want = self.buffers["action"]
        if not torch.is_tensor(actions):
            actions = torch.as_tensor(np.asarray(actions))
        actions = actions.to(device=self.device, dtype=want.dtype).reshape(self.num_envs).contiguous()
        with torch.cuda.device(self.device):
            self.step_raw(actions, inject_uniform, inject_normal)
        return self._package()

    # ---- result packaging (NSWrapper.step, base.py:314-363) ----
    def observation(self) -> torch.Tensor:
        b = self.buffers
        if b["obs"] is not None:
            return b["obs"]
        return b["state"]

    def relative_time(self) -> torch.Tensor:
        return self.buffers["t"] & nv.T_TIME_MASK

    def ground_truth_change(self) -> dict:
        c = self.buffers["change"]
        return {k: (c >> j) & 1 for j, k in enumerate(self.keys)}

    def constraint_violations(self, warn: bool = False) -> int:
        """Number of envs whose last step had a fired update REJECTED by the constraint checker
        (theta kept its old value): the batch counterpart of the reference's per-env
        ``ConstraintViolationWarning`` (classic_control.py:87-92).  Reads the flag bytes (one device
        reduction + a host sync), so it is on demand; ``warn=True`` also issues the warning.
        ``distributed.EpisodeStats`` accumulates the same count without a sync."""
        n = int(((self.buffers["flags"] & nv.FLAG_REJECTED) != 0).sum().item())
        if warn and n:
            import warnings

            from .wrappers import ConstraintViolationWarning
            warnings.warn(f"{n} of {self.num_envs} envs rejected a parameter update that violates the "
                          "environment's constraints; those parameters keep their previous value",
                          ConstraintViolationWarning)
        return n

    def ground_truth_delta(self) -> dict:
        d = self.buffers["delta"]
        if d is None:
            return {k: None for k in self.keys}
        return {k: d[j] for j, k in enumerate(self.keys)}

    def _package(self):
        b = self.buffers
        with torch.cuda.device(self.device):
            nv.check(self.lib.nsgym_unpack(self._h, _ptr(self._terminated), _ptr(self._truncated),
                                           _ptr(self._was_reset), _ptr(self._rel_time), _ptr(self._env_change),
                                           self._stream()), "nsgym_unpack")
        muted = self.frozen or (self.is_sim_env and not self.in_sim_change)
        gt_c = {k: self._env_change[j] for j, k in enumerate(self.keys)}
        gt_d = self.ground_truth_delta()
        zc = {k: self._zeros_u8 for k in self.keys}
        zd = {k: self._zeros_real for k in self.keys}
        env_change = zc if (not self.change_notification or muted) else gt_c
        delta_change = zd if (not self.delta_change_notification or muted) else gt_d
        rel = self._rel_time
        obs = {"state": self.observation(), "env_change": env_change, "delta_change": delta_change,
               "relative_time": rel}
        reward: Any = b["reward"]
        if not self.scalar_reward:
            reward = Reward(reward=b["reward"], env_change=env_change, delta_change=delta_change,
                            relative_time=rel)
        info = _LazyInfo({"Ground Truth Env Change": gt_c, "Ground Truth Delta Change": gt_d,
                          "was_reset": self._was_reset})
        if self.program.is_grid:
            info["transition_prob"] = self.transition_prob      # evaluated on first access
        else:
            info["prob"] = 1.0                                   # classic_control.py:98
        return obs, reward, self._terminated, self._truncated, info

    # ---- parameter views ----
    def theta(self) -> dict:
        """Current value of every bound parameter: ``{name: tensor[N]}`` (scalars) or
        ``{name: tensor[D, N]}`` (slip distributions, the wrapper's ``transition_prob``)."""
        th = self.buffers["theta"]
        if not self.program.is_grid:
            return {k: th[j] for j, k in enumerate(self.keys)}
        return self.transition_prob()

    def transition_prob(self) -> dict:
        """``transition_prob`` of the gridworld wrappers (toy_text.py:178-187, 362-373).  For
        FrozenLake / CliffWalking the stored planes are the probabilities baked into the
        sampling table; until the first fire after a reset the wrapper's ``transition_prob``
        is the initial distribution (stale-table rule)."""
        th, D = self.buffers["theta"], self.program.n_dist
        out = {}
        for j, k in enumerate(self.keys):
            planes = th[j * D:(j + 1) * D]
            if self.program.env_kind != nv.ENV_BRIDGE:
                fresh = (self.buffers["t"] & nv.T_TABLE_FRESH) != 0
                idx = self.program.spec.slots[j].theta_index
                init = torch.tensor([self.program.spec.theta_init[idx][q] for q in range(D)],
                                    dtype=torch.float64, device=self.device)[:, None]
                planes = torch.where(fresh[None, :], planes, init)
            out[k] = planes
        return out

    # ---- fused K-step rollout ----
    def policy_shape(self, per_env: bool = False):
        """Shape of a device-side rollout policy: classic control ``[A, O + 1]`` float32 (row a: O
        observation weights, then the bias; A = 1 for Box action spaces), gridworlds ``[n_cells]``
        uint8 (action per cell); one leading ``num_envs`` axis when every env has its own."""
        kind = self.program.env_kind
        if kind in (nv.ENV_FROZENLAKE, nv.ENV_CLIFFWALKING, nv.ENV_BRIDGE):
            shape = (int(self.program.spec.nrow) * int(self.program.spec.ncol),)
        else:
            n_obs = {nv.ENV_CARTPOLE: 4, nv.ENV_ACROBOT: 6, nv.ENV_MOUNTAINCAR: 2, nv.ENV_MOUNTAINCAR_CONT: 2,
                     nv.ENV_PENDULUM: 3}[kind]
            shape = (N_ACTIONS.get(kind, 1), n_obs + 1)
        return ((self.num_envs,) + shape) if per_env else shape

    def rollout(self, k_steps: int, gamma: float = 1.0, returns: Optional[torch.Tensor] = None,
                lengths: Optional[torch.Tensor] = None, policy: Optional[torch.Tensor] = None):
        """K fused steps, state in registers, under a device-side policy: uniform-random actions, or
        (``policy`` given, see ``policy_shape``) a linear policy on the observation -- ``argmax`` of
        ``W obs + b`` for Discrete action spaces, the score itself for Box ones; a per-cell action
        table for gridworlds.  A leading ``num_envs`` axis gives every env its own policy."""
        if returns is None:
            returns = torch.zeros(self.num_envs, dtype=torch.float32, device=self.device)
        if lengths is None:
            lengths = torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            if policy is None:
                nv.check(self.lib.nsgym_rollout(self._h, int(k_steps), 0, float(gamma), _ptr(returns), _ptr(lengths),
                                                int(self.skip_updates), self._stream()), "nsgym_rollout")
            else:
                shared, own = self.policy_shape(False), self.policy_shape(True)
                if tuple(policy.shape) not in (shared, own):
                    raise ValueError(f"policy shape {tuple(policy.shape)}: expected {shared} or {own}")
                want = torch.uint8 if len(shared) == 1 else torch.float32
                if policy.dtype != want or policy.device != self.device or not policy.is_contiguous():
                    raise ValueError(f"policy must be a contiguous {want} tensor on {self.device}")
                nv.check(self.lib.nsgym_rollout_linear(self._h, int(k_steps), _ptr(policy),
                                                       int(tuple(policy.shape) == own and own != shared), float(gamma),
                                                       _ptr(returns), _ptr(lengths), int(self.skip_updates),
                                                       self._stream()), "nsgym_rollout_linear")
        return returns, lengths

    # ---- host-buffer (end-to-end) step through the C ABI ----
    def make_host_io(self, want_state: bool = True):
        """Pinned host buffers for ``step_host``: (actions, outputs dict)."""
        b = self.buffers
        pin = dict(pin_memory=True)
        act = torch.zeros(self.num_envs, dtype=b["action"].dtype, **pin)
        out = {"reward": torch.zeros(self.num_envs, dtype=torch.float32, **pin),
               "flags": torch.zeros(self.num_envs, dtype=torch.uint8, **pin),
               "change": torch.zeros(self.num_envs, dtype=torch.uint8, **pin)}
        if b["obs"] is not None:
            out["obs"] = torch.zeros(tuple(b["obs"].shape), dtype=torch.float32, **pin)
        elif want_state:
            out["state"] = torch.zeros(tuple(b["state"].shape), dtype=b["state"].dtype, **pin)
        return act, out

    def step_host(self, h_actions: torch.Tensor, h_out: dict, n_chunks: int = 8):
        """``nsgym_step_host``: actions host->device, step, results device->host (synchronous)."""
        ho = nv.NsgymHostOut(
            h_reward=h_out["reward"].data_ptr(), h_flags=h_out["flags"].data_ptr(),
            h_change=h_out["change"].data_ptr() if "change" in h_out else 0,
            h_delta=0, h_state=h_out["state"].data_ptr() if "state" in h_out else 0,
            h_obs=h_out["obs"].data_ptr() if "obs" in h_out else 0)
        with torch.cuda.device(self.device):
            nv.check(self.lib.nsgym_step_host(self._h, C.c_void_p(h_actions.data_ptr()), C.byref(ho),
                                              int(n_chunks), self._stream()), "nsgym_step_host")

    def host_bytes_per_step(self, h_actions: torch.Tensor, h_out: dict):
        h2d = h_actions.numel() * h_actions.element_size()
        d2h = sum(t.numel() * t.element_size() for t in h_out.values())
        return h2d, d2h

    # ---- planning copies (classic_control.py:120-186, toy_text.py:471-511, 669-711) ----
    # TimeLimit of the gym.make chain the reference's __deepcopy__ builds
    _PLANNING_LIMIT = {nv.ENV_FROZENLAKE: 100, nv.ENV_CLIFFWALKING: 1000, nv.ENV_BRIDGE: 1000}

    def get_planning_env(self, fanout: int = 1, seed: Optional[int] = None) -> "NSVectorEnv":
        """Batched ``get_planning_env()``: a new batch whose lanes ``r * fanout + k`` are copies of
        env ``r`` (state, episode time, list cursors; parameters too when the agent is told about
        them -- ``delta_change_notification`` or a copy of a copy -- else the initial parameters).
        The copy is a simulation env (``is_sim_env``): with ``in_sim_change`` False its parameters
        are frozen and notifications muted.  ``fanout`` > 1 gives every root ``fanout`` lanes for
        Monte-Carlo rollouts (``rollout``), the device-side form of MCTS's per-simulation deepcopy
        (``benchmark_algorithms/MCTS.py:130-131``).  Philox streams are re-keyed (``seed``), as the
        reference reseeds the copy's generators."""
        assert self.has_reset, "The environment must be reset before getting the planning environment."
        c = self._ctor
        kw = dict(c["env_kwargs"])
        # the copy sits in the gym.make chain __deepcopy__ builds: registered TimeLimit for classic
        # control, FrozenLake-v1's 100, max_episode_steps=1000 for CliffWalking / Bridge
        kw.pop("max_episode_steps", None)
        limit = self._PLANNING_LIMIT.get(self.program.env_kind)
        if limit is not None:
            kw["max_episode_steps"] = limit
        rows, tp = c["rows"], c["tunable_params"]
        if rows is not None:                          # per-env rows (already in storage order): repeat them
            rows = np.repeat(self.rows, fanout, axis=0)
        if seed is None:
            # _reseed_planning_env_rngs (base.py:433-441) gives every copy fresh generators: the key
            # mixes the root's key, its Philox step counter AND the number of copies taken so far, so
            # two copies of the same root step (MCTS.py:130: one deepcopy per simulation) draw
            # different streams; `seed=` pins it for reproducible tests
            self._n_copies += 1
            seed = _mix64(int(self.program.spec.seed), int(self.lib.nsgym_step_index(self._h)), self._n_copies)
        plan = type(self).__new__(type(self))
        NSVectorEnv.__init__(
            plan, c["env_id"], tp, self.num_envs * fanout, change_notification=c["change_notification"],
            delta_change_notification=c["delta_change_notification"], in_sim_change=c["in_sim_change"],
            scalar_reward=c["scalar_reward"], persistent_params=c["persistent_params"], precision=c["precision"],
            autoreset="none", seed=seed, env_id_offset=self.env_id_offset * fanout, device=self.device,
            want_obs=c["want_obs"],
            want_delta=c["want_delta"], rows=rows, pools=c["pools"], **kw)
        theta_from_init = not (self.is_sim_env or self.delta_change_notification)
        with torch.cuda.device(self.device):
            nv.check(self.lib.nsgym_fanout(self._h, plan._h, int(fanout), int(theta_from_init), self._stream()),
                     "nsgym_fanout")
        plan.lib.nsgym_set_step_index(plan._h, int(self.lib.nsgym_step_index(self._h)))
        plan.is_sim_env = True
        plan.has_reset = True
        plan.fanout = fanout
        return plan

    def transition_table(self, T: Optional[int] = None, env: int = 0):
        """Time-indexed transition table of env ``env`` over NS times ``0..T-1`` of an episode (the
        view of ``unwrapped.P`` / ``Bridge.transition_matrix`` step after step): ``prob [T, S, 4, D]``
        plus the time-invariant ``next / reward / done [S, 4, D]``."""
        if not self.program.is_grid:
            raise ValueError("transition tables exist for the gridworld environments only")
        spec = self.program.spec
        T = int(T if T is not None else (spec.max_episode_steps or 100))
        S, D = spec.nrow * spec.ncol, spec.n_dist
        out = {"prob": torch.zeros((T, S, 4, D), dtype=torch.float64, device=self.device),
               "next": torch.zeros((S, 4, D), dtype=torch.int32, device=self.device),
               "reward": torch.zeros((S, 4, D), dtype=torch.float32, device=self.device),
               "done": torch.zeros((S, 4, D), dtype=torch.uint8, device=self.device)}
        with torch.cuda.device(self.device):
            nv.check(self.lib.nsgym_transition_table(self._h, int(env), T, _ptr(out["prob"]), _ptr(out["next"]),
                                                     _ptr(out["reward"]), _ptr(out["done"]), self._stream()),
                     "nsgym_transition_table")
        return out

    def set_option(self, name: str, value) -> None:
        """Handle options (``nsgym_set_option``): ``"general_kernels"`` forces the general kernel
        instantiations instead of the lean ones picked for this program (same results);
        ``"specialize"`` (-1 auto / 0 / 1) runs a lean program through a step kernel compiled at run
        time for exactly this program (NVRTC; auto = batches of >= 32768 envs)."""
        opt = {"general_kernels": nv.OPT_GENERAL_KERNELS, "specialize": nv.OPT_SPECIALIZE}[name]
        nv.check(self.lib.nsgym_set_option(self._h, opt, int(value)), "nsgym_set_option")

    @property
    def last_kernel_specialized(self) -> bool:
        """Did the last step / rollout launch run a program-specialised kernel?"""
        return bool(self.lib.nsgym_last_kernel_specialized(self._h))

    def snapshot(self):
        """Device copy of everything a step mutates; ``restore`` rewinds the batch to it."""
        buf = torch.empty(int(self.lib.nsgym_snapshot_bytes(self._h)), dtype=torch.uint8, device=self.device)
        info = nv.NsgymSnapshotInfo()
        with torch.cuda.device(self.device):
            nv.check(self.lib.nsgym_snapshot(self._h, _ptr(buf), C.byref(info), self._stream()), "nsgym_snapshot")
        return buf, info

    def restore(self, snap):
        buf, info = snap
        with torch.cuda.device(self.device):
            nv.check(self.lib.nsgym_restore(self._h, _ptr(buf), C.byref(info), self._stream()), "nsgym_restore")

    # ---- notification control (base.py:443-458) ----
    def freeze(self, mode: bool = True):
        if not isinstance(mode, bool):
            raise TypeError(f"Expected mode to be a boolean, got {type(mode)}")
        self.frozen = mode
        return self

    def unfreeze(self):
        return self.freeze(False)

    def get_default_params(self):
        from .base import TUNABLE_PARAMS
        return TUNABLE_PARAMS[self.program.env_class]


class MixedVectorEnv:
    """Several env kinds in one logical batch (BASELINE config C4: CartPole + FrozenLake with
    per-env rows).  Envs are bucketed by kind: one ``NSVectorEnv`` shard per kind, each stepped by
    its own kernel launch on the caller's stream; global env ids (Philox streams) are contiguous
    per shard.  ``step`` takes / returns one entry per shard, in shard order."""

    def __init__(self, shards):
        self.shards = list(shards)
        self.num_envs = sum(s.num_envs for s in self.shards)

    @property
    def bytes_per_step(self) -> float:
        """algorithmic bytes per env-step averaged over the batch"""
        return sum(s.bytes_per_step * s.num_envs for s in self.shards) / self.num_envs

    @property
    def launch_count(self) -> int:
        return sum(s.launch_count for s in self.shards)

    def reset(self, *, seed=None, **kw):
        return [s.reset(seed=seed, **kw) for s in self.shards]

    def step_raw(self, actions=None):
        """One host call (``nsgym_step_many``) steps every shard, back to back on the current stream."""
        n = len(self.shards)
        handles = (C.c_void_p * n)(*[s._h for s in self.shards])
        acts = (C.c_void_p * n)(*[None if actions is None or actions[k] is None else actions[k].data_ptr()
                                  for k in range(n)])
        s0 = self.shards[0]
        if any(s.skip_updates != s0.skip_updates or s.device != s0.device for s in self.shards):
            raise ValueError("the shards of a mixed batch share one device and one planning mode")
        with torch.cuda.device(s0.device):
            nv.check(s0.lib.nsgym_step_many(handles, acts, n, int(s0.skip_updates), s0._stream()), "nsgym_step_many")

    def step(self, actions):
        return [s.step(a) for s, a in zip(self.shards, actions)]
