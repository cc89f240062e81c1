"""Lower a ``tunable_params`` dict to the opcode-and-coefficient table of the step kernel.

Input: the same objects a user of the reference builds -- ``{param_name: UpdateFn}`` where
every update function owns a ``.scheduler`` (``ns_gym/base.py:222-265``).  The compiler is
duck-typed on class and attribute names, so instances of the reference's own
classes compile as well as ``ns_gym_b200``'s descriptions.

Output: ``CompiledProgram`` wrapping a ``native.NsgymSpec`` (``include/nsgym_b200.h``): one
``NsgymSlot`` per bound parameter, in dict insertion order (the order the reference iterates
in, ``classic_control.py:80-85``), plus the pools the slots reference.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np

from . import native as nv
from .base import TUNABLE_PARAMS

FROZEN_LAKE_MAPS = {
    "4x4": ["SFFF", "FHFH", "FFFH", "HFFG"],
    "8x8": ["SFFFFFFF", "FFFFFFFF", "FFFHFFFF", "FFFFFHFF", "FFFHFFFF", "FHHFFFHF", "FHFFHFHF",
            "FFFHFFFG"],
}
BRIDGE_MAP = ["HHHHHHHH", "FFFFFHHH", "GFHFSFFG", "FFFFFHHH", "HHHHHHHH"]   # envs/Bridge.py:12

LE0, LT0 = nv.CONS_REJECT_LE0, nv.CONS_REJECT_LT0

# env id -> (kind, reference class name, TimeLimit, parameter order, constraint per parameter)
ENV_TABLE = {
    "CartPole-v1": (nv.ENV_CARTPOLE, "CartPoleEnv", 500),
    "CartPole-v0": (nv.ENV_CARTPOLE, "CartPoleEnv", 200),
    "Acrobot-v1": (nv.ENV_ACROBOT, "AcrobotEnv", 500),
    "MountainCar-v0": (nv.ENV_MOUNTAINCAR, "MountainCarEnv", 200),
    "MountainCarContinuous-v0": (nv.ENV_MOUNTAINCAR_CONT, "Continuous_MountainCarEnv", 999),
    "Pendulum-v1": (nv.ENV_PENDULUM, "PendulumEnv", 200),
    "FrozenLake-v1": (nv.ENV_FROZENLAKE, "FrozenLakeEnv", 100),
    "FrozenLake8x8-v1": (nv.ENV_FROZENLAKE, "FrozenLakeEnv", 200),
    "CliffWalking-v1": (nv.ENV_CLIFFWALKING, "CliffWalkingEnv", 0),
    "ns_gym/Bridge-v0": (nv.ENV_BRIDGE, "Bridge", 100),
}

THETA_ORDER = {
    nv.ENV_CARTPOLE: ["gravity", "masscart", "masspole", "force_mag", "tau", "length"],
    nv.ENV_ACROBOT: ["dt", "LINK_LENGTH_1", "LINK_LENGTH_2", "LINK_MASS_1", "LINK_MASS_2",
                     "LINK_COM_POS_1", "LINK_COM_POS_2", "LINK_MOI"],
    nv.ENV_MOUNTAINCAR: ["gravity", "force"],
    nv.ENV_MOUNTAINCAR_CONT: ["power"],
    nv.ENV_PENDULUM: ["m", "l", "dt", "g"],
    nv.ENV_FROZENLAKE: ["P"],
    nv.ENV_CLIFFWALKING: ["P"],
    nv.ENV_BRIDGE: ["P", "P_left", "P_right"],
}

# wrappers/classic_control.py:193-422
CONSTRAINTS = {
    nv.ENV_CARTPOLE: {"length": LE0, "masscart": LE0, "masspole": LE0, "gravity": LT0},
    nv.ENV_ACROBOT: {"LINK_LENGTH_1": nv.CONS_ACRO_LENGTH1, "LINK_LENGTH_2": LE0, "LINK_MASS_1": LE0,
                     "LINK_MASS_2": LE0, "LINK_COM_POS_1": nv.CONS_ACRO_COM,
                     "LINK_COM_POS_2": nv.CONS_ACRO_COM},
    nv.ENV_MOUNTAINCAR: {"gravity": LE0, "force": LE0},
    nv.ENV_MOUNTAINCAR_CONT: {"power": LE0},
    nv.ENV_PENDULUM: {"m": LE0, "l": LE0, "dt": LE0, "g": LT0},
}
ACRO_PARTNER = {"LINK_LENGTH_1": "LINK_COM_POS_1", "LINK_COM_POS_1": "LINK_LENGTH_1",
                "LINK_COM_POS_2": "LINK_LENGTH_2"}

STATE_WORDS = {nv.ENV_CARTPOLE: 4, nv.ENV_ACROBOT: 4, nv.ENV_MOUNTAINCAR: 2, nv.ENV_MOUNTAINCAR_CONT: 2,
               nv.ENV_PENDULUM: 2, nv.ENV_FROZENLAKE: 1, nv.ENV_CLIFFWALKING: 1, nv.ENV_BRIDGE: 1}
OBS_WORDS = {nv.ENV_CARTPOLE: 4, nv.ENV_ACROBOT: 6, nv.ENV_MOUNTAINCAR: 2, nv.ENV_MOUNTAINCAR_CONT: 2,
             nv.ENV_PENDULUM: 3}
N_ACTIONS = {nv.ENV_CARTPOLE: 2, nv.ENV_ACROBOT: 3, nv.ENV_MOUNTAINCAR: 3, nv.ENV_FROZENLAKE: 4,
             nv.ENV_CLIFFWALKING: 4, nv.ENV_BRIDGE: 4}
BOX_ACTION = {nv.ENV_MOUNTAINCAR_CONT: (-1.0, 1.0), nv.ENV_PENDULUM: (-2.0, 2.0)}
GRID_KINDS = (nv.ENV_FROZENLAKE, nv.ENV_CLIFFWALKING, nv.ENV_BRIDGE)

_STATEFUL_UPDATES = {
    "RandomWalk", "RandomWalkWithDrift", "RandomWalkWithDriftAndTrend", "OrnsteinUhlenbeck",
    "BoundedRandomWalk", "RandomCategorical", "LCBoundedDistrubutionUpdate", "StepWiseUpdate", "CyclicUpdate",
    "DistributionStepWiseUpdate", "DistributionCyclicUpdate"}
_STATEFUL_SCHEDS = {"RandomScheduler", "DecayingProbabilityScheduler", "MemorylessScheduler",
                    "CustomScheduler"}


class CompileError(ValueError):
    pass


@dataclass
class CompiledProgram:
    spec: nv.NsgymSpec
    env_id: str
    env_kind: int
    env_class: str
    keys: list                     # bound parameter names, slot order
    precision: int
    n_dist: int
    _keepalive: list = field(default_factory=list, repr=False)

    @property
    def is_grid(self) -> bool:
        return self.env_kind in GRID_KINDS

    @property
    def state_words(self) -> int:
        return STATE_WORDS[self.env_kind]

    @property
    def obs_words(self) -> int:
        return OBS_WORDS.get(self.env_kind, 0)


def _int_bounds(start, end):
    """``start <= t <= end`` for integer t  ->  inclusive int32 range."""
    lo = nv.INT32_MAX if (isinstance(start, float) and math.isinf(start) and start > 0) else \
        max(int(math.ceil(start)), 0)
    if isinstance(end, float) and math.isinf(end):
        hi = nv.INT32_MAX if end > 0 else -1
    else:
        hi = int(math.floor(end))
    return min(lo, nv.INT32_MAX), max(min(hi, nv.INT32_MAX), -1)


class _Pools:
    def __init__(self):
        self.f: list = []
        self.i: list = []
        self.bits: list = []

    def add_f(self, values):
        off = len(self.f)
        self.f.extend(float(v) for v in values)
        return off

    def add_i(self, values):
        off = len(self.i)
        self.i.extend(int(v) for v in values)
        return off

    def add_bitmap(self, times, n_bits):
        words = [0] * ((n_bits + 31) // 32)
        for t in times:
            words[t >> 5] |= 1 << (t & 31)
        off = len(self.bits)
        self.bits.extend(words)
        return off


def _as_contiguous_range(slot: nv.NsgymSlot, times) -> bool:
    """Strength reduction: an event set whose live part (inside the scheduler's own [start, end]) is
    one run a, a+1, .., b fires exactly like a ContinuousScheduler on [a, b] -- the kernels' fast
    class (one unsigned range compare) instead of a bitmap word fetched from the pool.  A single
    event time (the usual "step change at t*") is the common case."""
    live = [t for t in times if slot.start <= t <= slot.end]
    if not live or live[-1] - live[0] + 1 != len(live):
        return False
    slot.sched_op = nv.SCHED_CONTINUOUS
    slot.start, slot.end = live[0], live[-1]
    return True


def _lower_scheduler(sch, slot: nv.NsgymSlot, pools: _Pools, horizon: int, planes: dict, j: int):
    kind = type(sch).__name__
    slot.start, slot.end = _int_bounds(sch.start, sch.end)
    if kind == "ContinuousScheduler":
        slot.sched_op = nv.SCHED_CONTINUOUS
    elif kind == "PeriodicScheduler":
        if sch.period != int(sch.period) or sch.period <= 0:
            raise CompileError(f"PeriodicScheduler period must be a positive integer, got {sch.period}")
        slot.sched_op = nv.SCHED_PERIODIC
        slot.si[0] = int(sch.period)
    elif kind == "DiscreteScheduler":
        times = sorted({int(e) for e in sch.event_list if e == int(e) and e >= 0})
        if not _as_contiguous_range(slot, times):
            n_bits = (times[-1] + 1) if times else 0
            slot.sched_op = nv.SCHED_BITMAP
            slot.si[0] = pools.add_bitmap(times, n_bits)
            slot.si[1] = n_bits
    elif kind == "BurstScheduler":
        slot.sched_op = nv.SCHED_BURST
        slot.si[0] = int(sch.on_duration)
        slot.si[1] = int(sch.cycle)
        if slot.si[1] <= 0:
            raise CompileError("BurstScheduler cycle must be positive")
    elif kind == "WindowScheduler":
        flat = []
        for ws, we in sch.windows:
            lo, hi = _int_bounds(ws, we)
            flat += [lo, hi]
        lo1, hi1 = (max(flat[0], slot.start), min(flat[1], slot.end)) if len(flat) == 2 else (1, 0)
        if lo1 <= hi1:         # a single window = a Continuous scheduler on it (fast class)
            slot.sched_op = nv.SCHED_CONTINUOUS
            slot.start, slot.end = lo1, hi1
        else:
            slot.sched_op = nv.SCHED_WINDOW
            slot.si[0] = pools.add_i(flat)
            slot.si[1] = len(flat) // 2
    elif kind == "RandomScheduler":
        slot.sched_op = nv.SCHED_RANDOM
        slot.sf[0] = float(sch.probability)
    elif kind == "DecayingProbabilityScheduler":
        slot.sched_op = nv.SCHED_DECAY
        slot.sf[0] = float(sch.initial_probability)
        slot.sf[1] = float(sch.decay_rate)
    elif kind == "MemorylessScheduler":
        slot.sched_op = nv.SCHED_MEMORYLESS
        slot.sf[0] = float(sch.p)
        slot.istate_plane = planes.setdefault(j, len(planes))
        slot.istate_init = int(np.asarray(sch.transition_time).reshape(-1)[0])
    elif kind == "CustomScheduler":
        # arbitrary Python cannot run on the device: pre-evaluate, in increasing t, only where
        # the range gate would have called it (base.py:79-81)
        lo, hi = slot.start, min(slot.end, horizon)
        times = [t for t in range(lo, hi + 1) if sch.event_function(t)]
        if not _as_contiguous_range(slot, times):
            slot.sched_op = nv.SCHED_BITMAP
            slot.si[0] = pools.add_bitmap(times, horizon + 1)
            slot.si[1] = horizon + 1
    else:
        raise CompileError(f"scheduler {kind} cannot be compiled")


def _need_plane(slot, planes, j):
    """One integer state plane per parameter slot (list cursor or Memoryless next-fire time);
    ``planes`` maps slot position -> plane, allocated on first need."""
    if slot.istate_plane >= 0:
        # MemorylessScheduler (schedulers.py:92-116) driving a list update (single_param.py:202-223,
        # 388-408; distribution.py:100-130, 334-356): next-fire time (low 24 bits) and cursor (bits 24..30)
        # share the slot's word (IstPack, nsgym_device.cuh)
        if slot.ui[1] > 127:
            raise CompileError("a Memoryless scheduler driving a StepWise / Cyclic update is limited to lists of "
                               "127 entries (next-fire time and cursor share one integer state word)")
        if not 0 <= slot.istate_init < (1 << 24):
            raise CompileError("MemorylessScheduler: first transition time out of the packed range")
        slot.ui[3] = 1
        return
    slot.istate_plane = planes.setdefault(j, len(planes))
    slot.istate_init = 0


def _lower_scalar_update(fn, slot, pools, planes, j):
    kind = type(fn).__name__
    uf = slot.uf
    if kind == "IncrementUpdate":
        slot.upd_op, uf[0] = nv.UPD_ADD, float(fn.k)
    elif kind == "DecrementUpdate":
        slot.upd_op, uf[0] = nv.UPD_ADD, -float(fn.k)
    elif kind == "DeterministicTrend":
        slot.upd_op, uf[0] = nv.UPD_ADD_T, float(fn.slope)
    elif kind == "PolynomialTrend":
        slot.upd_op = nv.UPD_POLY
        slot.ui[0], slot.ui[1] = pools.add_f(fn.coeffs), len(fn.coeffs)
    elif kind == "GeometricProgression":
        slot.upd_op, uf[0] = nv.UPD_MUL, float(fn.r)
    elif kind == "ExponentialDecay":
        slot.upd_op, uf[0] = nv.UPD_MUL_EXP, float(fn.decay_rate)
    elif kind == "OscillatingUpdate":
        slot.upd_op, uf[0] = nv.UPD_ADD_SIN, float(fn.delta)
    elif kind == "SigmoidTransition":
        slot.upd_op = nv.UPD_SIGMOID
        uf[0], uf[1], uf[2], uf[3] = float(fn.a), float(fn.b - fn.a), float(fn.k), float(fn.t0)
    elif kind == "LinearInterpolation":
        slot.upd_op = nv.UPD_LERP
        uf[0], uf[1], uf[2] = float(fn.start_val), float(fn.end_val - fn.start_val), float(fn.T)
    elif kind == "StepWiseUpdate":
        slot.upd_op = nv.UPD_STEPWISE
        slot.ui[0], slot.ui[1] = pools.add_f(fn.param_list), len(fn.param_list)
        _need_plane(slot, planes, j)
    elif kind == "CyclicUpdate":
        slot.upd_op = nv.UPD_CYCLIC
        slot.ui[0], slot.ui[1] = pools.add_f(fn.value_list), len(fn.value_list)
        _need_plane(slot, planes, j)
    elif kind == "NoUpdate":
        slot.upd_op = nv.UPD_NOP
    elif kind == "RandomWalk":
        slot.upd_op = nv.UPD_RW
        uf[0], uf[1], uf[2], uf[3] = 0.0, float(fn.mu), float(fn.sigma), 0.0
    elif kind == "RandomWalkWithDrift":
        slot.upd_op = nv.UPD_RW
        uf[0], uf[1], uf[2], uf[3] = float(fn.alpha), float(fn.mu), float(fn.sigma), 0.0
    elif kind == "RandomWalkWithDriftAndTrend":
        slot.upd_op = nv.UPD_RW
        uf[0], uf[1], uf[2], uf[3] = float(fn.alpha), float(fn.mu), float(fn.sigma), float(fn.slope)
    elif kind == "OrnsteinUhlenbeck":
        slot.upd_op = nv.UPD_OU
        uf[0], uf[1], uf[2] = float(fn.theta), float(fn.mu), float(fn.sigma)
    elif kind == "BoundedRandomWalk":
        slot.upd_op = nv.UPD_BRW
        uf[0], uf[1], uf[2], uf[3] = float(fn.mu), float(fn.sigma), float(fn.lo), float(fn.hi)
    else:
        raise CompileError(f"update function {kind} cannot drive a scalar parameter")


def _lower_dist_update(fn, slot, pools, planes, j, n_dist):
    kind = type(fn).__name__
    uf = slot.uf

    def check(d):
        if len(d) != n_dist:
            raise CompileError(f"{kind}: distribution of length {len(d)}, this env needs {n_dist}")
        return [float(x) for x in d]

    if kind == "DistributionIncrementUpdate":
        slot.upd_op, uf[0] = nv.UPD_D_INC, float(fn.k)
    elif kind == "DistributionDecrementUpdate":
        slot.upd_op, uf[0] = nv.UPD_D_DEC, float(fn.k)
    elif kind == "DistributionNoUpdate":
        slot.upd_op = nv.UPD_D_NOP
    elif kind == "UniformDrift":
        slot.upd_op = nv.UPD_D_UNIFORM
        uf[0], uf[1] = float(1 - fn.rate), float(fn.rate * (1.0 / n_dist))
    elif kind == "TargetReversion":
        slot.upd_op = nv.UPD_D_TARGET
        uf[0] = float(fn.theta)
        for k, v in enumerate(check(fn.target)):
            uf[1 + k] = v
    elif kind == "DistributionLinearInterpolation":
        s, e = check(fn.start_dist), check(fn.end_dist)
        slot.upd_op, uf[0] = nv.UPD_D_LERP, float(fn.T)
        slot.ui[0] = pools.add_f(s + [b - a for a, b in zip(s, e)])
    elif kind == "DistributionStepWiseUpdate":
        slot.upd_op = nv.UPD_D_STEPWISE
        flat = [x for d in fn.update_values for x in check(d)]
        slot.ui[0], slot.ui[1] = pools.add_f(flat), len(fn.update_values)
        _need_plane(slot, planes, j)
    elif kind == "DistributionCyclicUpdate":
        slot.upd_op = nv.UPD_D_CYCLIC
        flat = [x for d in fn.dist_list for x in check(d)]
        slot.ui[0], slot.ui[1] = pools.add_f(flat), len(fn.dist_list)
        _need_plane(slot, planes, j)
    elif kind == "RandomCategorical":
        slot.upd_op = nv.UPD_D_RANDOM
    elif kind == "LCBoundedDistrubutionUpdate":
        # distribution.py:155-164: the inner rule is built as update_fn(scheduler), so only rules whose
        # constructor takes the scheduler alone can be inner rules at all
        inner = getattr(fn, "update_fn", None)
        name = "RandomCategorical" if inner is None else (
            inner.__name__ if isinstance(inner, type) else type(inner).__name__)
        if name == "RandomCategorical":
            slot.upd_op = nv.UPD_D_RANDOM
        elif name == "DistributionNoUpdate":
            slot.upd_op = nv.UPD_D_NOP
        else:
            raise CompileError(f"LCBoundedDistrubutionUpdate around {name} cannot be compiled")
        slot.ui[2] = 1
        uf[5] = float(fn.L)
    elif kind == "BudgetBoundedIncrement":
        raise CompileError(f"{kind} is broken upstream (its __call__ unpacks a 3-tuple into 2 names) and is "
                           "not lowered")
    else:
        raise CompileError(f"update function {kind} cannot drive a slip distribution")


def _check_aliasing(tunable_params):
    seen_fn, seen_s = {}, {}
    for key, fn in tunable_params.items():
        if type(fn).__name__ in _STATEFUL_UPDATES:
            if id(fn) in seen_fn:
                raise CompileError(f"stateful update function shared by {seen_fn[id(fn)]!r} and {key!r}: "
                                   "the reference interleaves calls on the shared state; bind separate instances")
            seen_fn[id(fn)] = key
        sch = fn.scheduler
        if type(sch).__name__ in _STATEFUL_SCHEDS:
            if id(sch) in seen_s:
                raise CompileError(f"stateful scheduler shared by {seen_s[id(sch)]!r} and {key!r}")
            seen_s[id(sch)] = key


def _grid_masks(desc):
    rows = [r.decode() if isinstance(r, bytes) else "".join(
        c.decode() if isinstance(c, bytes) else c for c in r) for r in desc]
    nrow, ncol = len(rows), len(rows[0])
    if nrow * ncol > 256:
        raise CompileError("gridworld maps are limited to 256 cells (one byte per cell in the kernels' map tables)")
    hole = goal = start = 0
    starts, letters = [], []
    for r, line in enumerate(rows):
        if len(line) != ncol:
            raise CompileError("gridworld map rows differ in length")
        for c, ch in enumerate(line):
            bit = 1 << (r * ncol + c)
            letters.append({"F": nv.CELL_FROZEN, "H": nv.CELL_HOLE, "G": nv.CELL_GOAL, "S": nv.CELL_START}.get(ch))
            if letters[-1] is None:
                raise CompileError(f"unknown map letter {ch!r}")
            if ch == "H":
                hole |= bit
            elif ch == "G":
                goal |= bit
            elif ch == "S":
                start |= bit
                starts.append(r * ncol + c)
    return nrow, ncol, hole, goal, start, starts, letters


def _lower_slot(slot, j, key, fn, kind, keys, order, pools, planes, horizon, n_dist):
    """One (parameter name, update function) pair -> one NsgymSlot."""
    is_grid = kind in GRID_KINDS
    slot.theta_index = order.index(key)
    slot.istate_plane = -1
    slot.partner_slot = -1
    slot.partner_index = 0
    _lower_scheduler(fn.scheduler, slot, pools, horizon, planes, j)
    if is_grid:
        _lower_dist_update(fn, slot, pools, planes, j, n_dist)
    else:
        _lower_scalar_update(fn, slot, pools, planes, j)
        slot.constraint = CONSTRAINTS[kind].get(key, nv.CONS_NONE)
        if slot.constraint in (nv.CONS_ACRO_LENGTH1, nv.CONS_ACRO_COM):
            partner = ACRO_PARTNER[key]
            slot.partner_index = order.index(partner)
            slot.partner_slot = keys.index(partner) if partner in keys else -1


def compile_program(env_id: str, tunable_params: dict, n_envs: int, *, precision: str = "fp32",
                    autoreset: str = "next_step", seed: int = 0, env_id_offset: int = 0,
                    persistent_params: bool = False, max_episode_steps=None, base_params=None,
                    initial_prob_dist=None, modified_rewards=None, terminal_cliff: bool = False,
                    map_name=None, desc=None, custom_horizon: int = 65536, _pools=None, _planes=None,
                    _finish=True, **_ignored) -> CompiledProgram:
    if env_id not in ENV_TABLE:
        raise CompileError(f"unknown environment id {env_id!r}; supported: {sorted(ENV_TABLE)}")
    kind, env_class, limit = ENV_TABLE[env_id]
    if max_episode_steps is not None:
        limit = int(max_episode_steps)
    order = THETA_ORDER[kind]
    # base.py:257-261
    assert set(tunable_params.keys()) <= set(TUNABLE_PARAMS.get(env_class, {}).keys()), (
        f"Tunable parameters {list(tunable_params.keys())} not all in default tunable parameters "
        f"{list(TUNABLE_PARAMS.get(env_class, {}).keys())} for environment {env_class}")
    _check_aliasing(tunable_params)
    if persistent_params and any(type(f).__name__ == "LCBoundedDistrubutionUpdate" for f in tunable_params.values()):
        raise CompileError("LCBoundedDistrubutionUpdate with persistent_params: the bound L |t - prev_time| "
                           "carries prev_time across resets in the reference; not lowered")
    is_grid = kind in GRID_KINDS
    spec = nv.NsgymSpec()
    spec.abi_version = nv.ABI_VERSION
    spec.env_kind = kind
    spec.precision = {"fp32": nv.F32, "fp64": nv.F64}[precision] if not is_grid else nv.F64
    spec.autoreset = {"none": nv.AUTORESET_NONE, "next_step": nv.AUTORESET_NEXT_STEP}[autoreset]
    if not 0 < int(n_envs) <= (1 << 28):
        raise CompileError(f"n_envs must be in 1 .. 2^28 per handle (got {n_envs}): shard larger batches over handles")
    spec.n_envs = int(n_envs)
    spec.env_id_offset = int(env_id_offset)
    spec.seed = int(seed) & (2**64 - 1)
    spec.max_episode_steps = int(limit or 0)
    spec.persistent_params = int(bool(persistent_params))
    spec.n_slots = len(tunable_params)
    n_dist = 0
    if is_grid:
        n_dist = 4 if kind == nv.ENV_CLIFFWALKING else 3
    spec.n_dist = n_dist
    # A CustomScheduler is arbitrary Python: it is evaluated ONCE, ahead of time, for t = 0 .. horizon
    # (same for every env and episode -- a stateful or random event_function cannot be lowered
    # faithfully) into a fire bitmap.  The horizon covers every reachable t: an episode is at most
    # `limit` steps and a planning copy taken at t <= limit steps on for at most another limit of
    # its own (`_PLANNING_LIMIT`, <= 1000); without a TimeLimit (CliffWalking-v1, autoreset "none"
    # stepping past the end) `custom_horizon` (default 65536) is the documented bound -- past it a
    # CustomScheduler never fires.
    if spec.max_episode_steps > 0 and autoreset == "next_step":
        horizon = max(2 * spec.max_episode_steps, spec.max_episode_steps + 1000) + 2
    else:
        horizon = max(int(custom_horizon), 2 * spec.max_episode_steps + 1002)

    pools = _Pools() if _pools is None else _pools
    planes = {} if _planes is None else _planes
    keys = list(tunable_params.keys())
    for j, (key, fn) in enumerate(tunable_params.items()):
        _lower_slot(spec.slots[j], j, key, fn, kind, keys, order, pools, planes, horizon, n_dist)

    # ---- initial values ----
    if not is_grid:
        defaults = dict(TUNABLE_PARAMS[env_class])
        if base_params:
            unknown = set(base_params) - set(defaults)
            if unknown:
                raise CompileError(f"unknown base parameters {sorted(unknown)}")
            defaults.update(base_params)
        for i, name in enumerate(order):
            spec.theta_init[i][0] = float(defaults[name])
    else:
        default_dist = [1, 0, 0, 0][:n_dist]
        ipd = default_dist if initial_prob_dist is None else initial_prob_dist
        if kind == nv.ENV_BRIDGE and isinstance(ipd, tuple) and len(ipd) == 2:   # toy_text.py:573-579
            left, right, uni = list(ipd[0]), list(ipd[1]), list(ipd[0])
        else:
            left = right = uni = list(ipd)
        if kind == nv.ENV_FROZENLAKE:                                            # toy_text.py:329-334
            assert sum(uni) == 1 or math.isclose(sum(uni), 1), "The sum of transition probabilities must be 1."
            assert len(uni) == 3, ("The length of the transition probability distribution must be 3. "
                                   "Each action can have at most 3 possible outcomes.")
        for idx, dist in enumerate((uni, left, right)):
            if len(dist) != n_dist:
                raise CompileError(f"initial_prob_dist must have {n_dist} entries")
            for k in range(n_dist):
                spec.theta_init[idx][k] = float(dist[k])
        # ---- map ----
        if kind == nv.ENV_FROZENLAKE:
            if desc is None:
                desc = FROZEN_LAKE_MAPS[map_name or ("8x8" if env_id == "FrozenLake8x8-v1" else "4x4")]
            nrow, ncol, hole, goal, start, starts, letters = _grid_masks(desc)
            if len(starts) < 1:
                raise CompileError("the FrozenLake map has no start cell 'S'")
            rw = {"F": 0.0, "H": 0.0, "G": 1.0, "S": 0.0}
            if modified_rewards:
                rw = {k: float(modified_rewards[k]) for k in "FHGS"}
            # several 'S' cells: reset draws the start cell, categorical_sample(initial_state_distrib) over the
            # start cells in row-major order (gymnasium FrozenLakeEnv.reset; toy_text.py:314-319 accepts any desc)
            start_cell = starts[0] if len(starts) == 1 else -1
        elif kind == nv.ENV_CLIFFWALKING:
            nrow, ncol, letters = 4, 12, None
            hole = sum(1 << (3 * 12 + c) for c in range(1, 11))
            goal, start, start_cell = 1 << 47, 0, 36
            rw = {"H": -100.0, "G": 0.0, "F": -1.0, "S": -1.0}
            if modified_rewards:
                rw.update({k: float(v) for k, v in modified_rewards.items()})
        else:
            nrow, ncol, hole, goal, start, starts, letters = _grid_masks(BRIDGE_MAP)
            start_cell = 2 * ncol + 4                                            # envs/Bridge.py:110
            rw = {"F": 0.0, "H": -1.0, "G": 1.0, "S": 0.0}                       # envs/Bridge.py:159-174
            spec.split_mode = int(("P_left" in tunable_params) or ("P_right" in tunable_params))
        spec.nrow, spec.ncol = nrow, ncol
        if nrow * ncol <= 64:
            spec.hole_mask, spec.goal_mask, spec.start_mask = hole, goal, start
        else:       # the masks hold 64 bits: larger maps (toy_text.py:314-319 accepts any desc) travel as letters
            arr = (C.c_uint8 * len(letters))(*letters)
            spec.cell_class, spec.n_cell_class = C.cast(arr, C.POINTER(C.c_uint8)), len(letters)
            _map_keepalive = arr
        spec.start_cell = start_cell
        spec.reward_f, spec.reward_h, spec.reward_g, spec.reward_s = rw["F"], rw["H"], rw["G"], rw["S"]
        spec.terminal_cliff = int(bool(terminal_cliff))

    prog = CompiledProgram(spec=spec, env_id=env_id, env_kind=kind, env_class=env_class, keys=keys,
                           precision=int(spec.precision), n_dist=n_dist)
    prog.horizon = horizon
    if is_grid and spec.n_cell_class:
        prog._map_keepalive = _map_keepalive      # the spec points into it until nsgym_create has copied it
    if _finish:
        _attach_pools(prog, pools)
    return prog


def _attach_pools(prog: CompiledProgram, pools: _Pools):
    spec = prog.spec
    keep = prog._keepalive
    if pools.f:
        arr = (C.c_double * len(pools.f))(*pools.f)
        spec.pool_f, spec.n_pool_f = C.cast(arr, C.POINTER(C.c_double)), len(pools.f)
        keep.append(arr)
    if pools.i:
        arr = (C.c_int32 * len(pools.i))(*pools.i)
        spec.pool_i, spec.n_pool_i = C.cast(arr, C.POINTER(C.c_int32)), len(pools.i)
        keep.append(arr)
    if pools.bits:
        arr = (C.c_uint32 * len(pools.bits))(*pools.bits)
        spec.bitmap, spec.n_bitmap_words = C.cast(arr, C.POINTER(C.c_uint32)), len(pools.bits)
        keep.append(arr)


def compile_rows(env_id: str, params_per_env, **kwargs):
    """Heterogeneous batch (BASELINE config C4): ``params_per_env[e]`` is the ``tunable_params``
    dict env e's wrapper would be given.  Every dict must bind the same parameter names in the
    same order; schedulers, update functions and their coefficients are free per env.

    Returns ``(CompiledProgram, rows)`` with ``rows`` a C-contiguous numpy structured array
    ``[n_envs, n_slots]`` of ``NsgymSlot`` records -- the argument of ``nsgym_create_rows``.
    (Large synthetic batches can fill such an array directly with numpy: ``rows_dtype()``.)"""
    params_per_env = list(params_per_env)
    n = len(params_per_env)
    if n == 0:
        raise CompileError("empty batch")
    keys = list(params_per_env[0].keys())
    if not keys:
        raise CompileError("a heterogeneous batch needs at least one bound parameter")
    pools, planes = _Pools(), {}
    prog = compile_program(env_id, params_per_env[0], n, _pools=pools, _planes=planes, _finish=False, **kwargs)
    kind, order = prog.env_kind, THETA_ORDER[prog.env_kind]
    rows = np.zeros((n, len(keys)), dtype=rows_dtype())
    tmp = nv.NsgymSlot()
    size = C.sizeof(nv.NsgymSlot)
    flat = rows.reshape(-1).view(np.uint8).reshape(-1, size)
    for e, tp in enumerate(params_per_env):
        if list(tp.keys()) != keys:
            raise CompileError(f"env {e} binds {list(tp.keys())}, env 0 binds {keys}: the key set is shared")
        _check_aliasing(tp)
        for j, (key, fn) in enumerate(tp.items()):
            C.memset(C.byref(tmp), 0, size)
            _lower_slot(tmp, j, key, fn, kind, keys, order, pools, planes, prog.horizon, prog.n_dist)
            flat[e * len(keys) + j] = np.frombuffer(bytes(tmp), dtype=np.uint8)
    for j in range(len(keys)):                     # a cursor plane exists if ANY env needs one
        prog.spec.slots[j].istate_plane = planes.get(j, -1)
    _attach_pools(prog, pools)
    prog.pool_lists = (list(pools.f), list(pools.i), list(pools.bits))
    return prog, rows


def row_signature(rows):
    """Opcode signature of every env of a rows array: envs with equal signatures take the same
    branches in the heterogeneous step kernel (only their coefficients differ)."""
    sig = np.zeros(len(rows), dtype=np.int64)
    for j in range(rows.shape[1]):
        sig = sig * 4096 + rows["sched_op"][:, j].astype(np.int64) * 64 + rows["upd_op"][:, j].astype(np.int64)
    return sig


def rows_dtype():
    """numpy structured dtype with the memory layout of ``NsgymSlot``."""
    return np.dtype(nv.NsgymSlot)
