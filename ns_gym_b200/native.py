"""ctypes binding of ``include/nsgym_b200.h``.

The library is built in-tree by ``ns_gym_b200.build`` (nvcc, sm_100a) and loaded from
``ns_gym_b200/_lib/libnsgym_b200.so``.  Loading fails loudly when the library is missing;
there is no fallback implementation.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

MAX_SLOTS = 8
MAX_THETA = 8
MAX_DIST = 4
ABI_VERSION = 2

ENV_CARTPOLE, ENV_ACROBOT, ENV_MOUNTAINCAR, ENV_MOUNTAINCAR_CONT, ENV_PENDULUM = 0, 1, 2, 3, 4
ENV_FROZENLAKE, ENV_CLIFFWALKING, ENV_BRIDGE = 5, 6, 7
F32, F64 = 0, 1
AUTORESET_NONE, AUTORESET_NEXT_STEP = 0, 1
FLAG_TERMINATED, FLAG_TRUNCATED, FLAG_RESET, FLAG_REJECTED, FLAG_BAD_DIST = 1, 2, 4, 8, 128

SCHED_CONTINUOUS, SCHED_PERIODIC, SCHED_BITMAP, SCHED_BURST = 0, 1, 2, 3
SCHED_WINDOW, SCHED_RANDOM, SCHED_DECAY, SCHED_MEMORYLESS = 4, 5, 6, 7

UPD_NOP, UPD_ADD, UPD_ADD_T, UPD_POLY, UPD_MUL, UPD_MUL_EXP, UPD_ADD_SIN = 0, 1, 2, 3, 4, 5, 6
UPD_SIGMOID, UPD_LERP, UPD_STEPWISE, UPD_CYCLIC, UPD_RW, UPD_OU, UPD_BRW = 7, 8, 9, 10, 11, 12, 13
UPD_D_NOP, UPD_D_INC, UPD_D_DEC, UPD_D_UNIFORM = 32, 33, 34, 35
UPD_D_TARGET, UPD_D_LERP, UPD_D_STEPWISE, UPD_D_CYCLIC, UPD_D_RANDOM = 36, 37, 38, 39, 40

CONS_NONE, CONS_REJECT_LE0, CONS_REJECT_LT0, CONS_ACRO_LENGTH1, CONS_ACRO_COM = 0, 1, 2, 3, 4

OPT_GENERAL_KERNELS = 1
OPT_SPECIALIZE = 2
CELL_FROZEN, CELL_HOLE, CELL_GOAL, CELL_START = 0, 1, 2, 3
STAT_KEYS = ("steps", "episodes", "return_sum", "length_sum", "terminated", "truncated", "rejected_updates",
             "bad_dist")
KERNEL_LEAN_FAST, KERNEL_LEAN_MEDIUM, KERNEL_GENERAL, KERNEL_ROWS_LEAN, KERNEL_ROWS_GENERAL = range(5)
DRAW_NORMAL, DRAW_SCHED_UNIFORM, DRAW_RESET_UNIFORMS, DRAW_GEOMETRIC, DRAW_DYN_UNIFORM, DRAW_DIRICHLET, \
    DRAW_BOX_MULLER_SWEEP = range(7)

T_ENDED = 0x80000000
T_TERMINATED_ONCE = 0x40000000
T_TABLE_FRESH = 0x20000000
T_TIME_MASK = 0x0FFFFFFF
INT32_MAX = 2**31 - 1


class NsgymSlot(C.Structure):
    _fields_ = [
        ("sched_op", C.c_int32), ("upd_op", C.c_int32), ("theta_index", C.c_int32),
        ("constraint", C.c_int32), ("start", C.c_int32), ("end", C.c_int32),
        ("si", C.c_int32 * 4), ("ui", C.c_int32 * 4),
        ("partner_slot", C.c_int32), ("partner_index", C.c_int32),
        ("istate_plane", C.c_int32), ("istate_init", C.c_int32),
        ("sf", C.c_double * 2), ("uf", C.c_double * 6),
    ]


class NsgymSpec(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("env_kind", C.c_int32), ("precision", C.c_int32),
        ("autoreset", C.c_int32), ("n_envs", C.c_int64), ("env_id_offset", C.c_int64),
        ("seed", C.c_uint64), ("max_episode_steps", C.c_int32), ("persistent_params", C.c_int32),
        ("n_slots", C.c_int32), ("n_dist", C.c_int32),
        ("slots", NsgymSlot * MAX_SLOTS),
        ("theta_init", (C.c_double * MAX_DIST) * MAX_THETA),
        ("pool_f", C.POINTER(C.c_double)), ("n_pool_f", C.c_int32),
        ("pool_i", C.POINTER(C.c_int32)), ("n_pool_i", C.c_int32),
        ("bitmap", C.POINTER(C.c_uint32)), ("n_bitmap_words", C.c_int32),
        ("nrow", C.c_int32), ("ncol", C.c_int32),
        ("hole_mask", C.c_uint64), ("goal_mask", C.c_uint64), ("start_mask", C.c_uint64),
        ("start_cell", C.c_int32), ("split_mode", C.c_int32),
        ("reward_f", C.c_float), ("reward_h", C.c_float), ("reward_g", C.c_float),
        ("reward_s", C.c_float), ("terminal_cliff", C.c_int32), ("n_cell_class", C.c_int32),
        ("cell_class", C.POINTER(C.c_uint8)),
    ]


class NsgymLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in
                ("state", "theta", "t", "istate", "action", "reward", "flags", "change", "delta", "obs")] + [
        ("state_words", C.c_int32), ("obs_words", C.c_int32), ("n_istate", C.c_int32),
        ("theta_planes", C.c_int32), ("bytes_per_step", C.c_double), ("row_bytes_per_env", C.c_double)]


class NsgymBuffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("d_state", "d_theta", "d_t", "d_istate", "d_action", "d_reward", "d_flags",
                 "d_change", "d_delta", "d_obs")]


class NsgymHostOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("h_reward", "h_flags", "h_change", "h_delta", "h_state", "h_obs")]


class NsgymSnapshotInfo(C.Structure):
    _fields_ = [("step_index", C.c_uint64), ("plan_elapsed", C.c_int32), ("_reserved", C.c_int32)]


EXPORTS = [
    "nsgym_abi_version", "nsgym_sizeof", "nsgym_last_error", "nsgym_create", "nsgym_create_rows", "nsgym_destroy",
    "nsgym_layout", "nsgym_bind", "nsgym_reset", "nsgym_step", "nsgym_step_many", "nsgym_unpack", "nsgym_episode_stats", "nsgym_step_host", "nsgym_alloc_host", "nsgym_free_host",
    "nsgym_rollout", "nsgym_rollout_linear", "nsgym_fanout", "nsgym_snapshot_bytes", "nsgym_snapshot", "nsgym_restore", "nsgym_transition_table", "nsgym_set_option", "nsgym_eval_update", "nsgym_eval_w1", "nsgym_eval_draws", "nsgym_set_seed", "nsgym_step_index", "nsgym_set_step_index",
    "nsgym_launch_count", "nsgym_last_kernel_class", "nsgym_last_kernel_specialized", "nsgym_jit_check", "nsgym_jit_stats",
]

_lib = None


class NsgymError(RuntimeError):
    pass


def lib_path() -> str:
    """In-tree library; ``NSGYM_B200_LIB`` points at an alternative build (kernel experiments)."""
    return os.environ.get("NSGYM_B200_LIB") or _build.LIB_PATH


def load(build_if_missing: bool = False):
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        if build_if_missing:
            _build.build_library()
        else:
            raise NsgymError(
                f"{path} not found: build the CUDA library first (python -m ns_gym_b200.build or "
                "__graft_entry__.build()); ns_gym_b200 has no CPU fallback")
    if path == _build.LIB_PATH and not _build.is_current():
        # a library older than its sources is never loaded: rebuild it in place (nvcc, ~2 min) unless
        # told not to; without a toolchain this still fails loudly -- there is no CPU fallback
        if os.environ.get("NSGYM_B200_NO_AUTOBUILD"):
            raise NsgymError(f"{path} is older than the sources under ns_gym_b200/csrc: rebuild it "
                             "(python -m ns_gym_b200.build or __graft_entry__.build())")
        import sys
        sys.stderr.write(f"ns_gym_b200: {path} is older than its sources, rebuilding with nvcc ...\n")
        _build.build_library()
    lib = C.CDLL(path)
    lib.nsgym_abi_version.restype = C.c_int
    lib.nsgym_sizeof.restype = C.c_size_t
    lib.nsgym_sizeof.argtypes = [C.c_int]
    lib.nsgym_last_error.restype = C.c_char_p
    lib.nsgym_create.argtypes = [C.POINTER(NsgymSpec), C.POINTER(C.c_void_p)]
    lib.nsgym_create_rows.argtypes = [C.POINTER(NsgymSpec), C.c_void_p, C.POINTER(C.c_void_p)]
    lib.nsgym_destroy.argtypes = [C.c_void_p]
    lib.nsgym_destroy.restype = None
    lib.nsgym_layout.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(NsgymLayout)]
    lib.nsgym_bind.argtypes = [C.c_void_p, C.POINTER(NsgymBuffers)]
    lib.nsgym_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.nsgym_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.nsgym_step_many.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_void_p]
    lib.nsgym_unpack.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.nsgym_episode_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.nsgym_step_host.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(NsgymHostOut), C.c_int, C.c_void_p]
    lib.nsgym_alloc_host.argtypes = [C.c_size_t, C.c_int]
    lib.nsgym_alloc_host.restype = C.c_void_p
    lib.nsgym_free_host.argtypes = [C.c_void_p]
    lib.nsgym_free_host.restype = None
    lib.nsgym_rollout.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                  C.c_int, C.c_void_p]
    lib.nsgym_fanout.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.nsgym_snapshot_bytes.restype = C.c_size_t
    lib.nsgym_rollout_linear.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                         C.c_int, C.c_void_p]
    lib.nsgym_snapshot_bytes.argtypes = [C.c_void_p]
    lib.nsgym_snapshot.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(NsgymSnapshotInfo), C.c_void_p]
    lib.nsgym_restore.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(NsgymSnapshotInfo), C.c_void_p]
    lib.nsgym_transition_table.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p]
    lib.nsgym_set_option.argtypes = [C.c_void_p, C.c_int, C.c_int64]
    lib.nsgym_eval_w1.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.nsgym_eval_w1.restype = C.c_int
    lib.nsgym_eval_update.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                      C.c_void_p]
    lib.nsgym_eval_draws.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_uint64, C.c_void_p,
                                     C.c_int64, C.c_void_p]
    lib.nsgym_set_seed.argtypes = [C.c_void_p, C.c_uint64]
    lib.nsgym_set_seed.restype = None
    lib.nsgym_step_index.restype = C.c_uint64
    lib.nsgym_step_index.argtypes = [C.c_void_p]
    lib.nsgym_set_step_index.argtypes = [C.c_void_p, C.c_uint64]
    lib.nsgym_set_step_index.restype = None
    lib.nsgym_launch_count.restype = C.c_int64
    lib.nsgym_launch_count.argtypes = [C.c_void_p]
    lib.nsgym_last_kernel_class.argtypes = [C.c_void_p]
    lib.nsgym_last_kernel_specialized.argtypes = [C.c_void_p]
    lib.nsgym_jit_check.argtypes = [C.POINTER(NsgymSpec), C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    lib.nsgym_jit_stats.argtypes = [C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_char_p, C.c_size_t]
    if lib.nsgym_abi_version() != ABI_VERSION:
        raise NsgymError("libnsgym_b200.so ABI version differs from ns_gym_b200/native.py")
    for which, st in enumerate((NsgymSlot, NsgymSpec, NsgymLayout, NsgymBuffers, NsgymHostOut, NsgymSnapshotInfo)):
        if lib.nsgym_sizeof(which) != C.sizeof(st):
            raise NsgymError(f"struct layout mismatch for {st.__name__}: C {lib.nsgym_sizeof(which)} "
                             f"vs ctypes {C.sizeof(st)}")
    _lib = lib
    return lib


def check(rc: int, what: str = "nsgym"):
    if rc != 0:
        msg = load().nsgym_last_error()
        raise NsgymError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")
