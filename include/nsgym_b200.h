/* nsgym_b200.h -- C ABI of the B200-native batched simulator for ns_gym's
 * non-stationary env-step path.
 *
 * The reference has no FFI on this path: the boundary is a pure-Python object protocol,
 *   NSClassicControlWrapper.step   ns_gym/wrappers/classic_control.py:60-100
 *   NSFrozenLakeWrapper.step       ns_gym/wrappers/toy_text.py:342-380
 *   NSCliffWalkingWrapper.step     ns_gym/wrappers/toy_text.py:162-193
 *   NSBridgeWrapper.step           ns_gym/wrappers/toy_text.py:633-651
 *   NSWrapper.step / reset         ns_gym/base.py:296-410
 *   UpdateFn.__call__              ns_gym/base.py:124-149
 *   Scheduler.__call__             ns_gym/base.py:67-81
 * This header is the boundary a maintainer would bind (ctypes stub: INTEGRATION.md):
 * the wrapper's `tunable_params` dict is lowered on the host to one NsgymSlot per bound
 * parameter (scheduler opcode + update opcode + coefficients), and one kernel launch
 * advances every env of the batch by one step.
 *
 * Conventions
 *  - every `d_` pointer is a DEVICE pointer owned by the caller (a torch tensor); the
 *    library never frees caller memory.  `h_` pointers are host memory (pinned for speed).
 *  - every compute call takes a cudaStream_t (passed as void*) and is asynchronous unless
 *    stated; a handle is not thread-safe.
 *  - return value: 0 = OK, negative = error; text via nsgym_last_error().
 *  - SoA layout, N = number of envs of this handle, w = 4 (NSGYM_F32) or 8 (NSGYM_F64):
 *      state   classic control: one packed vector per env, real[N][S] (S = 4 or 2),
 *              16/32-byte aligned so a thread moves it with 128-bit accesses;
 *              gridworlds: int32[N] (cell index)
 *      theta   real[P][N]   one plane per BOUND scalar parameter (classic control);
 *              double[P*D][N] for gridworlds (D = 3 or 4 probabilities per parameter,
 *              always fp64 so the cumulative-sum comparisons are bit-exact)
 *      t       int32[N]; bits 0..27 episode time, bit 31 = episode ended at the last call
 *              (consumed by next-step autoreset), bit 30 = CartPole "already terminated
 *              once" (reward 0 on further steps), bit 29 = gridworld "P table rebuilt
 *              since the last reset" (stale-table rule)
 *      istate  int32[I][N]  list cursors / Memoryless next-fire time, I = nsgym layout
 *      action  int32[N] (discrete) or real[N] (Pendulum, MountainCarContinuous)
 *      reward  float[N];  flags uint8[N];  change uint8[N] (bit j = slot j changed);
 *      delta   real[P][N] (double for gridworlds) or NULL;  obs float[N][O] or NULL
 */
#ifndef NSGYM_B200_H
#define NSGYM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NSGYM_ABI_VERSION 2
#define NSGYM_MAX_SLOTS 8
#define NSGYM_MAX_THETA 8
#define NSGYM_MAX_DIST 4

/* ---- env kinds (reference class: ns_gym/base.py:611-635, envs/Bridge.py) ---- */
enum {
  NSGYM_ENV_CARTPOLE = 0,        /* CartPoleEnv   theta: gravity masscart masspole force_mag tau length */
  NSGYM_ENV_ACROBOT = 1,         /* AcrobotEnv    theta: dt L1 L2 M1 M2 COM1 COM2 MOI */
  NSGYM_ENV_MOUNTAINCAR = 2,     /* MountainCarEnv theta: gravity force */
  NSGYM_ENV_MOUNTAINCAR_CONT = 3,/* Continuous_MountainCarEnv theta: power */
  NSGYM_ENV_PENDULUM = 4,        /* PendulumEnv   theta: m l dt g */
  NSGYM_ENV_FROZENLAKE = 5,      /* FrozenLakeEnv theta: P */
  NSGYM_ENV_CLIFFWALKING = 6,    /* CliffWalkingEnv theta: P (4 outcomes) */
  NSGYM_ENV_BRIDGE = 7,          /* Bridge        theta: P | P_left, P_right */
  NSGYM_ENV_COUNT = 8
};

enum { NSGYM_F32 = 0, NSGYM_F64 = 1 };

/* autoreset: NONE = caller resets (single-env semantics, stepping past the end allowed);
 * NEXT_STEP = gymnasium 1.x vector default: an env that ended at call k is reset by call
 * k+1 (action ignored, reward 0, flags = NSGYM_FLAG_RESET only). */
enum { NSGYM_AUTORESET_NONE = 0, NSGYM_AUTORESET_NEXT_STEP = 1 };

enum {
  NSGYM_FLAG_TERMINATED = 1,
  NSGYM_FLAG_TRUNCATED = 2,
  NSGYM_FLAG_RESET = 4,          /* this call performed the autoreset of the env */
  NSGYM_FLAG_REJECTED = 8,       /* classic control: a fired update was rejected by the constraint checker and
                                    theta kept its old value (ConstraintViolationWarning, classic_control.py:87-92) */
  NSGYM_FLAG_BAD_DIST = 128      /* gridworld: negative weight / sum != 1 (reference raises) */
};

/* ---- scheduler opcodes (ns_gym/schedulers.py) ---- */
enum {
  NSGYM_SCHED_CONTINUOUS = 0,    /* :46-53   always */
  NSGYM_SCHED_PERIODIC = 1,      /* :77-89   t % si[0] == 0 */
  NSGYM_SCHED_BITMAP = 2,        /* :56-74 Discrete, :31-43 Custom pre-evaluated; bit t of bitmap[si[0]..], si[1] bits */
  NSGYM_SCHED_BURST = 3,         /* :119-140 (t % si[1]) < si[0] */
  NSGYM_SCHED_WINDOW = 4,        /* :180-198 any pool_i[si[0]+2k] <= t <= pool_i[si[0]+2k+1], k < si[1] */
  NSGYM_SCHED_RANDOM = 5,        /* :9-28    u < sf[0] */
  NSGYM_SCHED_DECAY = 6,         /* :143-177 u < sf[0] * exp(-sf[1] t) */
  NSGYM_SCHED_MEMORYLESS = 7,    /* :92-116  t == istate; istate = t + Geom(sf[0]); first time si[0] */
  NSGYM_SCHED_COUNT = 8
};

/* ---- update opcodes (update_functions/single_param.py, distribution.py) ---- */
enum {
  NSGYM_UPD_NOP = 0,             /* NoUpdate :226-240 */
  NSGYM_UPD_ADD = 1,             /* Increment :154-175 (uf0 = k), Decrement :178-199 (uf0 = -k) */
  NSGYM_UPD_ADD_T = 2,           /* DeterministicTrend :20-40   y + uf0 * t */
  NSGYM_UPD_POLY = 3,            /* PolynomialTrend :451-473    coeffs pool_f[ui0..ui0+ui1) */
  NSGYM_UPD_MUL = 4,             /* GeometricProgression :290-307 */
  NSGYM_UPD_MUL_EXP = 5,         /* ExponentialDecay :266-287   y * exp(-uf0 t) */
  NSGYM_UPD_ADD_SIN = 6,         /* OscillatingUpdate :243-264  y + uf0 sin t */
  NSGYM_UPD_SIGMOID = 7,         /* SigmoidTransition :349-385  uf = a, b-a, k, t0 */
  NSGYM_UPD_LERP = 8,            /* LinearInterpolation :476-508 uf = s, e-s, T */
  NSGYM_UPD_STEPWISE = 9,        /* StepWiseUpdate :202-223     values pool_f[ui0..ui0+ui1), cursor istate */
  NSGYM_UPD_CYCLIC = 10,         /* CyclicUpdate :388-408 */
  NSGYM_UPD_RW = 11,             /* RandomWalk :84-113, WithDrift :116-151, WithDriftAndTrend :43-81; uf = alpha, mu, sigma, slope */
  NSGYM_UPD_OU = 12,             /* OrnsteinUhlenbeck :310-346  uf = theta, mu, sigma */
  NSGYM_UPD_BRW = 13,            /* BoundedRandomWalk :411-448  uf = mu, sigma, lo, hi */
  /* distribution opcodes (distribution.py) */
  NSGYM_UPD_D_NOP = 32,          /* :217-231 */
  NSGYM_UPD_D_INC = 33,          /* :41-67   uf0 = k */
  NSGYM_UPD_D_DEC = 34,          /* :70-97   uf0 = k */
  NSGYM_UPD_D_UNIFORM = 35,      /* :234-261 uf0 = 1-rate, uf1 = rate * (1/n) */
  NSGYM_UPD_D_TARGET = 36,       /* :264-293 uf0 = theta, uf1..4 = target */
  NSGYM_UPD_D_LERP = 37,         /* :296-331 uf0 = T, pool_f[ui0..] = start[D], (end-start)[D] */
  NSGYM_UPD_D_STEPWISE = 38,     /* :100-130 ui1 distributions of D doubles at pool_f[ui0..] */
  NSGYM_UPD_D_CYCLIC = 39,       /* :334-356 */
  NSGYM_UPD_D_RANDOM = 40        /* RandomCategorical :11-38: Dirichlet(1,..,1) = standard exponentials
                                    from the env's Philox stream scaled by the reciprocal of their sum */
  /* Lipschitz bound (LCBoundedDistrubutionUpdate :133-183) on any distribution opcode: ui[2] = 1,
   * uf[5] = L; the candidate must satisfy W1(p, p') <= L |t - prev_time| (= L: the rule is called
   * every step).  D_RANDOM is redrawn until it does (<= 1e5 tries, as the reference); a
   * deterministic rule that fails sets NSGYM_FLAG_BAD_DIST (the reference raises). */
};

/* ---- constraint rules (wrappers/classic_control.py:193-422) ---- */
enum {
  NSGYM_CONS_NONE = 0,
  NSGYM_CONS_REJECT_LE0 = 1,     /* reject new <= 0 */
  NSGYM_CONS_REJECT_LT0 = 2,     /* reject new <  0 */
  NSGYM_CONS_ACRO_LENGTH1 = 3,   /* :241-265  <=0 | new partner COM > new | new < current COM */
  NSGYM_CONS_ACRO_COM = 4        /* :307-357  <=0 | new partner LEN < new | new > current LEN */
};

typedef struct {
  int32_t sched_op;
  int32_t upd_op;
  int32_t theta_index;           /* which physical parameter (order of the kind's list above) */
  int32_t constraint;
  int32_t start;                 /* inclusive; INT32_MAX = never in range */
  int32_t end;                   /* inclusive; INT32_MAX = +inf */
  int32_t si[4];
  int32_t ui[4];
  int32_t partner_slot;          /* Acrobot cross-checks: slot of the partner parameter or -1 */
  int32_t partner_index;         /* theta index of the partner parameter */
  int32_t istate_plane;          /* plane of this slot in istate[][] or -1 */
  int32_t istate_init;           /* cursor 0 / first Memoryless time */
  double sf[2];
  double uf[6];
} NsgymSlot;

typedef struct {
  int32_t abi_version;           /* NSGYM_ABI_VERSION */
  int32_t env_kind;
  int32_t precision;             /* NSGYM_F32 | NSGYM_F64 (classic control only) */
  int32_t autoreset;
  int64_t n_envs;                /* envs owned by THIS handle (one shard) */
  int64_t env_id_offset;         /* global id of local env 0: Philox keys use global ids */
  uint64_t seed;
  int32_t max_episode_steps;     /* TimeLimit; <= 0: none */
  int32_t persistent_params;     /* base.py:381-395 */
  int32_t n_slots;
  int32_t n_dist;                /* gridworlds: probabilities per parameter (3 | 4), else 0 */
  NsgymSlot slots[NSGYM_MAX_SLOTS];
  double theta_init[NSGYM_MAX_THETA][NSGYM_MAX_DIST]; /* per theta_index; scalars use [i][0] */
  /* pools referenced by the slots (host pointers, copied at create) */
  const double* pool_f; int32_t n_pool_f;
  const int32_t* pool_i; int32_t n_pool_i;
  const uint32_t* bitmap; int32_t n_bitmap_words;
  /* gridworld description (toy_text.py:426-469, :86-138; envs/Bridge.py:12,113-174) */
  int32_t nrow, ncol;
  uint64_t hole_mask, goal_mask, start_mask; /* bit = cell index; maps of <= 64 cells (else cell_class) */
  int32_t start_cell;            /* -1 (FrozenLake): the map has several start cells, a reset samples one uniformly
                                    (categorical_sample over them in row-major order, with this step's gridworld uniform) */
  int32_t split_mode;            /* Bridge: P_left / P_right by column of the current cell */
  float reward_f, reward_h, reward_g, reward_s; /* by destination cell letter */
  int32_t terminal_cliff;        /* CliffWalking */
  int32_t n_cell_class;          /* 0, or nrow * ncol when cell_class is given */
  const uint8_t* cell_class;     /* host, copied at create: letter of every cell (NSGYM_CELL_*), row-major; replaces
                                    the three masks -- required for maps of more than 64 cells (toy_text.py:314-319
                                    accepts any desc); nrow * ncol <= 256 */
} NsgymSpec;
enum { NSGYM_CELL_FROZEN = 0, NSGYM_CELL_HOLE = 1, NSGYM_CELL_GOAL = 2, NSGYM_CELL_START = 3 };

typedef struct {                 /* bytes the caller must allocate for each bound buffer */
  size_t state, theta, t, istate, action, reward, flags, change, delta, obs;
  int32_t state_words, obs_words, n_istate, theta_planes;
  double bytes_per_step;         /* algorithmic bytes per env-step, delta & obs as bound;
                                    heterogeneous handles include row_bytes_per_env */
  double row_bytes_per_env;      /* per-env opcode/coefficient row words read each step (0: homogeneous) */
} NsgymLayout;

typedef struct {
  void* d_state; void* d_theta; int32_t* d_t; int32_t* d_istate;
  void* d_action;                /* staging used by nsgym_step_host */
  float* d_reward; uint8_t* d_flags; uint8_t* d_change;
  void* d_delta;                 /* NULL: delta not produced */
  float* d_obs;                  /* NULL: obs not produced (fp32 CartPole/MountainCar: obs == state) */
} NsgymBuffers;

typedef struct {                 /* host-side results of nsgym_step_host; NULL members are skipped */
  float* h_reward; uint8_t* h_flags; uint8_t* h_change; void* h_delta;
  void* h_state;                 /* classic control: packed state (fp32: the observation itself); gridworld: int32 cell */
  float* h_obs;
} NsgymHostOut;

typedef struct NsgymHandle NsgymHandle;

int nsgym_abi_version(void);
size_t nsgym_sizeof(int which);  /* 0 NsgymSlot, 1 NsgymSpec, 2 NsgymLayout, 3 NsgymBuffers, 4 NsgymHostOut, 5 NsgymSnapshotInfo */
const char* nsgym_last_error(void);

/* replaces: wrapper construction (base.py:222-294, classic_control.py:27-58, toy_text.py:28-84,282-340,547-603) */
int nsgym_create(const NsgymSpec* spec, NsgymHandle** out);
/* Heterogeneous batch (BASELINE config C4: "per-env distinct scheduler/update-function
 * opcodes"): env e of the handle runs its OWN scheduler / update function for every bound
 * parameter.  `rows` = NsgymSlot[n_envs][n_slots] in HOST memory, env-major; row (e, j) is what
 * the wrapper of env e would have been given for its j-th tunable parameter.  The key set is
 * shared: theta_index, constraint, partner_* and istate_plane of row (e, j) must equal those
 * of spec->slots[j] (istate_plane >= 0 there if ANY env needs a cursor for slot j); everything
 * else -- opcodes, range, coefficients, pool offsets, istate_init -- is per env.  The library
 * lowers the rows, keeps only the words that actually differ between envs as SoA planes in
 * device memory (owned by the handle) and reports their size in NsgymLayout.row_bytes_per_env. */
int nsgym_create_rows(const NsgymSpec* spec, const NsgymSlot* rows, NsgymHandle** out);
void nsgym_destroy(NsgymHandle* h);
int nsgym_layout(const NsgymHandle* h, int want_delta, int want_obs, NsgymLayout* out);
int nsgym_bind(NsgymHandle* h, const NsgymBuffers* buffers);

/* replaces: NSWrapper.reset + subclass resets (base.py:365-431, classic_control.py:102-109,
 * toy_text.py:202-210,382-399,653-667).  d_mask NULL = all envs, else uint8[N] (non-zero = reset).
 * d_inj_uniform NULL = native Philox draws, else double[L][N] (lanes: oracle/streams.py). */
int nsgym_reset(NsgymHandle* h, const uint8_t* d_mask, const double* d_inj_uniform, void* stream);

/* replaces: <wrapper>.step (file:line list at the top).  d_action NULL = the bound staging
 * buffer.  d_inj_uniform double[L][N] / d_inj_normal double[P][N] or NULL (native Philox).
 * skip_updates != 0 = planning env with in_sim_change False (classic_control.py:70-75). */
int nsgym_step(NsgymHandle* h, const void* d_action, const double* d_inj_uniform,
               const double* d_inj_normal, int skip_updates, void* stream);

/* One host call for a logical batch made of several handles (BASELINE config C4: CartPole + FrozenLake
 * envs with per-env rows are bucketed by env kind, one handle per kind): steps every handle on `stream`,
 * back to back.  d_actions[k] NULL (or d_actions NULL) = handle k's bound staging buffer. */
int nsgym_step_many(NsgymHandle* const* handles, const void* const* d_actions, int n, int skip_updates, void* stream);

/* Result packaging of NSWrapper.step (base.py:314-361) for the batch, in one launch: splits the
 * flag / change-mask / time words the step left in the bound buffers into what the wrapper
 * returns -- d_terminated, d_truncated, d_was_reset uint8[N] (0 / 1), d_relative_time int32[N],
 * d_env_change uint8[n_slots][N] (ground-truth env_change of every bound parameter).  Any
 * pointer may be NULL; byte arrays must be 4-byte aligned, d_relative_time 16-byte aligned. */
int nsgym_unpack(NsgymHandle* h, uint8_t* d_terminated, uint8_t* d_truncated, uint8_t* d_was_reset,
                 int32_t* d_relative_time, uint8_t* d_env_change, void* stream);

/* Episode bookkeeping for the batch in ONE launch over the outputs of the last step (what
 * evaluate/run_experiment.py:91-148 does per env in Python: episode_reward += reward, count steps,
 * stop at done / truncated), plus the counters a caller would otherwise need warnings for.
 * d_running_return double[N] and d_running_length int32[N] are caller-owned running sums (zero them
 * once); d_totals double[NSGYM_STAT_COUNT] accumulates over calls (zero it to start a report
 * interval; sum it over ranks with one all-reduce):
 *   STEPS env-steps taken (autoreset calls excluded), EPISODES finished, RETURN_SUM / LENGTH_SUM of
 *   the finished episodes, TERMINATED, TRUNCATED, REJECTED steps on which the constraint checker
 *   rejected a fired update (ConstraintViolationWarning, classic_control.py:87-92), BAD_DIST steps
 *   on which the reference would have raised (NSGYM_FLAG_BAD_DIST). */
enum { NSGYM_STAT_STEPS = 0, NSGYM_STAT_EPISODES, NSGYM_STAT_RETURN_SUM, NSGYM_STAT_LENGTH_SUM,
       NSGYM_STAT_TERMINATED, NSGYM_STAT_TRUNCATED, NSGYM_STAT_REJECTED, NSGYM_STAT_BAD_DIST, NSGYM_STAT_COUNT };
int nsgym_episode_stats(NsgymHandle* h, double* d_running_return, int32_t* d_running_length, double* d_totals,
                        void* stream);

/* Same call with HOST buffers: copies actions host->device, steps, copies the requested
 * results device->host, in `n_chunks` pipelined chunks over the handle's own copy streams, and
 * returns after the last copy has completed (synchronous).  This is the end-to-end call a
 * host-side agent loop makes.  `stream` is the stream earlier calls on this handle were issued on
 * (nsgym_reset, nsgym_step, nsgym_fanout ...): the pipeline is ordered after the work queued there.
 * On an error nothing of the step is left in flight. */
int nsgym_step_host(NsgymHandle* h, const void* h_action, const NsgymHostOut* out, int n_chunks, void* stream);

/* Page-locked host memory for nsgym_step_host buffers (callers without torch): portable across
 * devices; write_combined != 0 asks for write-combined memory (host writes / device reads only: the
 * action buffer).  NULL on failure (nsgym_last_error). */
void* nsgym_alloc_host(size_t bytes, int write_combined);
void nsgym_free_host(void* p);

/* K fused steps with state, theta and cursors in registers under a device-side policy
 * (MCTS-style random rollouts, benchmark_algorithms/MCTS.py:162-181).  policy 0 = uniform
 * random action.  d_return float[N] += sum of rewards (discounted by gamma^k);
 * d_length int32[N] += steps taken before the first episode end.  With
 * NSGYM_AUTORESET_NEXT_STEP envs keep cycling through episodes inside the K steps; with
 * NSGYM_AUTORESET_NONE a lane stops at its first episode end (MCTS default policy,
 * MCTS.py:162-181) and a lane that had already ended contributes nothing. */
int nsgym_rollout(NsgymHandle* h, int k_steps, int policy, float gamma, float* d_return,
                  int32_t* d_length, int skip_updates, void* stream);

/* The same fused rollout under a device-side LINEAR policy (shared by the batch, or one per env
 * when per_env != 0 -- population search evaluates N policies in one launch).
 * Classic control: d_policy = float[A][O + 1] (row a: O weights, then the bias) per policy, over the
 * float32 observation of the env kind (O = 4 CartPole, 6 Acrobot, 2 MountainCar(Continuous),
 * 3 Pendulum); Discrete action spaces take argmax_a (first maximum; A = 2 CartPole, 3 Acrobot /
 * MountainCar), Box action spaces (A = 1) take the score itself, which the env clips.
 * Gridworlds: a linear policy on the one-hot cell is a table: d_policy = uint8[nrow * ncol],
 * action = table[cell].  Steps, flags, autoreset and accumulators as in nsgym_rollout. */
int nsgym_rollout_linear(NsgymHandle* h, int k_steps, const void* d_policy, int per_env, float gamma,
                         float* d_return, int32_t* d_length, int skip_updates, void* stream);

/* Planning envs.
 * replaces: get_planning_env + __deepcopy__ (classic_control.py:120-186, toy_text.py:471-511,
 * 669-711) for a batch, with the fan-out MCTS-style consumers need (benchmark_algorithms/MCTS.py:
 * 130-131 makes one deepcopy per simulation): lane r * fanout + k of `dst` becomes a copy of env r
 * of `src` -- state, episode time, parameters / sampling table, list cursors.  `dst` must be a
 * bound handle of the same env kind, precision and slot list with src.n_envs * fanout envs.
 * theta_from_init != 0: the planner is not told the current parameters
 * (delta_change_notification False): theta <- initial values, cursors still carried.
 * The copy starts a fresh TimeLimit count (the reference builds a new gym.make chain), while
 * the NS time t carries on; dst is left "reset" and steps / rolls out like any handle (use
 * skip_updates for in_sim_change False).  Philox streams of dst are keyed by dst's own seed. */
int nsgym_fanout(const NsgymHandle* src, NsgymHandle* dst, int fanout, int theta_from_init, void* stream);

/* Device-side snapshot / restore of everything a step mutates (state, theta, t, cursors), packed
 * in that order into a caller buffer of nsgym_snapshot_bytes(); the host-side counters -- the Philox
 * step counter and, for a planning copy, the TimeLimit steps counted since the copy -- travel
 * through *info.  restore(snapshot(x)) followed by the same calls reproduces the same results bit
 * for bit (rewinding a search: snapshot, rollout, restore, any number of times). */
typedef struct { uint64_t step_index; int32_t plan_elapsed; int32_t _reserved; } NsgymSnapshotInfo;
size_t nsgym_snapshot_bytes(const NsgymHandle* h);
int nsgym_snapshot(NsgymHandle* h, void* d_dst, NsgymSnapshotInfo* info, void* stream);
int nsgym_restore(NsgymHandle* h, const void* d_src, const NsgymSnapshotInfo* info, void* stream);

/* Time-indexed transition table of one gridworld env.
 * replaces: reading unwrapped.P (NSFrozenLakeWrapper._update_transition_prob_table toy_text.py:
 * 426-447, NSCliffWalkingWrapper._build_P_from_outcomes :86-138) / Bridge.transition_matrix
 * (envs/Bridge.py:189-221) after every step of an episode -- the input of value iteration and of
 * the PAMCTS / RATS-style planners (evaluate/metrics.py:244-297, benchmark_algorithms/rats.py) and
 * what the unimplemented extract_oracle_transition_table (toy_text.py:513-519) was to return.
 * For env `env` and NS times t = 0 .. n_times-1 of an episode started by a reset:
 *   d_prob   double[n_times][S][4][D]  probability of outcome k of action a in cell s at time t
 *   d_next   int32 [S][4][D], d_reward float[S][4][D], d_done uint8[S][4][D]   (time-invariant)
 * D = 3 (4 for CliffWalking) outcomes in the reference's order [a, a+1, a-1(, a+2)]; absorbing rows
 * (FrozenLake G / H, Bridge H / G) have the single outcome (1.0, s, r, True) in k = 0 and zeros
 * after it.  Deterministic rules give THE table; stochastic rules are sampled from the env's
 * Philox stream as an episode starting at the handle's current step index would draw them. */
int nsgym_transition_table(NsgymHandle* h, int64_t env, int n_times, double* d_prob, int32_t* d_next,
                           float* d_reward, uint8_t* d_done, void* stream);

/* Fire-test + advance stages only (a1 + a2/a3), for known-answer checks of the update
 * functions: applies slot `slot` to d_param (real[N] scalars, or double[D][N]) at times
 * d_time[N]; writes new values in place, d_flag uint8[N], d_delta real[N]. */
int nsgym_eval_update(NsgymHandle* h, int slot, void* d_param, const int32_t* d_time,
                      int32_t* d_istate, uint8_t* d_flag, void* d_delta,
                      const double* d_inj_uniform, const double* d_inj_normal, int64_t n,
                      void* stream);

/* replaces: ns_gym/utils.py:55-94 (wasserstein_distance on indices) as the step kernels compute it
 * for delta_change (base.py:192-203): d_u, d_v are double[dim][n] (dim 3 or 4); d_out[n] = the
 * kernels' value (quotients through one shared reciprocal refinement per divisor), d_ref[n] = the
 * same sum with plain IEEE divisions.  Test entry: the two must agree bit for bit. */
int nsgym_eval_w1(int dim, const double* d_u, const double* d_v, double* d_out, double* d_ref,
                  int64_t n, void* stream);

/* Test entry: the NATIVE random draws of the handle's Philox streams (key = seed, counter = global env
 * id, step index, block), made by the very device functions the step kernels call, for envs 0..n-1
 * at Philox step `step_index` -- so the distributions of the draws the throughput kernels consume can
 * be tested directly, and a host restatement (tests/philox_np.py) can be held to them value by value.
 * d_out is double[planes][n] (float draws convert exactly).  Classic-control handles (in their
 * precision): NORMAL the standard normal of parameter lane `lane` (update_functions/single_param.py:
 * 79,111,149,345,446 draw rng.normal(mu, sigma) = mu + sigma z); RESET_UNIFORMS 4 planes, the
 * initial-state uniforms.  Gridworld handles: DYN_UNIFORM the slip uniform (FrozenLake / Cliff
 * categorical_sample, Bridge np.random.choice); DIRICHLET n_dist planes, RandomCategorical's
 * Dirichlet(1,..,1) (distribution.py:37-38), `t` = redraw attempt.  Both: SCHED_UNIFORM the scheduler
 * uniform of lane `lane` at episode time `t` (schedulers.py:28,177); GEOMETRIC MemorylessScheduler's
 * inter-fire time Geometric(p) drawn from it (schedulers.py:112-113). */
enum { NSGYM_DRAW_NORMAL = 0, NSGYM_DRAW_SCHED_UNIFORM = 1, NSGYM_DRAW_RESET_UNIFORMS = 2, NSGYM_DRAW_GEOMETRIC = 3,
       NSGYM_DRAW_DYN_UNIFORM = 4, NSGYM_DRAW_DIRICHLET = 5,
       NSGYM_DRAW_BOX_MULLER_SWEEP = 6 /* fp32 handles: the Box-Muller transform itself on chosen random bits --
                                          radius bits step_index + i (sweep all 2^24), angle bits t */ };
int nsgym_eval_draws(NsgymHandle* h, int what, int lane, int t, double p, uint64_t step_index, double* d_out,
                     int64_t n, void* stream);

/* Handle options.  NSGYM_OPT_GENERAL_KERNELS != 0: always launch the general kernel instantiations
 * (all rule classes, injection-capable) instead of the lean ones the library would pick for this
 * program -- same results (bit for bit in fp64 mode), used by the tests to tie the lean kernels to
 * the oracle-checked general ones.
 * NSGYM_OPT_SPECIALIZE: program-specialised kernels.  A handle whose program stays in the lean classes
 * can run a step kernel compiled at run time (NVRTC, ~0.3 s once per distinct program in the process) from
 * the library's own device code with the lowered program as a compile-time constant: the interpreter's
 * constant-bank loads, uniform branches, range / modulo tests and select masks fold away.  Same results as
 * the precompiled lean kernel (bit for bit in fp64 mode; fp32 may differ in the last bit where the folded
 * constants change an FMA contraction).  -1 (default): batches of >= 32768 envs; 0: never; 1: always.
 * Programs with slow-class slots specialise as well (general kernel class, slow class unrolled over the slots);
 * so do per-env rows of either class; injected tables and NSGYM_OPT_GENERAL_KERNELS keep the precompiled kernels.  Gridworld
 * batches of >= 2^21 envs with 16-byte aligned planes run a tiled variant whose env records arrive in shared
 * memory through TMA bulk copies (NSGYM_B200_NO_TILED=1 turns it off).
 * Environment: NSGYM_B200_NO_JIT=1 disables it process-wide, NSGYM_B200_NVRTC names libnvrtc.so.12 ("none": act
 * as if absent), NSGYM_B200_JIT_VERBOSE=1 prints why a specialisation was not possible (the precompiled kernel
 * runs then), NSGYM_B200_JIT_DUMP=dir keeps the generated sources and cubins. */
enum { NSGYM_OPT_GENERAL_KERNELS = 1, NSGYM_OPT_SPECIALIZE = 2 };
int nsgym_set_option(NsgymHandle* h, int option, int64_t value);
/* 1 when the handle's last step / rollout launch went to a program-specialised kernel */
int nsgym_last_kernel_specialized(const NsgymHandle* h);
/* Generate and compile (no device needed) the specialised single-step (rollout = 0) or fused-rollout
 * (rollout = 1, uniform-random policy) kernel of `spec` as a handle with / without delta and
 * float32-observation buffers would run it; copies the source and the compiler log out when asked.
 * Returns the cubin size, -2 when the program does not specialise, -3 when NVRTC is missing or the
 * compilation fails (nsgym_last_error says which). */
int nsgym_jit_check(const NsgymSpec* spec, int rollout, int want_delta, int want_obs, char* source, size_t source_len,
                    char* log, size_t log_len);
/* process-wide counters of the specialiser; returns 1 when it is enabled */
int nsgym_jit_stats(int64_t* compiled, int64_t* hits, int64_t* failed, char* last_failure, size_t len);

/* replaces: reset(seed=...) reseeding (base.py:386-388, 412-421): re-keys the Philox streams */
void nsgym_set_seed(NsgymHandle* h, uint64_t seed);
uint64_t nsgym_step_index(const NsgymHandle* h);        /* Philox counter (launches so far) */
void nsgym_set_step_index(NsgymHandle* h, uint64_t v);
int64_t nsgym_launch_count(const NsgymHandle* h);       /* kernels launched by this handle */
/* Which kernel instantiation the handle's last step / rollout launched: the LEAN ones (what throughput
 * runs use: native draws, fast [+ medium] rule classes only) or the GENERAL ones (every rule class,
 * injection-capable -- what parity tests with injected tables run); -1 before the first step. */
enum { NSGYM_KERNEL_LEAN_FAST = 0, NSGYM_KERNEL_LEAN_MEDIUM = 1, NSGYM_KERNEL_GENERAL = 2,
       NSGYM_KERNEL_ROWS_LEAN = 3, NSGYM_KERNEL_ROWS_GENERAL = 4 };
int nsgym_last_kernel_class(const NsgymHandle* h);

#ifdef __cplusplus
}
#endif
#endif /* NSGYM_B200_H */
