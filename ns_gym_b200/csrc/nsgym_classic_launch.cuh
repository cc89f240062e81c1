// nsgym_classic_launch.cuh -- host launcher for the classic-control kernels, instantiated
// once per precision (nsgym_f32.cu with FMA contraction, nsgym_f64.cu with -fmad=false).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <limits>
#include <string>
#include <type_traits>
#include <utility>

#include "nsgym_device.cuh"
#include "nsgym_host.h"

namespace nsg {

template <typename R> struct TrueMin;
template <> struct TrueMin<float> { static constexpr float value = 1.401298464324817e-45f; };
template <> struct TrueMin<double> { static constexpr double value = 4.9406564584124654e-324; };

// Lower one ABI slot to the kernel's form (SlotT): fast / slow class, range gate in unsigned
// form, multiply-high modulo, affine coefficients, constraint threshold.
// t_max: largest NS time the handle can reach (TimeLimit + autoreset), or 2^28
inline int64_t reachable_t_max(const NsgymSpec& spec) {
  return (spec.autoreset == NSGYM_AUTORESET_NEXT_STEP && spec.max_episode_steps > 0)
             ? int64_t(spec.max_episode_steps) + 1 : (int64_t(1) << 28);
}

template <typename R>
static SlotT<R> lower_slot(const NsgymSlot& a, int lane, int64_t t_max = (int64_t(1) << 28)) {
  SlotT<R> b{};
  b.lane = lane;
  b.sched_op = a.sched_op; b.upd_op = a.upd_op; b.constraint = a.constraint;
  b.istate_plane = a.istate_plane; b.istate_init = a.istate_init;
  for (int k = 0; k < 4; ++k) { b.si[k] = a.si[k]; b.ui[k] = a.ui[k]; }
  b.sf[0] = a.sf[0]; b.sf[1] = a.sf[1];
  for (int k = 0; k < 6; ++k) b.uf[k] = R(a.uf[k]);
  // ---- range gate: start <= t <= end  <=>  unsigned(t - start) <= unsigned(end - start) ----
  const int32_t t_cap = (1 << 28);                       // t never exceeds T_TIME_MASK
  int32_t lo = a.start < 0 ? 0 : a.start, hi = a.end > t_cap ? t_cap : a.end;
  if (lo > t_cap || hi < lo) { b.start = INT32_MAX; b.span = 0; }   // never in range
  else { b.start = lo; b.span = hi - lo; }
  // ---- scheduler: fast class = in range && (t mod d) < on ----
  b.mod_d = 0; b.mod_magic = 0; b.mod_on = INT32_MAX;    // Continuous: t - 0 < INT_MAX
  switch (a.sched_op) {
    case NSGYM_SCHED_CONTINUOUS: break;
    case NSGYM_SCHED_PERIODIC:
      if (a.si[0] == 1) break;                           // t % 1 == 0 always
      if (a.si[2]) { b.mod_d = a.si[0]; b.mod_magic = a.si[2]; b.mod_on = 1; }
      else b.flags |= SF_SLOW_SCHED;
      break;
    case NSGYM_SCHED_BURST:
      if (a.si[1] == 1) { b.mod_on = a.si[0] > 0 ? INT32_MAX : 0; break; }   // (t % 1) < on
      if (a.si[2]) { b.mod_d = a.si[1]; b.mod_magic = a.si[2]; b.mod_on = a.si[0]; }
      else b.flags |= SF_SLOW_SCHED;
      break;
    default: b.flags |= SF_SLOW_SCHED; break;
  }
  // ---- update: fast class = ((A y + B) + noise) + C t ----
  double A = 1.0, B = 0.0, Ct = 0.0;
  switch (a.upd_op) {
    case NSGYM_UPD_NOP: break;
    case NSGYM_UPD_ADD: B = a.uf[0]; break;
    case NSGYM_UPD_ADD_T: Ct = a.uf[0]; break;
    case NSGYM_UPD_MUL: A = a.uf[0]; break;
    case NSGYM_UPD_RW: B = a.uf[0]; Ct = a.uf[3]; b.flags |= SF_NORMAL; break;
    case NSGYM_UPD_LERP: case NSGYM_UPD_MUL_EXP: case NSGYM_UPD_SIGMOID:
      b.flags |= SF_MEDIUM;
      break;
    case NSGYM_UPD_ADD_SIN:    // fp32: the bounded sine is good for t <= 1e5; fp64 calls the accurate sin
      b.flags |= (std::is_same<R, double>::value || t_max <= 100000) ? SF_MEDIUM : SF_SLOW_UPD;
      break;
    case NSGYM_UPD_D_UNIFORM: b.flags |= SF_SLOW_UPD | SF_D_AFFINE; break;
    default: b.flags |= SF_SLOW_UPD; break;
  }
  b.fa[0] = R(A); b.fa[1] = R(B); b.fa[2] = R(Ct);
  b.mu = R(a.uf[1]); b.sigma = R(a.uf[2]);
  // `v <= 0` -> v <= 0;  `v < 0` -> v <= -(smallest subnormal);  none -> v <= -inf (never)
  if (a.constraint == NSGYM_CONS_REJECT_LE0) b.reject_le = R(0);
  else if (a.constraint == NSGYM_CONS_REJECT_LT0) b.reject_le = -TrueMin<R>::value;
  else b.reject_le = -std::numeric_limits<R>::infinity();
  return b;
}

inline bool slot_draws_block0(const NsgymSlot& a, int lane) {
  return lane < 2 && (a.upd_op == NSGYM_UPD_RW || a.upd_op == NSGYM_UPD_OU || a.upd_op == NSGYM_UPD_BRW);
}

template <typename R, int NP>
static void program_common(ProgramT<R, NP>& P, const NsgymSpec& spec, const DevicePools& pools) {
  P.max_steps = spec.max_episode_steps;
  P.autoreset = spec.autoreset;
  P.persistent = spec.persistent_params;
  for (int i = 0; i < NSGYM_MAX_THETA; ++i) P.theta_default[i] = R(spec.theta_init[i][0]);
  P.pool_f = pools.pool_f; P.pool_i = pools.pool_i; P.bitmap = pools.bitmap;
}

// classic control: slots in tunable_params order, NP = spec.n_slots exactly
template <typename R, int NP>
static ProgramT<R, NP> build_program(const NsgymSpec& spec, const DevicePools& pools) {
  ProgramT<R, NP> P{};
  program_common(P, spec, pools);
  for (int q = 0; q < NSGYM_MAX_THETA; ++q) P.base[q] = P.theta_default[q];
  for (int j = 0; j < NP; ++j) {
    const NsgymSlot& a = spec.slots[j];
    SlotT<R>& b = P.slot[j];
    b = lower_slot<R>(a, j, reachable_t_max(spec));
    b.theta_index = a.theta_index;
    b.init = R(spec.theta_init[a.theta_index][0]);
    b.partner_slot = a.partner_slot;
    b.partner_default = R(spec.theta_init[a.partner_index >= 0 && a.partner_index < NSGYM_MAX_THETA ? a.partner_index : 0][0]);
    P.sel[j][a.theta_index] = 0xFFFFFFFFu;
    P.base[a.theta_index] = R(0);
    P.bound_mask |= 1 << a.theta_index;
    if (b.flags & (SF_SLOW_SCHED | SF_SLOW_UPD)) P.slow_j[P.n_slow++] = j;
  }
  P.n_bound = NP;
  return P;
}

// gridworlds: slots indexed by theta index (0 = P, 1 = P_left, 2 = P_right), dict position in `lane`
template <int MAXP>
static ProgramT<double, MAXP> build_program_by_index(const NsgymSpec& spec, const DevicePools& pools) {
  ProgramT<double, MAXP> P{};
  program_common(P, spec, pools);
  for (int q = 0; q < MAXP; ++q) P.slot[q].istate_plane = -1;
  for (int j = 0; j < spec.n_slots; ++j) {
    const int q = spec.slots[j].theta_index;
    if (q < 0 || q >= MAXP) continue;
    P.slot[q] = lower_slot<double>(spec.slots[j], j);
    P.slot[q].theta_index = q;
    P.bound_mask |= 1 << q;
    P.n_bound += 1;
  }
  return P;
}

// canonical row words of one lowered slot (layout: nsgym_device.cuh, HetT)
template <typename R>
static void row_words(const SlotT<R>& b, const NsgymSlot& a, int32_t (&iw)[kRowInt], double (&rw)[kRowReal],
                      double (&dw)[kRowDbl]) {
  int flags = b.flags;
  int mod_d = 0, mod_on_enc = 0;
  if (!(flags & SF_SLOW_SCHED)) {
    // the fast modulo as (modulus, on-count + 1 | 0 = always); a pair that does not fit 16 bits each takes
    // the slow scheduler switch instead
    if (b.mod_d > 0xFFFF || (b.mod_on != INT32_MAX && (b.mod_on < 0 || b.mod_on >= 0xFFFF))) flags |= SF_SLOW_SCHED;
    else { mod_d = b.mod_d; mod_on_enc = b.mod_on == INT32_MAX ? 0 : b.mod_on + 1; }
  }
  iw[RI_OPS] = flags | (a.sched_op << 5) | (a.upd_op << 9);
  iw[RI_START] = b.start; iw[RI_SPAN] = b.span;
  iw[RI_MODD] = mod_d; iw[RI_MODON] = mod_on_enc;
  iw[RI_SI0] = a.si[0]; iw[RI_SI1] = a.si[1]; iw[RI_UI0] = a.ui[0]; iw[RI_UI1] = a.ui[1];
  iw[RI_IINIT] = a.istate_init;
  if (flags & (SF_SLOW_UPD | SF_MEDIUM)) {
    for (int k = 0; k < kRowReal; ++k) rw[k] = double(b.uf[k]);
  } else {
    rw[0] = double(b.fa[0]); rw[1] = double(b.fa[1]); rw[2] = double(b.fa[2]);
    rw[3] = double(b.mu); rw[4] = double(b.sigma);
  }
  dw[0] = a.sf[0]; dw[1] = a.sf[1];
}

template <typename R, int NP>
static HetT<R, NP> build_het(const RowTable& t) {
  HetT<R, NP> H{};
  H.ints = t.d_int;
  H.reals = reinterpret_cast<const R*>(t.d_real);
  H.dbls = t.d_dbl;
  for (int j = 0; j < NP; ++j) {
    H.mask[j] = t.mask[j];
    for (int w = 0; w < kRowWords; ++w) H.plane[j][w] = t.plane[j][w];
    for (int w = 0; w < kRowInt; ++w) { H.idef[j][w] = t.def_int[j][w]; H.shift[j][w] = t.shift[j][w]; H.bits[j][w] = t.bits[j][w]; }
    for (int w = 0; w < kRowReal; ++w) H.rdef[j][w] = R(t.def_real[j][w]);
    for (int w = 0; w < kRowDbl; ++w) H.ddef[j][w] = t.def_dbl[j][w];
  }
  return H;
}

// block 0 of the Philox stream is consumed by the normals of lanes 0 / 1 and by the gridworld slip
// draw on every step: compute it once, before any branch.  When only next-step autoreset needs it
// (initial-state draws of the few lanes that reset), it is computed in the reset branch instead:
// a warp without a resetting lane -- most warps of MountainCar / Pendulum / Acrobot, whose
// episodes last hundreds of steps -- skips the ten rounds altogether.
inline bool wants_prefetch(const NsgymSpec& spec) {
  if (is_grid_kind(spec.env_kind)) return true;
  for (int j = 0; j < spec.n_slots; ++j)
    if (slot_draws_block0(spec.slots[j], j)) return true;
  return false;
}

template <typename R>
static StepIO<R> build_io(const LaunchIO& a) {
  StepIO<R> io{};
  io.state = reinterpret_cast<R*>(a.state); io.theta = reinterpret_cast<R*>(a.theta);
  io.t = a.t; io.istate = a.istate; io.action = a.action;
  io.reward = a.reward; io.flags = a.flags; io.change = a.change;
  io.delta = reinterpret_cast<R*>(a.delta); io.obs = a.obs;
  io.inj_u = a.inj_u; io.inj_z = a.inj_z; io.mask = a.mask;
  io.n = uint32_t(a.n); io.begin = uint32_t(a.begin); io.count = uint32_t(a.count);
  io.gid_offset = a.gid_offset; io.step_index = a.step_index;
  io.skip_updates = a.skip_updates; io.force_init = a.force_init;
  io.prefetch = (a.prefetch && !a.inj_u && !a.inj_z) ? 1 : 0;
  io.plan_elapsed = a.plan_elapsed;
  io.sched_replay = a.sched_replay;
  uint32_t k0 = uint32_t(a.seed), k1 = uint32_t(a.seed >> 32);
  for (int r = 0; r < 10; ++r) {            // Philox4x32 key schedule (Weyl sequence)
    io.rk[r][0] = k0; io.rk[r][1] = k1;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return io;
}

// ---- sources of the program-specialised kernels (nsgym_jit.cu) ----
// `dst = the host object's words`, for a pointer-free object of type `type`
template <typename T>
static std::string spec_assign(const std::string& dst, const std::string& type, const T& obj) {
  static_assert(sizeof(T) % 4 == 0, "a whole number of words");
  return "  " + dst + " = __builtin_bit_cast(" + type + ", SpecWords<" + std::to_string(sizeof(T) / 4) + ">{{" +
         jit::words(&obj, sizeof(T)) + "}});\n";
}
// launch facts fixed at compile time (device side: NoFix / SpecFix)
template <typename R>
static std::string spec_prelude(const char* header, const StepIO<R>& io, bool root) {
  return std::string("#include \"") + header + "\"\nnamespace nsg {\ntemplate <int N> struct SpecWords { uint32_t w[N]; };\n" +
         "struct SpecFix { static constexpr int prefetch = " + std::to_string(io.prefetch ? 1 : 0) +
         ", want_delta = " + std::to_string(io.delta ? 1 : 0) + ", has_obs = " + std::to_string(io.obs ? 1 : 0) +
         ", root = " + std::to_string(root ? 1 : -1) + ", rows_early = 1; };\n";
}
template <typename R>
static uint32_t spec_facts(const StepIO<R>& io, bool root) {
  return (io.prefetch ? 4u : 0u) | (io.delta ? 8u : 0u) | (io.obs ? 16u : 0u) | (root ? 32u : 0u);
}

// The single-step kernel of a lean classic-control program: the body the precompiled kernel runs
// (classic_step_body), with the lowered program rebuilt as a constexpr object from the words of the
// host's ProgramHeadT and the launch facts fixed in SpecFix.
// `pools`: a program with slow-class slots (level 2) reads its value lists / window lists / bitmaps through
// the pool pointers, which stay kernel parameters
static std::string spec_program_object(const std::string& prog, bool pools) {
  if (!pools) return "  constexpr nsg::" + prog + " P = nsg::spec_program();\n";
  return "  constexpr nsg::" + prog + " P0 = nsg::spec_program();\n  nsg::" + prog + " P = P0;\n"
         "  P.pool_f = pp.pool_f; P.pool_i = pp.pool_i; P.bitmap = pp.bitmap;\n";
}
template <typename R, int KIND, int NP>
static std::string spec_step_source(const ProgramT<R, NP>& P, int level, const StepIO<R>& io, bool root) {
  const std::string real = std::is_same<R, float>::value ? "float" : "double";
  const ProgramHeadT<R, NP>& head = P;
  const std::string prog = "ProgramT<" + real + ", " + std::to_string(NP) + ">";
  const std::string headt = "ProgramHeadT<" + real + ", " + std::to_string(NP) + ">";
  const bool pools = level >= 2;
  std::string s = spec_prelude<R>("nsgym_device.cuh", io, root);
  s += "__device__ constexpr " + prog + " spec_program() {\n  " + prog + " P{};\n" +
       spec_assign("static_cast<" + headt + "&>(P)", headt, head) + "  return P;\n}\n}  // namespace nsg\n";
  s += "extern \"C\" __global__ void __launch_bounds__(256, nsg::classic_spec_min_blocks<" + real + ", " +
       std::to_string(KIND) + ", " + std::to_string(level) + ">())\nnsgym_spec_classic_step(const __grid_constant__ nsg::StepIO<" +
       real + "> io" + (pools ? ", const __grid_constant__ nsg::PoolPtrs pp" : "") + ") {\n" + spec_program_object(prog, pools) +
       "  nsg::classic_step_body<" + real + ", " +
       std::to_string(KIND) + ", " + std::to_string(NP) + ", " + std::to_string(level) + ", nsg::SpecFix>(P, io);\n}\n";
  return s;
}

// Can the tiled kernels stream this launch's planes with TMA bulk copies?  Every plane offset must be a
// multiple of 16 bytes: base pointers, the sub-range start and the plane stride (n envs).
template <typename R>
static bool tiled_ok(const StepIO<R>& io, bool kind_supported) {
  static const bool off = [] { const char* e = std::getenv("NSGYM_B200_NO_TILED"); return e && *e && *e != '0'; }();
  // (batches below 2^21 envs keep the plain kernel: FrozenLake 2^20 envs -- L2-resident -- 13.7 us plain, 15.6 us tiled)
  if (off || !kind_supported || io.count < (1u << 21)) return false;
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  return (io.n % 4u) == 0 && (io.begin % 4u) == 0 && al(io.state) && al(io.theta) && al(io.t) && al(io.istate) && al(io.action);
}
// tiles a block advances back to back (NSGYM_B200_TILES overrides): measured on 2^24 envs (Bridge /
// FrozenLake, 5 resident blocks): 2 tiles 0.91 / 0.84 of the copy peak, 4: 0.985 / 0.93, 6: 0.985 / 0.95,
// 16: 0.90 / 0.81; smaller batches keep at least ~3 waves of blocks
static int tiles_per_block_for(uint32_t full_tiles) {
  static const int forced = [] { const char* e = std::getenv("NSGYM_B200_TILES"); return e ? std::atoi(e) : 0; }();
  if (forced > 0) return forced;
  const int t = int(full_tiles / 2220u);
  return t < 2 ? 2 : (t > 6 ? 6 : t);
}

// `spec_rows()`: the row layout (which words vary, their planes, the shared defaults) as a constant
template <typename R, int NP>
static std::string spec_rows_source(const HetT<R, NP>& H) {
  const std::string real = std::is_same<R, float>::value ? "float" : "double";
  const std::string headt = "HetHeadT<" + real + ", " + std::to_string(NP) + ">";
  const HetHeadT<R, NP>& head = H;
  return "namespace nsg {\n__device__ constexpr " + headt + " spec_rows() {\n  " + headt + " H{};\n" +
         spec_assign("H", headt, head) + "  return H;\n}\n}  // namespace nsg\n";
}
// statements that rebuild a HetT named H from spec_rows() and the kernel parameter `hp` (HetPtrs)
static std::string spec_rows_object(const std::string& real, int np) {
  const std::string args = real + ", " + std::to_string(np);
  return "  constexpr nsg::HetHeadT<" + args + "> HH = nsg::spec_rows();\n  nsg::HetT<" + args + "> H{};\n"
         "  static_cast<nsg::HetHeadT<" + args + ">&>(H) = HH;\n"
         "  H.ints = hp.ints; H.reals = static_cast<const " + real + "*>(hp.reals); H.dbls = hp.dbls;\n";
}

// The single-step kernel of a batch with lean per-env rows (classic_step_het_body): program and row layout
// are constants, so every lane loads exactly the row words that vary and the shared ones are immediates.
template <typename R, int KIND, int NP>
static std::string spec_step_rows_source(const ProgramT<R, NP>& P, const HetT<R, NP>& H, const StepIO<R>& io, bool root,
                                         bool lean = true) {
  const std::string real = std::is_same<R, float>::value ? "float" : "double";
  const ProgramHeadT<R, NP>& head = P;
  const std::string prog = "ProgramT<" + real + ", " + std::to_string(NP) + ">";
  const std::string headt = "ProgramHeadT<" + real + ", " + std::to_string(NP) + ">";
  std::string s = spec_prelude<R>("nsgym_device.cuh", io, root);
  s += "__device__ constexpr " + prog + " spec_program() {\n  " + prog + " P{};\n" +
       spec_assign("static_cast<" + headt + "&>(P)", headt, head) + "  return P;\n}\n}  // namespace nsg\n";
  s += spec_rows_source<R, NP>(H);
  s += std::string("extern \"C\" __global__ void __launch_bounds__(256, ") + (lean ? "NSGYM_HET_LEAN_MIN_BLOCKS" : "NSGYM_HET_MIN_BLOCKS") +
       ")\nnsgym_spec_classic_step_rows("
       "const __grid_constant__ nsg::StepIO<" + real + "> io, const __grid_constant__ nsg::HetPtrs hp, "
       "const __grid_constant__ nsg::PoolPtrs pp) {\n"
       "  constexpr nsg::" + prog + " P0 = nsg::spec_program();\n  nsg::" + prog + " P = P0;\n"
       "  P.pool_f = pp.pool_f; P.pool_i = pp.pool_i; P.bitmap = pp.bitmap;\n" + spec_rows_object(real, NP) +
       "  nsg::classic_step_het_body<" + real + ", " + std::to_string(KIND) + ", " + std::to_string(NP) +
       ", " + (lean ? "true" : "false") + ", nsg::SpecFix>(P, H, io);\n}\n";
  return s;
}

// K fused steps of a lean classic-control program (classic_rollout_body); `lin`: linear rollout policy
template <typename R, int KIND, int NP>
static std::string spec_rollout_source(const ProgramT<R, NP>& P, int level, const StepIO<R>& io, bool root, bool lin) {
  const std::string real = std::is_same<R, float>::value ? "float" : "double";
  const ProgramHeadT<R, NP>& head = P;
  const std::string prog = "ProgramT<" + real + ", " + std::to_string(NP) + ">";
  const std::string headt = "ProgramHeadT<" + real + ", " + std::to_string(NP) + ">";
  std::string s = spec_prelude<R>("nsgym_device.cuh", io, root);
  s += "__device__ constexpr " + prog + " spec_program() {\n  " + prog + " P{};\n" +
       spec_assign("static_cast<" + headt + "&>(P)", headt, head) + "  return P;\n}\n}  // namespace nsg\n";
  const bool pools = level >= 2;
  s += "extern \"C\" __global__ void __launch_bounds__(256)\nnsgym_spec_classic_rollout(const __grid_constant__ nsg::StepIO<" + real +
       "> io, const __grid_constant__ nsg::RolloutArgs ra" + (pools ? ", const __grid_constant__ nsg::PoolPtrs pp" : "") +
       ") {\n" + spec_program_object(prog, pools) +
       "  const nsg::HetT<" + real + ", " + std::to_string(NP) + "> no_rows{};\n  nsg::classic_rollout_body<" + real + ", " +
       std::to_string(KIND) + ", " + std::to_string(NP) + ", " + std::to_string(level) + ", false, " + (lin ? "true" : "false") +
       ", nsg::SpecFix>(P, no_rows, io, ra.k_steps, ra.gamma, ra.ret, ra.len, static_cast<const float*>(ra.pol), ra.pol_per_env);\n}\n";
  return s;
}

template <typename R, int KIND, int NP>
static cudaError_t launch_classic_knp(LaunchOp op, const NsgymSpec& spec, const DevicePools& pools,
                                      const LaunchIO& a, cudaStream_t stream) {
  const ProgramT<R, NP> P = build_program<R, NP>(spec, pools);
  LaunchIO a2 = a;
  a2.prefetch = wants_prefetch(spec) ? 1 : 0;
  const StepIO<R> io = build_io<R>(a2);
  const int block = 256;
  const unsigned grid = unsigned((a.count + block - 1) / block);
  if (grid == 0) return cudaSuccess;
  if (a.rows && a.rows->active) {       // heterogeneous handle: per-env rows
    if (a.spec_source) return cudaErrorNotSupported;
    if constexpr (NP == 0) return cudaErrorInvalidValue;
    else {
      const HetT<R, NP> H = build_het<R, NP>(*a.rows);
      const bool lean_rows = a.rows->lean && !a.general_kernels && !a.inj_u && !a.inj_z;
      if (a.kernel_class) *a.kernel_class = lean_rows ? NSGYM_KERNEL_ROWS_LEAN : NSGYM_KERNEL_ROWS_GENERAL;
      if (a.specialized) *a.specialized = 0;
      // rows of the general class (stochastic schedulers, cursor rules) specialise too: slot loop unrolled,
      // row layout constant, no injection code
      const bool rows_spec_ok = lean_rows || (!a.general_kernels && !a.inj_u && !a.inj_z);
      if (op == OP_STEP && rows_spec_ok && a.specialize) {
        const bool root = a.plan_elapsed < 0 && !a.skip_updates;
        const uint32_t facts = spec_facts(io, root) | 256u | (lean_rows ? 0u : 2048u);
        cudaKernel_t k = nullptr;
        if (!a.spec_cache || !a.spec_cache->find(facts, &k)) {
          k = jit::kernel(spec_step_rows_source<R, KIND, NP>(P, H, io, root, lean_rows), "nsgym_spec_classic_step_rows",
                          std::is_same<R, float>::value, nullptr);
          if (a.spec_cache) a.spec_cache->put(facts, k);
        }
        if (k) {
          HetPtrs hp{H.ints, H.reals, H.dbls};
          PoolPtrs pp{P.pool_f, P.pool_i, P.bitmap};
          void* args[] = {const_cast<StepIO<R>*>(&io), &hp, &pp};
          if (a.specialized) *a.specialized = 1;
          return cudaLaunchKernel(reinterpret_cast<const void*>(k), dim3(grid), dim3(block), args, 0, stream);
        }
      }
      switch (op) {
        case OP_STEP:
          if (lean_rows) classic_step_het_kernel<R, KIND, NP, true><<<grid, block, 0, stream>>>(P, H, io);
          else classic_step_het_kernel<R, KIND, NP, false><<<grid, block, 0, stream>>>(P, H, io);
          break;
        case OP_RESET: classic_reset_het_kernel<R, KIND, NP><<<grid, block, 0, stream>>>(P, H, io); break;
        case OP_ROLLOUT:
          classic_rollout_kernel<R, KIND, NP, 2, true><<<grid, block, 0, stream>>>(
              P, H, io, a.k_steps, a.gamma, a.ret, a.len, static_cast<const float*>(a.policy), a.policy_per_env);
          break;
      }
      return cudaGetLastError();
    }
  }
  const HetT<R, NP> no_rows{};
  // programs without slow-class slots run a lean instantiation (no rule switches compiled in):
  // level 0 = fast class only, level 1 = + inline medium rules, level 2 = everything
  // (injected random tables -- parity tests -- also take level 2: the lean kernels fold the
  // "injected?" tests away)
  int level = (P.n_slow > 0 || a.inj_u || a.inj_z || a.general_kernels) ? 2 : 0;
  if (level == 0)
    for (int j = 0; j < NP; ++j)
      if (P.slot[j].flags & SF_MEDIUM) level = 1;
  if (a.kernel_class) *a.kernel_class = level == 2 ? NSGYM_KERNEL_GENERAL : (level == 1 ? NSGYM_KERNEL_LEAN_MEDIUM : NSGYM_KERNEL_LEAN_FAST);
  constexpr bool kHasMedium = true;
  const unsigned lean_grid = unsigned((a.count + block * NSGYM_LEAN_EPT - 1) / (block * NSGYM_LEAN_EPT));
  if (a.specialized) *a.specialized = 0;
  // programs with slow-class slots specialise too (level 2: the rule switches fold to the one rule of each slot);
  // injected tables and NSGYM_OPT_GENERAL_KERNELS keep the precompiled general kernel
  const bool spec_ok = level < 2 || (P.n_slow > 0 && !a.inj_u && !a.inj_z && !a.general_kernels);
  PoolPtrs pp{P.pool_f, P.pool_i, P.bitmap};
  if (op == OP_STEP && spec_ok && (a.specialize || a.spec_source)) {
    // lean program: the kernel compiled for exactly this program (cached per distinct source)
    const bool root = a.plan_elapsed < 0 && !a.skip_updates;
    const uint32_t facts = uint32_t(level) | spec_facts(io, root);
    cudaKernel_t k = nullptr;
    if (a.spec_source || !a.spec_cache || !a.spec_cache->find(facts, &k)) {
      const std::string src = spec_step_source<R, KIND, NP>(P, level, io, root);
      if (a.spec_source) { *a.spec_source = src; return cudaSuccess; }
      k = jit::kernel(src, "nsgym_spec_classic_step", std::is_same<R, float>::value, nullptr);
      if (a.spec_cache) a.spec_cache->put(facts, k);
    }
    if (k) {
      // (a tiled variant with TMA-prefetched planes, as the gridworld kernels have, was measured for the
      // classic-control kernels too: no gain -- C1 fp32 166 -> 171 us, fp64 CartPole +2.5 %, the others
      // -1..-3 % -- their record is one 128-bit load and three words, and they already run at full occupancy)
      void* args[] = {const_cast<StepIO<R>*>(&io), &pp};      // (pp: level 2 only -- a lean kernel takes one parameter)
      if (a.specialized) *a.specialized = 1;
      return cudaLaunchKernel(reinterpret_cast<const void*>(k), dim3(level < 2 ? lean_grid : grid), dim3(block), args, 0, stream);
    }
  }
  if (op == OP_ROLLOUT && spec_ok && NP > 0 && (a.specialize || a.spec_source)) {
    const bool root = a.plan_elapsed < 0 && !a.skip_updates;
    const bool lin = a.policy != nullptr;
    const uint32_t facts = uint32_t(level) | spec_facts(io, root) | 64u | (lin ? 128u : 0u);
    cudaKernel_t k = nullptr;
    if (a.spec_source || !a.spec_cache || !a.spec_cache->find(facts, &k)) {
      const std::string src = spec_rollout_source<R, KIND, NP>(P, level, io, root, lin);
      if (a.spec_source) { *a.spec_source = src; return cudaSuccess; }
      k = jit::kernel(src, "nsgym_spec_classic_rollout", std::is_same<R, float>::value, nullptr);
      if (a.spec_cache) a.spec_cache->put(facts, k);
    }
    if (k) {
      RolloutArgs ra{a.k_steps, a.gamma, a.ret, a.len, a.policy, a.policy_per_env};
      void* args[] = {const_cast<StepIO<R>*>(&io), &ra, &pp};
      if (a.specialized) *a.specialized = 1;
      return cudaLaunchKernel(reinterpret_cast<const void*>(k), dim3(grid), dim3(block), args, 0, stream);
    }
  }
  if (a.spec_source) return cudaErrorNotSupported;   // nsgym_jit_check on a program of the general class: nothing to generate
  switch (op) {
    case OP_STEP:
      if (level == 2) classic_step_kernel<R, KIND, NP, 2><<<grid, block, 0, stream>>>(P, io);
      else if (level == 1) {
        if constexpr (kHasMedium) classic_step_kernel<R, KIND, NP, 1><<<lean_grid, block, 0, stream>>>(P, io);
        else return cudaErrorInvalidValue;
      } else classic_step_kernel<R, KIND, NP, 0><<<lean_grid, block, 0, stream>>>(P, io);
      break;
    case OP_RESET: classic_reset_kernel<R, KIND, NP, 2><<<grid, block, 0, stream>>>(P, io); break;
    case OP_ROLLOUT: {
      const float* pol = static_cast<const float*>(a.policy);   // linear policy (nsgym_rollout_linear) or NULL
      if (level == 2 || NP == 0) {
        classic_rollout_kernel<R, KIND, NP, 2><<<grid, block, 0, stream>>>(P, no_rows, io, a.k_steps, a.gamma, a.ret,
                                                                            a.len, pol, a.policy_per_env);
      } else if (level == 1) {
        if (pol)
          classic_rollout_kernel<R, KIND, NP, 1, false, true><<<grid, block, 0, stream>>>(
              P, no_rows, io, a.k_steps, a.gamma, a.ret, a.len, pol, a.policy_per_env);
        else
          classic_rollout_kernel<R, KIND, NP, 1><<<grid, block, 0, stream>>>(P, no_rows, io, a.k_steps, a.gamma, a.ret,
                                                                              a.len);
      } else {
        if (pol)
          classic_rollout_kernel<R, KIND, NP, 0, false, true><<<grid, block, 0, stream>>>(
              P, no_rows, io, a.k_steps, a.gamma, a.ret, a.len, pol, a.policy_per_env);
        else
          classic_rollout_kernel<R, KIND, NP, 0><<<grid, block, 0, stream>>>(P, no_rows, io, a.k_steps, a.gamma, a.ret,
                                                                              a.len);
      }
      break;
    }
  }
  return cudaGetLastError();
}

// one instantiation per exact slot count 0..NTH: the slot loops of the kernels are static
template <typename R, int KIND, int... NP>
static cudaError_t launch_classic_np(std::integer_sequence<int, NP...>, LaunchOp op, const NsgymSpec& spec,
                                     const DevicePools& pools, const LaunchIO& a, cudaStream_t stream) {
  cudaError_t e = cudaErrorInvalidValue;
  ((spec.n_slots == NP ? (e = launch_classic_knp<R, KIND, NP>(op, spec, pools, a, stream), 0) : 0), ...);
  return e;
}

// One translation unit per (precision, env kind) -- nsgym_classic_kind.cu compiled with
// -DNSGYM_TU_REAL / -DNSGYM_TU_KIND -- holds the kernels of that kind; the dispatchers in
// nsgym_f32.cu / nsgym_f64.cu only see this declaration.
template <typename R, int KIND>
cudaError_t launch_classic_kind(LaunchOp op, const NsgymSpec& spec, const DevicePools& pools, const LaunchIO& a,
                                cudaStream_t stream);

#ifdef NSGYM_TU_KIND
template <typename R, int KIND>
static cudaError_t launch_classic_k(LaunchOp op, const NsgymSpec& spec, const DevicePools& pools,
                                    const LaunchIO& a, cudaStream_t stream) {
  return launch_classic_np<R, KIND>(std::make_integer_sequence<int, KindTraits<KIND>::NTH + 1>{}, op, spec, pools, a,
                                    stream);
}

template <typename R, int KIND>
cudaError_t launch_classic_kind(LaunchOp op, const NsgymSpec& spec, const DevicePools& pools, const LaunchIO& a,
                                cudaStream_t stream) {
  return launch_classic_k<R, KIND>(op, spec, pools, a, stream);
}
#endif  // NSGYM_TU_KIND

template <typename R>
static cudaError_t launch_classic_t(LaunchOp op, const NsgymSpec& spec, const DevicePools& pools,
                                    const LaunchIO& a, cudaStream_t stream) {
  switch (spec.env_kind) {
    case NSGYM_ENV_CARTPOLE: return launch_classic_kind<R, NSGYM_ENV_CARTPOLE>(op, spec, pools, a, stream);
    case NSGYM_ENV_ACROBOT: return launch_classic_kind<R, NSGYM_ENV_ACROBOT>(op, spec, pools, a, stream);
    case NSGYM_ENV_MOUNTAINCAR: return launch_classic_kind<R, NSGYM_ENV_MOUNTAINCAR>(op, spec, pools, a, stream);
    case NSGYM_ENV_MOUNTAINCAR_CONT:
      return launch_classic_kind<R, NSGYM_ENV_MOUNTAINCAR_CONT>(op, spec, pools, a, stream);
    case NSGYM_ENV_PENDULUM: return launch_classic_kind<R, NSGYM_ENV_PENDULUM>(op, spec, pools, a, stream);
    default: return cudaErrorInvalidValue;
  }
}

template <typename R>
static cudaError_t launch_eval_scalar_t(const NsgymSpec& spec, const DevicePools& pools, int slot, void* param,
                                        const int32_t* time, int32_t* istate, uint8_t* flag, void* delta,
                                        const double* inj_u, const double* inj_z, int64_t n, uint64_t seed,
                                        uint64_t step_index, cudaStream_t stream) {
  ProgramT<R, 1> P{};
  P.max_steps = spec.max_episode_steps;
  P.slot[0] = lower_slot<R>(spec.slots[slot], slot);
  P.slot[0].istate_plane = istate ? 0 : -1;
  P.n_bound = 1;
  P.pool_f = pools.pool_f; P.pool_i = pools.pool_i; P.bitmap = pools.bitmap;
  LaunchIO a{};
  a.inj_u = inj_u; a.inj_z = inj_z; a.n = n; a.count = n; a.seed = seed; a.step_index = step_index;
  const StepIO<R> io = build_io<R>(a);
  const int block = 256;
  const unsigned grid = unsigned((n + block - 1) / block);
  if (grid == 0) return cudaSuccess;
  eval_scalar_update_kernel<R><<<grid, block, 0, stream>>>(P, io, reinterpret_cast<R*>(param), time, istate, flag,
                                                           reinterpret_cast<R*>(delta));
  return cudaGetLastError();
}

template <typename R>
static cudaError_t launch_eval_draws_t(const LaunchIO& a, int what, int lane, int t, double p, double* out,
                                       cudaStream_t stream) {
  const StepIO<R> io = build_io<R>(a);
  const unsigned grid = unsigned((a.count + 255) / 256);
  if (grid == 0) return cudaSuccess;
  eval_draws_kernel<R><<<grid, 256, 0, stream>>>(io, what, lane, t, p, out);
  return cudaGetLastError();
}

}  // namespace nsg
