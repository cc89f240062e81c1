// nsgym_classic_launch.cuh -- host launcher for the classic-control kernels, instantiated
// once per precision (nsgym_f32.cu with FMA contraction, nsgym_f64.cu with -fmad=false).
#pragma once
#include <limits>
#include <type_traits>

#include "nsgym_device.cuh"
#include "nsgym_host.h"

namespace nsg {

inline bool is_fast_affine(int op) {
  return op == NSGYM_UPD_NOP || op == NSGYM_UPD_ADD || op == NSGYM_UPD_ADD_T || op == NSGYM_UPD_MUL ||
         op == NSGYM_UPD_RW;
}

template <typename R> struct TrueMin;
template <> struct TrueMin<float> { static constexpr float value = 1.401298464324817e-45f; };
template <> struct TrueMin<double> { static constexpr double value = 4.9406564584124654e-324; };

template <typename R, int MAXP>
static ProgramT<R, MAXP> build_program(const NsgymSpec& spec, const DevicePools& pools) {
  ProgramT<R, MAXP> P{};
  P.n_slots = spec.n_slots < MAXP ? spec.n_slots : MAXP;
  P.max_steps = spec.max_episode_steps;
  P.autoreset = spec.autoreset;
  P.persistent = spec.persistent_params;
  for (int i = 0; i < NSGYM_MAX_THETA; ++i) P.theta_default[i] = R(spec.theta_init[i][0]);
  bool stochastic01 = false;
  for (int j = 0; j < P.n_slots; ++j) {
    const NsgymSlot& a = spec.slots[j];
    SlotT<R>& b = P.slot[j];
    b.theta_index = a.theta_index;
    b.sched_op = a.sched_op; b.upd_op = a.upd_op; b.constraint = a.constraint;
    b.start = a.start; b.end = a.end;
    b.gated = (a.start > 0 || a.end < (1 << 28)) ? 1 : 0;
    for (int k = 0; k < 4; ++k) { b.si[k] = a.si[k]; b.ui[k] = a.ui[k]; }
    b.partner_slot = a.partner_slot; b.partner_index = a.partner_index;
    b.istate_plane = a.istate_plane; b.istate_init = a.istate_init;
    if (a.istate_plane >= 0) P.has_istate = 1;
    b.sf[0] = a.sf[0]; b.sf[1] = a.sf[1];
    for (int k = 0; k < 6; ++k) b.uf[k] = R(a.uf[k]);
    b.fast = is_fast_affine(a.upd_op) ? 1 : 0;
    // ((A y + B) + noise) + C t
    double A = 1.0, B = 0.0, Ct = 0.0;
    switch (a.upd_op) {
      case NSGYM_UPD_ADD: B = a.uf[0]; break;
      case NSGYM_UPD_ADD_T: Ct = a.uf[0]; break;
      case NSGYM_UPD_MUL: A = a.uf[0]; break;
      case NSGYM_UPD_RW: B = a.uf[0]; Ct = a.uf[3]; break;
      default: break;
    }
    b.fa[0] = R(A); b.fa[1] = R(B); b.fa[2] = R(Ct);
    // `v <= 0` -> v <= 0;  `v < 0` -> v <= -(smallest subnormal);  none -> v <= -inf (never)
    if (a.constraint == NSGYM_CONS_REJECT_LE0) b.reject_le = R(0);
    else if (a.constraint == NSGYM_CONS_REJECT_LT0) b.reject_le = -TrueMin<R>::value;
    else b.reject_le = -std::numeric_limits<R>::infinity();
    if (j < 2) stochastic01 |= (a.upd_op == NSGYM_UPD_RW || a.upd_op == NSGYM_UPD_OU || a.upd_op == NSGYM_UPD_BRW);
  }
  // block 0 of the Philox stream is consumed by next-step autoreset (initial-state draws), by the
  // normals of slots 0 / 1 and by the gridworld slip draw: compute it once, before any branch
  P.rng_prefetch = (spec.autoreset == NSGYM_AUTORESET_NEXT_STEP && !is_grid_kind(spec.env_kind)) || stochastic01 ||
                   is_grid_kind(spec.env_kind);
  P.pool_f = pools.pool_f; P.pool_i = pools.pool_i; P.bitmap = pools.bitmap;
  return P;
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
  uint32_t k0 = uint32_t(a.seed), k1 = uint32_t(a.seed >> 32);
  for (int r = 0; r < 10; ++r) {            // Philox4x32 key schedule (Weyl sequence)
    io.rk[r][0] = k0; io.rk[r][1] = k1;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return io;
}

template <typename R, int KIND, int MAXP>
static cudaError_t launch_classic_kmp(LaunchOp op, const NsgymSpec& spec, const DevicePools& pools,
                                      const LaunchIO& a, cudaStream_t stream) {
  const ProgramT<R, MAXP> P = build_program<R, MAXP>(spec, pools);
  const StepIO<R> io = build_io<R>(a);
  const int block = 256;
  const unsigned grid = unsigned((a.count + block - 1) / block);
  if (grid == 0) return cudaSuccess;
  switch (op) {
    case OP_STEP: classic_step_kernel<R, KIND, MAXP><<<grid, block, 0, stream>>>(P, io); break;
    case OP_RESET: classic_reset_kernel<R, KIND, MAXP><<<grid, block, 0, stream>>>(P, io); break;
    case OP_ROLLOUT:
      classic_rollout_kernel<R, KIND, MAXP><<<grid, block, 0, stream>>>(P, io, a.k_steps, a.gamma, a.ret, a.len);
      break;
  }
  return cudaGetLastError();
}

// slot-count buckets: registers and unrolled interpreter iterations scale with MAXP
template <typename R, int KIND>
static cudaError_t launch_classic_k(LaunchOp op, const NsgymSpec& spec, const DevicePools& pools,
                                    const LaunchIO& a, cudaStream_t stream) {
  constexpr int NTH = KindTraits<KIND>::NTH;
  const int p = spec.n_slots;
  if (p > NTH) return cudaErrorInvalidValue;
  if constexpr (NTH <= 2) {
    return launch_classic_kmp<R, KIND, 2>(op, spec, pools, a, stream);
  } else if constexpr (NTH <= 4) {
    if (p <= 2) return launch_classic_kmp<R, KIND, 2>(op, spec, pools, a, stream);
    return launch_classic_kmp<R, KIND, 4>(op, spec, pools, a, stream);
  } else {
    if (p <= 2) return launch_classic_kmp<R, KIND, 2>(op, spec, pools, a, stream);
    if (p <= 4) return launch_classic_kmp<R, KIND, 4>(op, spec, pools, a, stream);
    return launch_classic_kmp<R, KIND, 8>(op, spec, pools, a, stream);
  }
}

template <typename R>
static cudaError_t launch_classic_t(LaunchOp op, const NsgymSpec& spec, const DevicePools& pools,
                                    const LaunchIO& a, cudaStream_t stream) {
  switch (spec.env_kind) {
    case NSGYM_ENV_CARTPOLE: return launch_classic_k<R, NSGYM_ENV_CARTPOLE>(op, spec, pools, a, stream);
    case NSGYM_ENV_ACROBOT: return launch_classic_k<R, NSGYM_ENV_ACROBOT>(op, spec, pools, a, stream);
    case NSGYM_ENV_MOUNTAINCAR: return launch_classic_k<R, NSGYM_ENV_MOUNTAINCAR>(op, spec, pools, a, stream);
    case NSGYM_ENV_MOUNTAINCAR_CONT:
      return launch_classic_k<R, NSGYM_ENV_MOUNTAINCAR_CONT>(op, spec, pools, a, stream);
    case NSGYM_ENV_PENDULUM: return launch_classic_k<R, NSGYM_ENV_PENDULUM>(op, spec, pools, a, stream);
    default: return cudaErrorInvalidValue;
  }
}

template <typename R>
static cudaError_t launch_eval_scalar_t(const NsgymSpec& spec, const DevicePools& pools, int slot, void* param,
                                        const int32_t* time, int32_t* istate, uint8_t* flag, void* delta,
                                        const double* inj_u, const double* inj_z, int64_t n, uint64_t seed,
                                        uint64_t step_index, cudaStream_t stream) {
  const ProgramT<R, 8> P = build_program<R, 8>(spec, pools);
  LaunchIO a{};
  a.inj_u = inj_u; a.inj_z = inj_z; a.n = n; a.count = n; a.seed = seed; a.step_index = step_index;
  const StepIO<R> io = build_io<R>(a);
  const int block = 256;
  const unsigned grid = unsigned((n + block - 1) / block);
  if (grid == 0) return cudaSuccess;
  eval_scalar_update_kernel<R, 8><<<grid, block, 0, stream>>>(P, io, slot,
                                                              reinterpret_cast<R*>(param), time, istate, flag,
                                                              reinterpret_cast<R*>(delta));
  return cudaGetLastError();
}

}  // namespace nsg
