// nsgym_classic_launch.cuh -- host launcher for the classic-control kernels, instantiated
// once per precision (nsgym_f32.cu with FMA contraction, nsgym_f64.cu with -fmad=false).
#pragma once
#include <type_traits>

#include "nsgym_device.cuh"
#include "nsgym_host.h"

namespace nsg {

template <typename R, int MAXP>
static ProgramT<R, MAXP> build_program(const NsgymSpec& spec, const DevicePools& pools) {
  ProgramT<R, MAXP> P{};
  P.n_slots = spec.n_slots;
  P.max_steps = spec.max_episode_steps;
  P.autoreset = spec.autoreset;
  P.persistent = spec.persistent_params;
  // block 0 of the Philox stream is consumed by next-step autoreset (initial-state draws), by the
  // normals of slots 0 / 1 and by the gridworld slip draw: compute it once, before any branch
  bool stochastic01 = false;
  for (int j = 0; j < spec.n_slots && j < 2; ++j) {
    const int op = spec.slots[j].upd_op;
    stochastic01 |= (op == NSGYM_UPD_RW || op == NSGYM_UPD_OU || op == NSGYM_UPD_BRW);
  }
  P.rng_prefetch = (spec.autoreset == NSGYM_AUTORESET_NEXT_STEP && !is_grid_kind(spec.env_kind)) || stochastic01 ||
                   is_grid_kind(spec.env_kind);
  for (int i = 0; i < NSGYM_MAX_THETA; ++i) P.theta_default[i] = R(spec.theta_init[i][0]);
  for (int j = 0; j < spec.n_slots && j < MAXP; ++j) {
    const NsgymSlot& a = spec.slots[j];
    SlotT<R>& b = P.slot[j];
    b.sched_op = a.sched_op; b.upd_op = a.upd_op; b.theta_index = a.theta_index; b.constraint = a.constraint;
    b.start = a.start; b.end = a.end;
    for (int k = 0; k < 4; ++k) { b.si[k] = a.si[k]; b.ui[k] = a.ui[k]; }
    b.partner_slot = a.partner_slot; b.partner_index = a.partner_index;
    b.istate_plane = a.istate_plane; b.istate_init = a.istate_init;
    b.sf[0] = a.sf[0]; b.sf[1] = a.sf[1];
    for (int k = 0; k < 6; ++k) b.uf[k] = R(a.uf[k]);
  }
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
  io.n = a.n; io.begin = a.begin; io.count = a.count;
  io.gid_offset = a.gid_offset; io.seed = a.seed; io.step_index = a.step_index;
  io.skip_updates = a.skip_updates; io.force_init = a.force_init;
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
  const int block = 256;
  const unsigned grid = unsigned((n + block - 1) / block);
  if (grid == 0) return cudaSuccess;
  eval_scalar_update_kernel<R, 8><<<grid, block, 0, stream>>>(P, slot, reinterpret_cast<R*>(param), time, istate,
                                                              flag, reinterpret_cast<R*>(delta), inj_u, inj_z, n,
                                                              seed, step_index);
  return cudaGetLastError();
}

}  // namespace nsg
