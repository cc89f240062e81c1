// Gridworld kernels (FrozenLake / CliffWalking / Bridge).  Probabilities are always fp64 and
// this unit is built with -fmad=false: cumulative sums and W1 distances match NumPy exactly.
#include "nsgym_grid.cuh"
#include "nsgym_classic_launch.cuh"

namespace nsg {

template <int MAXP>
static GridProgram<MAXP> build_grid_program(const NsgymSpec& spec, const DevicePools& pools) {
  GridProgram<MAXP> G{};
  // dict-order program first, then permuted so that slot index == theta index
  const ProgramT<double, 8> dict = build_program<double, 8>(spec, pools);
  G.base.n_slots = dict.n_slots; G.base.max_steps = dict.max_steps; G.base.autoreset = dict.autoreset;
  G.base.persistent = dict.persistent; G.base.rng_prefetch = dict.rng_prefetch; G.base.has_istate = dict.has_istate;
  G.base.pool_f = dict.pool_f; G.base.pool_i = dict.pool_i; G.base.bitmap = dict.bitmap;
  for (int i = 0; i < 4; ++i) { G.bound[i] = 0; G.plane[i] = 0; }
  for (int i = 0; i < MAXP; ++i) G.base.slot[i].istate_plane = -1;
  for (int j = 0; j < dict.n_slots; ++j) {
    const int idx = dict.slot[j].theta_index;
    if (idx < 0 || idx >= MAXP) continue;
    G.base.slot[idx] = dict.slot[j];
    G.bound[idx] = 1;
    G.plane[idx] = j;
  }
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < NSGYM_MAX_DIST; ++k) G.dist_init[i][k] = spec.theta_init[i][k];
  G.hole_mask = spec.hole_mask; G.goal_mask = spec.goal_mask; G.start_mask = spec.start_mask;
  G.nrow = spec.nrow; G.ncol = spec.ncol;
  G.inv_ncol = (65536 + spec.ncol - 1) / spec.ncol;
  G.start_cell = spec.start_cell;
  G.n_dist = spec.n_dist; G.split_mode = spec.split_mode; G.terminal_cliff = spec.terminal_cliff;
  G.reward_f = spec.reward_f; G.reward_h = spec.reward_h; G.reward_g = spec.reward_g; G.reward_s = spec.reward_s;
  return G;
}

template <int KIND, int D, int MAXP>
static cudaError_t launch_grid_k(LaunchOp op, const NsgymSpec& spec, const DevicePools& pools, const LaunchIO& a,
                                 cudaStream_t stream) {
  if (spec.n_slots > MAXP || spec.n_dist != D) return cudaErrorInvalidValue;
  const GridProgram<MAXP> G = build_grid_program<MAXP>(spec, pools);
  const StepIO<double> io = build_io<double>(a);
  const int block = 256;
  const unsigned grid = unsigned((a.count + block - 1) / block);
  if (grid == 0) return cudaSuccess;
  switch (op) {
    case OP_STEP: grid_step_kernel<KIND, D, MAXP><<<grid, block, 0, stream>>>(G, io); break;
    case OP_RESET: grid_reset_kernel<KIND, D, MAXP><<<grid, block, 0, stream>>>(G, io); break;
    case OP_ROLLOUT:
      grid_rollout_kernel<KIND, D, MAXP><<<grid, block, 0, stream>>>(G, io, a.k_steps, a.gamma, a.ret, a.len);
      break;
  }
  return cudaGetLastError();
}

cudaError_t launch_grid(LaunchOp op, const NsgymSpec& spec, const DevicePools& pools, const LaunchIO& io,
                        cudaStream_t stream) {
  switch (spec.env_kind) {
    case NSGYM_ENV_FROZENLAKE: return launch_grid_k<NSGYM_ENV_FROZENLAKE, 3, 1>(op, spec, pools, io, stream);
    case NSGYM_ENV_CLIFFWALKING: return launch_grid_k<NSGYM_ENV_CLIFFWALKING, 4, 1>(op, spec, pools, io, stream);
    case NSGYM_ENV_BRIDGE:   // uniform mode binds only P (index 0); split mode needs P_left / P_right too
      if (!spec.split_mode) return launch_grid_k<NSGYM_ENV_BRIDGE, 3, 1>(op, spec, pools, io, stream);
      return launch_grid_k<NSGYM_ENV_BRIDGE, 3, 3>(op, spec, pools, io, stream);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_eval_dist(const NsgymSpec& spec, const DevicePools& pools, int slot, double* param,
                             const int32_t* time, int32_t* istate, uint8_t* flag, double* delta,
                             const double* inj_u, int64_t n, uint64_t seed, uint64_t step_index,
                             cudaStream_t stream) {
  const GridProgram<3> G = build_grid_program<3>(spec, pools);
  LaunchIO a{};
  a.inj_u = inj_u; a.n = n; a.count = n; a.seed = seed; a.step_index = step_index;
  const StepIO<double> io = build_io<double>(a);
  const int index = spec.slots[slot].theta_index;
  const int block = 256;
  const unsigned grid = unsigned((n + block - 1) / block);
  if (grid == 0) return cudaSuccess;
  if (spec.n_dist == 4)
    eval_dist_update_kernel<4, 3><<<grid, block, 0, stream>>>(G, io, index, param, time, istate, flag, delta);
  else
    eval_dist_update_kernel<3, 3><<<grid, block, 0, stream>>>(G, io, index, param, time, istate, flag, delta);
  return cudaGetLastError();
}

}  // namespace nsg
