// Gridworld kernels (FrozenLake / CliffWalking / Bridge).  Probabilities are always fp64 and
// this unit is built with -fmad=false: cumulative sums and W1 distances match NumPy exactly.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "nsgym_grid.cuh"
#include "nsgym_classic_launch.cuh"

namespace nsg {

static_assert(kGridTabWords == GRID_TAB_TOTAL_WORDS, "nsgym_host.h: kGridTabWords");

static bool is_start_cell(const NsgymSpec& spec, int c) {
  if (spec.cell_class) return spec.cell_class[c] == NSGYM_CELL_START;
  return c < 64 && ((spec.start_mask >> c) & 1ull);
}
int resolve_start_cells(NsgymSpec* spec) {
  if (spec->start_cell >= 0) return 1;
  int n = 0, first = 0;
  for (int c = spec->nrow * spec->ncol - 1; c >= 0; --c)
    if (is_start_cell(*spec, c)) { ++n; first = c; }
  spec->start_cell = first;
  return n > 0 ? n : 1;
}

template <int MAXP>
static GridProgram<MAXP> build_grid_program(const NsgymSpec& spec, const DevicePools& pools) {
  GridProgram<MAXP> G{};
  G.base = build_program_by_index<MAXP>(spec, pools);   // slot index == theta index
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < NSGYM_MAX_DIST; ++k) G.dist_init[i][k] = spec.theta_init[i][k];
  G.tab = pools.grid_tab;
  G.nrow = spec.nrow; G.ncol = spec.ncol;
  G.n_cells = spec.nrow * spec.ncol;
  // several start cells (FrozenLake): a reset samples one; the list is in the map table (resolve_start_cells)
  G.n_start = pools.grid_n_start > 1 ? pools.grid_n_start : 1;
  G.start_cell = spec.start_cell;
  G.n_dist = spec.n_dist; G.split_mode = spec.split_mode; G.terminal_cliff = spec.terminal_cliff;
  G.reward_f = spec.reward_f; G.reward_h = spec.reward_h; G.reward_g = spec.reward_g; G.reward_s = spec.reward_s;
  return G;
}

// the grid program indexes its slots by theta index; the row table by dict position
template <int MAXP>
static HetT<double, MAXP> build_het_by_index(const NsgymSpec& spec, const RowTable& t) {
  HetT<double, MAXP> H{};
  H.ints = t.d_int;
  H.reals = reinterpret_cast<const double*>(t.d_real);
  H.dbls = t.d_dbl;
  for (int j = 0; j < spec.n_slots; ++j) {
    const int q = spec.slots[j].theta_index;
    if (q < 0 || q >= MAXP) continue;
    H.mask[q] = t.mask[j];
    for (int w = 0; w < kRowWords; ++w) H.plane[q][w] = t.plane[j][w];
    for (int w = 0; w < kRowInt; ++w) { H.idef[q][w] = t.def_int[j][w]; H.shift[q][w] = t.shift[j][w]; H.bits[q][w] = t.bits[j][w]; }
    for (int w = 0; w < kRowReal; ++w) H.rdef[q][w] = t.def_real[j][w];
    for (int w = 0; w < kRowDbl; ++w) H.ddef[q][w] = t.def_dbl[j][w];
  }
  return H;
}

// the pointer-free parts of a GridProgram as a constexpr function of the specialised sources
template <int MAXP>
static std::string spec_grid_program_source(const GridProgram<MAXP>& G) {
  const std::string prog = "GridProgram<" + std::to_string(MAXP) + ">";
  const std::string headt = "ProgramHeadT<double, " + std::to_string(MAXP) + ">";
  const GridConsts& consts = G;
  const ProgramHeadT<double, MAXP>& head = G.base;
  return "__device__ constexpr " + prog + " spec_program() {\n  " + prog + " G{};\n" +
         spec_assign("static_cast<GridConsts&>(G)", "GridConsts", consts) +
         spec_assign("static_cast<" + headt + "&>(G.base)", headt, head) + "  return G;\n}\n}  // namespace nsg\n";
}

// The single-step kernel of a lean gridworld program (deterministic schedulers and rules): grid_step_body
// with everything but the four device pointers of the program as a compile-time constant.
template <int KIND, int D, int MAXP>
static std::string spec_grid_step_source(const GridProgram<MAXP>& G, const StepIO<double>& io, bool root, bool slow = false) {
  const std::string prog = "GridProgram<" + std::to_string(MAXP) + ">";
  std::string s = spec_prelude<double>("nsgym_grid.cuh", io, root) + spec_grid_program_source<MAXP>(G);
  // (slow: stochastic schedulers / Dirichlet draws / the Lipschitz loop -- the general class, 4 resident blocks)
  s += "extern \"C\" __global__ void __launch_bounds__(256, " +
       (slow ? std::string("4") : "nsg::grid_spec_min_blocks<" + std::to_string(KIND) + ">()") +
       ")\nnsgym_spec_grid_step(const __grid_constant__ nsg::StepIO<double> io, const __grid_constant__ nsg::GridPtrs ptrs) {\n"
       "  constexpr nsg::" + prog + " G0 = nsg::spec_program();\n  nsg::" + prog + " G = G0;\n"
       "  G.base.pool_f = ptrs.pool_f; G.base.pool_i = ptrs.pool_i; G.base.bitmap = ptrs.bitmap; G.tab = ptrs.tab;\n"
       "  nsg::grid_step_body<" + std::to_string(KIND) + ", " + std::to_string(D) + ", " + std::to_string(MAXP) +
       ", " + (slow ? "true" : "false") + ", nsg::SpecFix>(G, io);\n}\n";
  return s;
}

// the tiled single-step kernel (grid_step_body_tiled): TMA bulk copies stream the planes of the next tiles
// into shared memory while the block advances the current one
template <int KIND, int D, int MAXP>
static std::string spec_grid_step_tiled_source(const GridProgram<MAXP>& G, const StepIO<double>& io, bool root) {
  const std::string prog = "GridProgram<" + std::to_string(MAXP) + ">";
  std::string s = spec_prelude<double>("nsgym_grid.cuh", io, root) + spec_grid_program_source<MAXP>(G);
  s += "extern \"C\" __global__ void __launch_bounds__(256, nsg::grid_tiled_min_blocks<" + std::to_string(KIND) +
       ">())\nnsgym_spec_grid_step_tiled(const __grid_constant__ nsg::StepIO<double> io, const __grid_constant__ nsg::GridPtrs ptrs, "
       "int tiles_per_block) {\n"
       "  constexpr nsg::" + prog + " G0 = nsg::spec_program();\n  nsg::" + prog + " G = G0;\n"
       "  G.base.pool_f = ptrs.pool_f; G.base.pool_i = ptrs.pool_i; G.base.bitmap = ptrs.bitmap; G.tab = ptrs.tab;\n"
       "  nsg::grid_step_body_tiled<" + std::to_string(KIND) + ", " + std::to_string(D) + ", " + std::to_string(MAXP) +
       ", nsg::SpecFix>(G, io, tiles_per_block);\n}\n";
  return s;
}

// single step of a gridworld batch with lean per-env rows (grid_step_het_body)
template <int KIND, int D, int MAXP>
static std::string spec_grid_step_rows_source(const GridProgram<MAXP>& G, const HetT<double, MAXP>& H,
                                              const StepIO<double>& io, bool root, bool lean = true) {
  const std::string prog = "GridProgram<" + std::to_string(MAXP) + ">";
  std::string s = spec_prelude<double>("nsgym_grid.cuh", io, root) + spec_grid_program_source<MAXP>(G) +
                  spec_rows_source<double, MAXP>(H);
  s += std::string("extern \"C\" __global__ void __launch_bounds__(256, ") + (lean ? "NSGYM_HET_LEAN_MIN_BLOCKS" : "NSGYM_HET_MIN_BLOCKS") +
       ")\nnsgym_spec_grid_step_rows("
       "const __grid_constant__ nsg::StepIO<double> io, const __grid_constant__ nsg::GridPtrs ptrs, "
       "const __grid_constant__ nsg::HetPtrs hp) {\n"
       "  constexpr nsg::" + prog + " G0 = nsg::spec_program();\n  nsg::" + prog + " G = G0;\n"
       "  G.base.pool_f = ptrs.pool_f; G.base.pool_i = ptrs.pool_i; G.base.bitmap = ptrs.bitmap; G.tab = ptrs.tab;\n" +
       spec_rows_object("double", MAXP) +
       "  nsg::grid_step_het_body<" + std::to_string(KIND) + ", " + std::to_string(D) + ", " + std::to_string(MAXP) +
       ", " + (lean ? "true" : "false") + ", nsg::SpecFix>(G, H, io);\n}\n";
  return s;
}

// K fused steps of a lean gridworld program under the uniform-random policy (grid_rollout_body)
template <int KIND, int D, int MAXP>
static std::string spec_grid_rollout_source(const GridProgram<MAXP>& G, const StepIO<double>& io, bool root, bool slow = false,
                                            bool tab = false) {
  const std::string prog = "GridProgram<" + std::to_string(MAXP) + ">";
  std::string s = spec_prelude<double>("nsgym_grid.cuh", io, root) + spec_grid_program_source<MAXP>(G);
  s += "extern \"C\" __global__ void __launch_bounds__(256)\nnsgym_spec_grid_rollout(const __grid_constant__ nsg::StepIO<double> io, "
       "const __grid_constant__ nsg::GridPtrs ptrs, const __grid_constant__ nsg::RolloutArgs ra) {\n"
       "  constexpr nsg::" + prog + " G0 = nsg::spec_program();\n  nsg::" + prog + " G = G0;\n"
       "  G.base.pool_f = ptrs.pool_f; G.base.pool_i = ptrs.pool_i; G.base.bitmap = ptrs.bitmap; G.tab = ptrs.tab;\n"
       "  const nsg::HetT<double, " + std::to_string(MAXP) + "> no_rows{};\n"
       "  nsg::grid_rollout_body<" + std::to_string(KIND) + ", " + std::to_string(D) + ", " + std::to_string(MAXP) +
       ", " + (slow ? "true" : "false") + ", false, " + (tab ? "true" : "false") +
       ", nsg::SpecFix>(G, no_rows, io, ra.k_steps, ra.gamma, ra.ret, ra.len, static_cast<const uint8_t*>(ra.pol), ra.pol_per_env);\n}\n";
  return s;
}

template <int KIND, int D, int MAXP>
static cudaError_t launch_grid_k(LaunchOp op, const NsgymSpec& spec, const DevicePools& pools, const LaunchIO& a,
                                 cudaStream_t stream) {
  if (spec.n_slots > MAXP || spec.n_dist != D) return cudaErrorInvalidValue;
  const GridProgram<MAXP> G = build_grid_program<MAXP>(spec, pools);
  // lean instantiation when every bound rule is deterministic
  bool slow = a.general_kernels != 0 || a.inj_u != nullptr;     // the lean kernels fold the injection tests away
  for (int j = 0; j < spec.n_slots; ++j) {
    const NsgymSlot& sl = spec.slots[j];
    slow |= sl.sched_op == NSGYM_SCHED_RANDOM || sl.sched_op == NSGYM_SCHED_DECAY ||
            sl.sched_op == NSGYM_SCHED_MEMORYLESS || sl.upd_op == NSGYM_UPD_D_RANDOM || sl.ui[2] != 0;
  }
  const bool het = a.rows && a.rows->active;
  LaunchIO a2 = a;
  a2.prefetch = 1;       // the step pair's Philox block before any divergent branch (general kernels; the lean step kernels draw lazily)
  const StepIO<double> io = build_io<double>(a2);
  const int block = 256;
  const unsigned grid = unsigned((a.count + block - 1) / block);
  if (grid == 0) return cudaSuccess;
  if (a.kernel_class) {
    if (het) *a.kernel_class = (a.rows->lean && !a.general_kernels && !a.inj_u) ? NSGYM_KERNEL_ROWS_LEAN : NSGYM_KERNEL_ROWS_GENERAL;
    else *a.kernel_class = slow ? NSGYM_KERNEL_GENERAL : NSGYM_KERNEL_LEAN_FAST;
  }
  if (het) {
    if (a.spec_source) return cudaErrorNotSupported;
    const HetT<double, MAXP> H = build_het_by_index<MAXP>(spec, *a.rows);
    if (a.specialized) *a.specialized = 0;
    // rows of the general class (stochastic schedulers per env) specialise too: row layout constant, row loads up
    // front, no injection code
    if (op == OP_STEP && !a.general_kernels && !a.inj_u && a.specialize) {
      const bool root = a.plan_elapsed < 0 && !a.skip_updates;
      const bool lean_rows = a.rows->lean;
      const uint32_t facts = spec_facts(io, root) | 256u | (lean_rows ? 0u : 2048u);
      cudaKernel_t k = nullptr;
      if (!a.spec_cache || !a.spec_cache->find(facts, &k)) {
        k = jit::kernel(spec_grid_step_rows_source<KIND, D, MAXP>(G, H, io, root, lean_rows), "nsgym_spec_grid_step_rows", false, nullptr);
        if (a.spec_cache) a.spec_cache->put(facts, k);
      }
      if (k) {
        GridPtrs ptrs{G.base.pool_f, G.base.pool_i, G.base.bitmap, G.tab};
        HetPtrs hp{H.ints, H.reals, H.dbls};
        void* args[] = {const_cast<StepIO<double>*>(&io), &ptrs, &hp};
        if (a.specialized) *a.specialized = 1;
        return cudaLaunchKernel(reinterpret_cast<const void*>(k), dim3(grid), dim3(block), args, 0, stream);
      }
    }
    switch (op) {
      case OP_STEP:
        if (a.rows->lean && !a.general_kernels && !a.inj_u) grid_step_het_kernel<KIND, D, MAXP, true><<<grid, block, 0, stream>>>(G, H, io);
        else grid_step_het_kernel<KIND, D, MAXP, false><<<grid, block, 0, stream>>>(G, H, io);
        break;
      case OP_RESET: grid_reset_het_kernel<KIND, D, MAXP><<<grid, block, 0, stream>>>(G, H, io); break;
      case OP_ROLLOUT:
        if (a.policy)
          grid_rollout_kernel<KIND, D, MAXP, true, true, true><<<grid, block, 0, stream>>>(
              G, H, io, a.k_steps, a.gamma, a.ret, a.len, static_cast<const uint8_t*>(a.policy), a.policy_per_env);
        else
          grid_rollout_kernel<KIND, D, MAXP, true, true><<<grid, block, 0, stream>>>(G, H, io, a.k_steps, a.gamma,
                                                                                    a.ret, a.len);
        break;
    }
    return cudaGetLastError();
  }
  const HetT<double, MAXP> no_rows{};
  if (a.specialized) *a.specialized = 0;
  // programs with stochastic rules specialise as well (general class: every scheduler / rule switch folds to
  // the slot's own); injected tables and NSGYM_OPT_GENERAL_KERNELS keep the precompiled general kernel
  const bool spec_ok = !slow || (!a.inj_u && !a.general_kernels);
  if (op == OP_STEP && spec_ok && (a.specialize || a.spec_source)) {
    const bool root = a.plan_elapsed < 0 && !a.skip_updates;
    const uint32_t facts = spec_facts(io, root) | (slow ? 1024u : 0u);
    cudaKernel_t k = nullptr;
    if (a.spec_source || !a.spec_cache || !a.spec_cache->find(facts, &k)) {
      const std::string src = spec_grid_step_source<KIND, D, MAXP>(G, io, root, slow);
      if (a.spec_source) { *a.spec_source = src; return cudaSuccess; }
      k = jit::kernel(src, "nsgym_spec_grid_step", false, nullptr);
      if (a.spec_cache) a.spec_cache->put(facts, k);
    }
    if (k) {
      GridPtrs ptrs{G.base.pool_f, G.base.pool_i, G.base.bitmap, G.tab};
      if (a.specialized) *a.specialized = 1;
      // full tiles of 256 envs through the tiled kernel (TMA-prefetched planes) when every plane is 16-byte
      // aligned; the remainder, and everything else, through the one-thread-one-env kernel
      StepIO<double> rest = io;
      const uint32_t full_tiles = tiled_ok(io, MAXP == 1 && !slow) ? io.count / 256u : 0u;
      if (full_tiles) {
        cudaKernel_t kt = nullptr;
        const uint32_t tfacts = facts | 512u;
        if (!a.spec_cache || !a.spec_cache->find(tfacts, &kt)) {
          kt = jit::kernel(spec_grid_step_tiled_source<KIND, D, MAXP>(G, io, root), "nsgym_spec_grid_step_tiled", false, nullptr);
          if (a.spec_cache) a.spec_cache->put(tfacts, kt);
        }
        if (kt) {
          StepIO<double> head = io;
          head.count = full_tiles * 256u;
          int tiles_per_block = tiles_per_block_for(full_tiles);
          void* targs[] = {&head, &ptrs, &tiles_per_block};
          const unsigned tgrid = (full_tiles + unsigned(tiles_per_block) - 1) / unsigned(tiles_per_block);
          cudaError_t e = cudaLaunchKernel(reinterpret_cast<const void*>(kt), dim3(tgrid), dim3(256), targs, 0, stream);
          if (e != cudaSuccess) return e;
          rest.begin += head.count;
          rest.count -= head.count;
          if (rest.count == 0) return cudaSuccess;
        }
      }
      void* args[] = {&rest, &ptrs};
      return cudaLaunchKernel(reinterpret_cast<const void*>(k), dim3((rest.count + block - 1) / block), dim3(block), args, 0, stream);
    }
  }
  if (op == OP_ROLLOUT && spec_ok && (a.specialize || a.spec_source)) {
    const bool root = a.plan_elapsed < 0 && !a.skip_updates;
    const bool tab = a.policy != nullptr;            // tabular policy (nsgym_rollout_linear on a gridworld)
    const uint32_t facts = spec_facts(io, root) | 64u | (slow ? 1024u : 0u) | (tab ? 128u : 0u);
    cudaKernel_t k = nullptr;
    if (a.spec_source || !a.spec_cache || !a.spec_cache->find(facts, &k)) {
      const std::string src = spec_grid_rollout_source<KIND, D, MAXP>(G, io, root, slow, tab);
      if (a.spec_source) { *a.spec_source = src; return cudaSuccess; }
      k = jit::kernel(src, "nsgym_spec_grid_rollout", false, nullptr);
      if (a.spec_cache) a.spec_cache->put(facts, k);
    }
    if (k) {
      GridPtrs ptrs{G.base.pool_f, G.base.pool_i, G.base.bitmap, G.tab};
      RolloutArgs ra{a.k_steps, a.gamma, a.ret, a.len, a.policy, a.policy_per_env};
      void* args[] = {const_cast<StepIO<double>*>(&io), &ptrs, &ra};
      if (a.specialized) *a.specialized = 1;
      return cudaLaunchKernel(reinterpret_cast<const void*>(k), dim3(grid), dim3(block), args, 0, stream);
    }
  }
  if (a.spec_source) return cudaErrorNotSupported;
  switch (op) {
    case OP_STEP:
      if (slow) grid_step_kernel<KIND, D, MAXP, true><<<grid, block, 0, stream>>>(G, io);
      else grid_step_kernel<KIND, D, MAXP, false><<<grid, block, 0, stream>>>(G, io);
      break;
    case OP_RESET: grid_reset_kernel<KIND, D, MAXP><<<grid, block, 0, stream>>>(G, io); break;
    case OP_ROLLOUT:
      if (a.policy)       // tabular policy: the general instantiation
        grid_rollout_kernel<KIND, D, MAXP, true, false, true><<<grid, block, 0, stream>>>(
            G, no_rows, io, a.k_steps, a.gamma, a.ret, a.len, static_cast<const uint8_t*>(a.policy), a.policy_per_env);
      else if (slow) grid_rollout_kernel<KIND, D, MAXP, true><<<grid, block, 0, stream>>>(G, no_rows, io, a.k_steps, a.gamma, a.ret, a.len);
      else grid_rollout_kernel<KIND, D, MAXP, false><<<grid, block, 0, stream>>>(G, no_rows, io, a.k_steps, a.gamma, a.ret, a.len);
      break;
  }
  return cudaGetLastError();
}

template <int KIND, int D, int MAXP>
static cudaError_t launch_table_k(const NsgymSpec& spec, const DevicePools& pools, const LaunchIO& a, int64_t env,
                                  int n_times, double* prob, int32_t* next, float* reward, uint8_t* done,
                                  cudaStream_t stream) {
  if (spec.n_slots > MAXP || spec.n_dist != D) return cudaErrorInvalidValue;
  const GridProgram<MAXP> G = build_grid_program<MAXP>(spec, pools);
  const StepIO<double> io = build_io<double>(a);
  double* traj = nullptr;
  cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&traj), sizeof(double) * size_t(n_times) * MAXP * D, stream);
  if (e != cudaSuccess) return e;
  HetT<double, MAXP> H{};
  if (a.rows && a.rows->active) {
    H = build_het_by_index<MAXP>(spec, *a.rows);
    grid_trajectory_kernel<KIND, D, MAXP, true><<<1, 32, 0, stream>>>(G, H, io, uint32_t(env), n_times, traj);
  } else {
    grid_trajectory_kernel<KIND, D, MAXP, false><<<1, 32, 0, stream>>>(G, H, io, uint32_t(env), n_times, traj);
  }
  const uint64_t total = uint64_t(n_times) * spec.nrow * spec.ncol * 4;
  grid_table_kernel<KIND, D, MAXP><<<unsigned((total + 255) / 256), 256, 0, stream>>>(G, traj, n_times, prob, next,
                                                                                      reward, done);
  e = cudaGetLastError();
  cudaFreeAsync(traj, stream);
  return e;
}

cudaError_t launch_table(const NsgymSpec& spec, const DevicePools& pools, const LaunchIO& io, int64_t env, int n_times,
                         double* prob, int32_t* next, float* reward, uint8_t* done, cudaStream_t stream) {
  switch (spec.env_kind) {
    case NSGYM_ENV_FROZENLAKE:
      return launch_table_k<NSGYM_ENV_FROZENLAKE, 3, 1>(spec, pools, io, env, n_times, prob, next, reward, done, stream);
    case NSGYM_ENV_CLIFFWALKING:
      return launch_table_k<NSGYM_ENV_CLIFFWALKING, 4, 1>(spec, pools, io, env, n_times, prob, next, reward, done, stream);
    case NSGYM_ENV_BRIDGE:
      if (!spec.split_mode)
        return launch_table_k<NSGYM_ENV_BRIDGE, 3, 1>(spec, pools, io, env, n_times, prob, next, reward, done, stream);
      return launch_table_k<NSGYM_ENV_BRIDGE, 3, 3>(spec, pools, io, env, n_times, prob, next, reward, done, stream);
    default: return cudaErrorInvalidValue;
  }
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
  GridProgram<1> G = build_grid_program<1>(spec, pools);
  G.base.slot[0] = lower_slot<double>(spec.slots[slot], slot);
  G.base.bound_mask = 1;
  LaunchIO a{};
  a.inj_u = inj_u; a.n = n; a.count = n; a.seed = seed; a.step_index = step_index;
  const StepIO<double> io = build_io<double>(a);
  const int block = 256;
  const unsigned grid = unsigned((n + block - 1) / block);
  if (grid == 0) return cudaSuccess;
  if (spec.n_dist == 4)
    eval_dist_update_kernel<4><<<grid, block, 0, stream>>>(G, io, param, time, istate, flag, delta);
  else
    eval_dist_update_kernel<3><<<grid, block, 0, stream>>>(G, io, param, time, istate, flag, delta);
  return cudaGetLastError();
}

int build_grid_tables(const NsgymSpec& spec, uint32_t* words, char* err, size_t err_len) {
  const int nrow = spec.nrow, ncol = spec.ncol, n = nrow * ncol;
  if (nrow <= 0 || ncol <= 0 || n > GRID_MAX_CELLS) {
    snprintf(err, err_len, "gridworld maps are limited to %d cells (got %dx%d)", GRID_MAX_CELLS, nrow, ncol);
    return -1;
  }
  if (spec.cell_class ? spec.n_cell_class != n : n > 64) {
    snprintf(err, err_len, spec.cell_class ? "n_cell_class must equal nrow * ncol"
                                           : "maps of more than 64 cells need cell_class (the masks hold 64 bits)");
    return -1;
  }
  uint8_t bytes[GRID_TAB_WORDS * 4] = {};
  const bool cliff = spec.env_kind == NSGYM_ENV_CLIFFWALKING;
  for (int c = 0; c < n; ++c) {
    const int row = c / ncol, col = c % ncol;
    for (int b = 0; b < 4; ++b) {
      // CliffWalking: UP RIGHT DOWN LEFT (toy_text.py:74-76); FrozenLake / Bridge: LEFT DOWN RIGHT UP
      // (toy_text.py:321-324, envs/Bridge.py:14-17); the clamp is "out of bounds -> stay" for unit moves
      const int dr = cliff ? (b == 2) - (b == 0) : (b == 1) - (b == 3);
      const int dc = cliff ? (b == 1) - (b == 3) : (b == 2) - (b == 0);
      const int r2 = row + dr < 0 ? 0 : (row + dr > nrow - 1 ? nrow - 1 : row + dr);
      const int c2 = col + dc < 0 ? 0 : (col + dc > ncol - 1 ? ncol - 1 : col + dc);
      bytes[c * 4 + b] = uint8_t(r2 * ncol + c2);
    }
    uint32_t cls = 0;
    if (spec.cell_class) {
      const uint8_t letter = spec.cell_class[c];
      if (letter > NSGYM_CELL_START) { snprintf(err, err_len, "cell_class[%d] = %d is not a NSGYM_CELL_* letter", c, letter); return -1; }
      cls = letter == NSGYM_CELL_HOLE ? CELL_HOLE : letter == NSGYM_CELL_GOAL ? CELL_GOAL : letter == NSGYM_CELL_START ? CELL_START : 0;
    } else {
      const uint64_t bit = 1ull << c;
      cls = ((spec.hole_mask & bit) ? CELL_HOLE : 0) | ((spec.goal_mask & bit) ? CELL_GOAL : 0) |
            ((spec.start_mask & bit) ? CELL_START : 0);
    }
    if (col < (ncol >> 1)) cls |= CELL_LEFT;                     // envs/Bridge.py:148-157
    bytes[4 * n + c] = uint8_t(cls);
  }
  std::memcpy(words, bytes, sizeof bytes);
  for (uint32_t c = 0; c < 8; ++c) {        // reward of the destination letter: hole > goal > start > frozen
    const float r = (c & CELL_HOLE) ? spec.reward_h : (c & CELL_GOAL) ? spec.reward_g : (c & CELL_START) ? spec.reward_s : spec.reward_f;
    std::memcpy(words + GRID_TAB_WORDS + c, &r, 4);
  }
  uint8_t starts[GRID_MAX_CELLS] = {};       // the start cells in row-major order (categorical_sample's cumsum order)
  int n_start = 0;
  for (int c = 0; c < n; ++c)
    if (is_start_cell(spec, c)) starts[n_start++] = uint8_t(c);
  std::memcpy(words + GRID_TAB_START_WORD, starts, sizeof starts);
  return GRID_TAB_TOTAL_WORDS;
}

cudaError_t launch_eval_draws_grid(const LaunchIO& a, int n_dist, int what, int lane, int t, double p, double* out,
                                   cudaStream_t stream) {
  const StepIO<double> io = build_io<double>(a);
  const unsigned grid = unsigned((a.count + 255) / 256);
  if (grid == 0) return cudaSuccess;
  if (n_dist == 4) eval_grid_draws_kernel<4><<<grid, 256, 0, stream>>>(io, what, lane, t, p, out);
  else eval_grid_draws_kernel<3><<<grid, 256, 0, stream>>>(io, what, lane, t, p, out);
  return cudaGetLastError();
}

cudaError_t launch_eval_w1(int dim, const double* u, const double* v, double* out, double* ref, int64_t n,
                           cudaStream_t stream) {
  const unsigned grid = unsigned((n + 255) / 256);
  if (grid == 0) return cudaSuccess;
  if (dim == 3) eval_w1_kernel<3><<<grid, 256, 0, stream>>>(u, v, out, ref, uint32_t(n));
  else if (dim == 4) eval_w1_kernel<4><<<grid, 256, 0, stream>>>(u, v, out, ref, uint32_t(n));
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

}  // namespace nsg
