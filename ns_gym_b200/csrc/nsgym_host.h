// nsgym_host.h -- internal host-side declarations shared by the translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <utility>
#include <vector>

#include "nsgym_b200.h"

namespace nsg {

enum LaunchOp { OP_STEP = 0, OP_RESET = 1, OP_ROLLOUT = 2 };

struct DevicePools {
  const double* pool_f;
  const int32_t* pool_i;
  const uint32_t* bitmap;
  const uint32_t* grid_tab;   // gridworld map tables (nsgym_grid.cuh: next[4 n] ++ cls[n]), NULL for classic control
  int32_t grid_n_start;       // > 1: a reset samples one of the map's start cells (listed in grid_tab)
};

// Map tables of a gridworld spec, as the kernels read them (host side; nsgym_create uploads them).
// Returns the number of 32-bit words written to `words` (capacity kGridTabWords), or -1 with `err` set.
constexpr int kGridTabWords = 392;    // = GRID_TAB_TOTAL_WORDS (nsgym_grid.cuh; asserted in nsgym_gridworld.cu)
// A spec with start_cell == -1 (FrozenLake map with several 'S' cells, toy_text.py:314-319 accepts any desc):
// returns the number of start cells and replaces start_cell by the first one; 1 otherwise.  Needs the map
// (masks / cell_class), i.e. runs before nsgym_create drops the host pointers.
int resolve_start_cells(NsgymSpec* spec);
int build_grid_tables(const NsgymSpec& spec, uint32_t* words, char* err, size_t err_len);

// Per-env rows of a heterogeneous handle (nsgym_create_rows): lowered, only the words that vary
// between envs kept, as SoA planes in device memory owned by the handle.
constexpr int kRowInt = 10, kRowReal = 5, kRowDbl = 2, kRowWords = kRowInt + kRowReal + kRowDbl;
struct RowTable {
  bool active = false;
  int precision = NSGYM_F32;                // word size of the real planes
  int32_t* d_int = nullptr;
  void* d_real = nullptr;
  double* d_dbl = nullptr;
  int n_int = 0, n_real = 0, n_dbl = 0;     // plane counts
  uint32_t mask[NSGYM_MAX_SLOTS] = {};
  uint8_t plane[NSGYM_MAX_SLOTS][kRowWords] = {};
  uint8_t shift[NSGYM_MAX_SLOTS][kRowInt] = {};   // int words are bit-packed: field position / width in their plane
  uint8_t bits[NSGYM_MAX_SLOTS][kRowInt] = {};
  int32_t def_int[NSGYM_MAX_SLOTS][kRowInt] = {};
  double def_real[NSGYM_MAX_SLOTS][kRowReal] = {};
  double def_dbl[NSGYM_MAX_SLOTS][kRowDbl] = {};
  double bytes_per_env = 0.0;
  bool lean = false;     // no row uses a stochastic scheduler, a cursor / slow update rule or a Dirichlet draw
};

// untyped view of StepIO<R>; the typed launchers reinterpret the real-valued pointers
struct LaunchIO {
  void* state; void* theta; int32_t* t; int32_t* istate; const void* action;
  float* reward; uint8_t* flags; uint8_t* change; void* delta; float* obs;
  const double* inj_u; const double* inj_z; const uint8_t* mask;
  int64_t n, begin, count;
  uint64_t gid_offset, seed, step_index;
  int32_t skip_updates, force_init, prefetch, plan_elapsed;
  int32_t general_kernels;   // NSGYM_OPT_GENERAL_KERNELS
  int32_t sched_replay;      // !persistent_params: stochastic schedulers replay their pattern every episode (base.py:383)
  // rollout
  int32_t k_steps; float gamma; float* ret; int32_t* len;
  const void* policy; int32_t policy_per_env;   // linear (float) / tabular (uint8) rollout policy, NULL = uniform random
  const RowTable* rows;   // heterogeneous handles
  int32_t* kernel_class;  // out (host, may be NULL): which instantiation the launcher picked (NSGYM_KERNEL_*)
  int32_t specialize;     // lean programs: launch the program-specialised kernel (nsgym_jit.cu) when it can be built
  int32_t* specialized;   // out (host, may be NULL): 1 when the launch went to a program-specialised kernel
  std::string* spec_source;   // test entry (nsgym_jit_check): receive the generated source instead of launching
  struct SpecCache* spec_cache;   // per-handle memo of the specialised kernels (the program of a handle never changes)
};

// launch facts -> kernel (NULL = could not be built: keep the precompiled kernel); owned by the handle
struct SpecCache {
  std::vector<std::pair<uint32_t, cudaKernel_t>> slots;
  bool find(uint32_t key, cudaKernel_t* k) const {
    for (const auto& s : slots)
      if (s.first == key) { *k = s.second; return true; }
    return false;
  }
  void put(uint32_t key, cudaKernel_t k) { slots.emplace_back(key, k); }
};

// Program-specialised kernels (nsgym_jit.cu): the device headers compiled again at run time by NVRTC with
// the handle's lowered program as a compile-time constant.
namespace jit {
bool enabled();                                           // false when NSGYM_B200_NO_JIT is set
std::string words(const void* p, size_t bytes);           // a pointer-free object as 32-bit literals "0x..u,0x..u,..."
// compile to an sm_100a cubin (needs NVRTC, no device); 0 or a negative status with `log` filled
int compile(const std::string& source, bool fmad, std::vector<char>* cubin, std::string* log);
// the extern "C" kernel `entry` of `source`, compiled (once per distinct source in the process) and loaded,
// or NULL with `why`
cudaKernel_t kernel(const std::string& source, const char* entry, bool fmad, std::string* why);
struct Stats { int64_t compiled = 0, hits = 0, failed = 0; std::string last_failure; };
Stats stats();
}  // namespace jit

// each returns cudaError_t of the launch (cudaGetLastError)
cudaError_t launch_classic_f32(LaunchOp op, const NsgymSpec& spec, const DevicePools& pools,
                               const LaunchIO& io, cudaStream_t stream);
cudaError_t launch_classic_f64(LaunchOp op, const NsgymSpec& spec, const DevicePools& pools,
                               const LaunchIO& io, cudaStream_t stream);
cudaError_t launch_grid(LaunchOp op, const NsgymSpec& spec, const DevicePools& pools, const LaunchIO& io,
                        cudaStream_t stream);

cudaError_t launch_table(const NsgymSpec& spec, const DevicePools& pools, const LaunchIO& io, int64_t env, int n_times,
                         double* prob, int32_t* next, float* reward, uint8_t* done, cudaStream_t stream);

cudaError_t launch_eval_scalar_f32(const NsgymSpec& spec, const DevicePools& pools, int slot, void* param,
                                   const int32_t* time, int32_t* istate, uint8_t* flag, void* delta,
                                   const double* inj_u, const double* inj_z, int64_t n, uint64_t seed,
                                   uint64_t step_index, cudaStream_t stream);
cudaError_t launch_eval_scalar_f64(const NsgymSpec& spec, const DevicePools& pools, int slot, void* param,
                                   const int32_t* time, int32_t* istate, uint8_t* flag, void* delta,
                                   const double* inj_u, const double* inj_z, int64_t n, uint64_t seed,
                                   uint64_t step_index, cudaStream_t stream);
// native draws of the Philox streams (nsgym_eval_draws)
cudaError_t launch_eval_draws_f32(const LaunchIO& io, int what, int lane, int t, double p, double* out,
                                  cudaStream_t stream);
cudaError_t launch_eval_draws_f64(const LaunchIO& io, int what, int lane, int t, double p, double* out,
                                  cudaStream_t stream);
cudaError_t launch_eval_draws_grid(const LaunchIO& io, int n_dist, int what, int lane, int t, double p, double* out,
                                   cudaStream_t stream);
cudaError_t launch_eval_w1(int dim, const double* u, const double* v, double* out, double* ref, int64_t n,
                           cudaStream_t stream);
cudaError_t launch_eval_dist(const NsgymSpec& spec, const DevicePools& pools, int slot, double* param,
                             const int32_t* time, int32_t* istate, uint8_t* flag, double* delta,
                             const double* inj_u, int64_t n, uint64_t seed, uint64_t step_index,
                             cudaStream_t stream);

// fast_mod magic (device: mod_fire): ceil(2^32 / d), exact while t * d < 2^32 over the reachable t;
// stored in si[2] of Periodic / Burst slots, 0 = use the real modulo
inline void set_mod_magic(NsgymSlot* sl, const NsgymSpec& spec) {
  const uint64_t t_max = (spec.autoreset == NSGYM_AUTORESET_NEXT_STEP && spec.max_episode_steps > 0)
                             ? uint64_t(spec.max_episode_steps) + 1 : (1ull << 28);
  int d = 0;
  if (sl->sched_op == NSGYM_SCHED_PERIODIC) d = sl->si[0];
  if (sl->sched_op == NSGYM_SCHED_BURST) d = sl->si[1];
  if (sl->sched_op == NSGYM_SCHED_PERIODIC || sl->sched_op == NSGYM_SCHED_BURST) {
    sl->si[2] = 0;
    if (d >= 2 && t_max * uint64_t(d) < (1ull << 32))
      sl->si[2] = int32_t(uint32_t(((1ull << 32) + uint64_t(d) - 1) / uint64_t(d)));
  }
}

// per-slot checks of nsgym_create, reused for every row of nsgym_create_rows (nsgym_abi.cu)
int validate_row_slot(const NsgymSpec* spec, const NsgymSlot* slot, int j, char* err, size_t err_len);

// lower rows (NsgymSlot[n_envs][n_slots], host) into a RowTable; `real_is_double` selects the
// word type of the real planes.  Returns 0 or a negative status with `err` filled.
int build_rows(const NsgymSpec& spec, const NsgymSlot* rows, bool real_is_double, RowTable* out, char* err,
               size_t err_len);
void free_rows(RowTable* t);

inline bool is_grid_kind(int k) {
  return k == NSGYM_ENV_FROZENLAKE || k == NSGYM_ENV_CLIFFWALKING || k == NSGYM_ENV_BRIDGE;
}

}  // namespace nsg
