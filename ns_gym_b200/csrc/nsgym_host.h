// nsgym_host.h -- internal host-side declarations shared by the translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nsgym_b200.h"

namespace nsg {

enum LaunchOp { OP_STEP = 0, OP_RESET = 1, OP_ROLLOUT = 2 };

struct DevicePools {
  const double* pool_f;
  const int32_t* pool_i;
  const uint32_t* bitmap;
};

// untyped view of StepIO<R>; the typed launchers reinterpret the real-valued pointers
struct LaunchIO {
  void* state; void* theta; int32_t* t; int32_t* istate; const void* action;
  float* reward; uint8_t* flags; uint8_t* change; void* delta; float* obs;
  const double* inj_u; const double* inj_z; const uint8_t* mask;
  int64_t n, begin, count;
  uint64_t gid_offset, seed, step_index;
  int32_t skip_updates, force_init, prefetch;
  // rollout
  int32_t k_steps; float gamma; float* ret; int32_t* len;
};

// each returns cudaError_t of the launch (cudaGetLastError)
cudaError_t launch_classic_f32(LaunchOp op, const NsgymSpec& spec, const DevicePools& pools,
                               const LaunchIO& io, cudaStream_t stream);
cudaError_t launch_classic_f64(LaunchOp op, const NsgymSpec& spec, const DevicePools& pools,
                               const LaunchIO& io, cudaStream_t stream);
cudaError_t launch_grid(LaunchOp op, const NsgymSpec& spec, const DevicePools& pools, const LaunchIO& io,
                        cudaStream_t stream);

cudaError_t launch_eval_scalar_f32(const NsgymSpec& spec, const DevicePools& pools, int slot, void* param,
                                   const int32_t* time, int32_t* istate, uint8_t* flag, void* delta,
                                   const double* inj_u, const double* inj_z, int64_t n, uint64_t seed,
                                   uint64_t step_index, cudaStream_t stream);
cudaError_t launch_eval_scalar_f64(const NsgymSpec& spec, const DevicePools& pools, int slot, void* param,
                                   const int32_t* time, int32_t* istate, uint8_t* flag, void* delta,
                                   const double* inj_u, const double* inj_z, int64_t n, uint64_t seed,
                                   uint64_t step_index, cudaStream_t stream);
cudaError_t launch_eval_dist(const NsgymSpec& spec, const DevicePools& pools, int slot, double* param,
                             const int32_t* time, int32_t* istate, uint8_t* flag, double* delta,
                             const double* inj_u, int64_t n, uint64_t seed, uint64_t step_index,
                             cudaStream_t stream);

inline bool is_grid_kind(int k) {
  return k == NSGYM_ENV_FROZENLAKE || k == NSGYM_ENV_CLIFFWALKING || k == NSGYM_ENV_BRIDGE;
}

}  // namespace nsg
