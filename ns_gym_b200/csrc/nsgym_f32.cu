// fp32 fast mode (FMA contraction on): classic-control kernels.
#include "nsgym_classic_launch.cuh"

namespace nsg {
cudaError_t launch_classic_f32(LaunchOp op, const NsgymSpec& spec, const DevicePools& pools, const LaunchIO& io,
                               cudaStream_t stream) {
  return launch_classic_t<float>(op, spec, pools, io, stream);
}
cudaError_t launch_eval_scalar_f32(const NsgymSpec& spec, const DevicePools& pools, int slot, void* param,
                                   const int32_t* time, int32_t* istate, uint8_t* flag, void* delta,
                                   const double* inj_u, const double* inj_z, int64_t n, uint64_t seed,
                                   uint64_t step_index, cudaStream_t stream) {
  return launch_eval_scalar_t<float>(spec, pools, slot, param, time, istate, flag, delta, inj_u, inj_z, n, seed,
                                     step_index, stream);
}
cudaError_t launch_eval_draws_f32(const LaunchIO& io, int what, int lane, int t, double p, double* out,
                                  cudaStream_t stream) {
  return launch_eval_draws_t<float>(io, what, lane, t, p, out, stream);
}
}  // namespace nsg
