// fp64 parity mode: classic-control kernels.  Built with -fmad=false so that every product and
// sum rounds separately, as NumPy's do; with FMA contraction `mu + sigma * z` and the Euler
// updates would differ from the reference in the last bit.
#include "nsgym_classic_launch.cuh"

namespace nsg {
cudaError_t launch_classic_f64(LaunchOp op, const NsgymSpec& spec, const DevicePools& pools, const LaunchIO& io,
                               cudaStream_t stream) {
  return launch_classic_t<double>(op, spec, pools, io, stream);
}
cudaError_t launch_eval_scalar_f64(const NsgymSpec& spec, const DevicePools& pools, int slot, void* param,
                                   const int32_t* time, int32_t* istate, uint8_t* flag, void* delta,
                                   const double* inj_u, const double* inj_z, int64_t n, uint64_t seed,
                                   uint64_t step_index, cudaStream_t stream) {
  return launch_eval_scalar_t<double>(spec, pools, slot, param, time, istate, flag, delta, inj_u, inj_z, n, seed,
                                      step_index, stream);
}
cudaError_t launch_eval_draws_f64(const LaunchIO& io, int what, int lane, int t, double p, double* out,
                                  cudaStream_t stream) {
  return launch_eval_draws_t<double>(io, what, lane, t, p, out, stream);
}
}  // namespace nsg
