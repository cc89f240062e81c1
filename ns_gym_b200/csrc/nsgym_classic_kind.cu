// Kernels of ONE classic-control env kind in ONE precision: compiled once per
// (-DNSGYM_TU_REAL=float|double, -DNSGYM_TU_KIND=0..4) so that the ten units build in parallel.
// fp64 units are built with -fmad=false (every product and sum rounds separately, as NumPy's do).
#include "nsgym_classic_launch.cuh"

namespace nsg {
template cudaError_t launch_classic_kind<NSGYM_TU_REAL, NSGYM_TU_KIND>(LaunchOp, const NsgymSpec&, const DevicePools&,
                                                                        const LaunchIO&, cudaStream_t);
}  // namespace nsg
