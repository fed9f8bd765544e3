// lbm_bulk_tma.cu -- TMA-pipelined persistent variant of the bulk collide-stream kernel.
// (placeholder until the kernel lands: reports "not supported" so that launch_bulk falls back)
#include "lbm_kernels.cuh"
namespace lbm {
cudaError_t launch_bulk_tma(bool, const StepArgs&, cudaStream_t, int, int) { return cudaErrorNotSupported; }
}  // namespace lbm
