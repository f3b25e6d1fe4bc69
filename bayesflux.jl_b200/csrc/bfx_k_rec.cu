// One launch shape of the small-P kernels (split so the shapes compile in parallel).
#include "bfx_kernels_impl.cuh"

namespace bfx {
cudaError_t launch_eval_rec(const EvalDev& E, cudaStream_t st) { return launch_eval_t<256, 16>(E, st); }
cudaError_t launch_mcmc_rec(const RunDev& R, cudaStream_t st) { return launch_mcmc_t<256, 16>(R, st); }
}  // namespace bfx
