// One launch shape of the small-P kernels (split so the shapes compile in parallel).
#include "bfx_kernels_impl.cuh"

namespace bfx {
cudaError_t launch_eval_32(const EvalDev& E, cudaStream_t st) { return launch_eval_t<32, 32>(E, st); }
cudaError_t launch_mcmc_32(const RunDev& R, cudaStream_t st) { return launch_mcmc_t<32, 32>(R, st); }
cudaError_t launch_eval_128(const EvalDev& E, cudaStream_t st) { return launch_eval_t<128, 64>(E, st); }
cudaError_t launch_mcmc_128(const RunDev& R, cudaStream_t st) { return launch_mcmc_t<128, 64>(R, st); }
}  // namespace bfx
