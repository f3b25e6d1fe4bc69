// One launch shape of the small-P kernels (split so the shapes compile in parallel).
#include "bfx_kernels_impl.cuh"

namespace bfx {
cudaError_t launch_eval_tc(const EvalDev& E, cudaStream_t st) { return launch_eval_t<256, 64, true>(E, st); }
cudaError_t launch_mcmc_tc(const RunDev& R, cudaStream_t st) { return launch_mcmc_t<256, 64, true>(R, st); }
#ifdef BFX_TIMING
extern "C" int bfx_debug_timing(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out, g_bfx_timing, sizeof(unsigned long long) * 24) != cudaSuccess) return 1;
  if (reset) {
    unsigned long long z[24] = {0};
    cudaMemcpyToSymbol(g_bfx_timing, z, sizeof z);
  }
  return 0;
}
#endif
}  // namespace bfx
