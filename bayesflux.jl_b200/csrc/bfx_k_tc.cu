// One launch shape of the small-P kernels (split so the shapes compile in parallel).
#include "bfx_kernels_impl.cuh"

namespace bfx {
cudaError_t launch_eval_tc(const EvalDev& E, cudaStream_t st) { return launch_eval_t<256, 64, true>(E, st); }
// Shape-specialised instances (TcFixed, bfx_common.cuh): layouts that `matches` proves identical to the run-time
// descriptors run a build whose layer loops are unrolled and whose offsets are literals; everything else (and every
// sampler kind not compiled for a shape) runs the generic TcDyn instance.  BFX_TC_GENERIC=1 forces the generic one.
using ShapeMlp10x64x64Relu = TcFixed<10, 64, 64, A_RELU, A_RELU>;
constexpr int kFixedKinds = (1 << S_SGLD) | (1 << S_SGNHT) | (1 << S_SGNHTS) | (1 << S_GGMC) | (1 << S_HMC);
cudaError_t launch_mcmc_tc(const RunDev& R, cudaStream_t st) {
  static const bool generic_only = std::getenv("BFX_TC_GENERIC") != nullptr;
  if (!generic_only && ((kFixedKinds >> R.S.kind) & 1) && R.M.L[2].act == A_IDENTITY && ShapeMlp10x64x64Relu::matches(R.M))
    return launch_mcmc_t<256, 64, true, ShapeMlp10x64x64Relu, kFixedKinds>(R, st);
  return launch_mcmc_t<256, 64, true>(R, st);
}
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
