// Host-side declarations of the kernel launchers (implemented in bfx_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include "bfx_common.cuh"

namespace bfx {

struct LaunchCfg {
  int nt;  // threads per CTA: 32 | 128 | 256
  int bt;  // batch tile: 32 | 64
};

size_t smallp_smem_bytes(const ModelDev& M, int bt);
cudaError_t launch_eval(const EvalDev& E, const LaunchCfg& cfg, cudaStream_t st);
cudaError_t launch_mcmc(const RunDev& R, const LaunchCfg& cfg, cudaStream_t st);
cudaError_t launch_split_x(const float* x, void* out, long long n, int d, int kpad, cudaStream_t st);
cudaError_t launch_debug_batch(const RunDev& R, int c, unsigned long long gstep, long long* out, long long* Bout,
                               cudaStream_t st);
cudaError_t launch_debug_noise(unsigned long long seed, unsigned chain, unsigned long long gstep, unsigned slot, int P,
                               const int* flat_pad, float* normals, float* uni, cudaStream_t st);

cudaError_t launch_flat_to_padded(const float* flat, float* pad, const int* flat_pad, int P, int S, int C, float fill,
                                  cudaStream_t st);
cudaError_t launch_padded_to_flat(const float* pad, float* flat, const int* flat_pad, int P, int S, int C, cudaStream_t st);

}  // namespace bfx
