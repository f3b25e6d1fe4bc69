// Dispatch over the launch shapes (instantiated in bfx_k_*.cu) and the test-hook kernels.
#include "bfx_kernels_impl.cuh"

namespace bfx {

cudaError_t launch_eval_32(const EvalDev& E, cudaStream_t st);
cudaError_t launch_mcmc_32(const RunDev& R, cudaStream_t st);
cudaError_t launch_eval_128(const EvalDev& E, cudaStream_t st);
cudaError_t launch_mcmc_128(const RunDev& R, cudaStream_t st);
cudaError_t launch_eval_256(const EvalDev& E, cudaStream_t st);
cudaError_t launch_mcmc_256(const RunDev& R, cudaStream_t st);
cudaError_t launch_eval_tc(const EvalDev& E, cudaStream_t st);
cudaError_t launch_mcmc_tc(const RunDev& R, cudaStream_t st);
cudaError_t launch_eval_rec(const EvalDev& E, cudaStream_t st);
cudaError_t launch_mcmc_rec(const RunDev& R, cudaStream_t st);

// ---- test hooks -----------------------------------------------------------------
__global__ void k_debug_batch(RunDev R, int c, unsigned long long gstep, long long* out, long long* Bout) {
  BatchSrc bs;
  long long Bk;
  make_batch(R, c, gstep, bs, Bk);
  if (threadIdx.x == 0 && blockIdx.x == 0) *Bout = Bk;
  for (long long j = blockIdx.x * blockDim.x + threadIdx.x; j < Bk; j += (long long)gridDim.x * blockDim.x)
    out[j] = batch_obs(bs, j);
}

// the normal parameter J receives = component (k % 4) of the Philox quad of its padded slot k
__global__ void k_debug_noise(unsigned long long seed, unsigned chain, unsigned long long gstep, unsigned slot, int P,
                              const int* flat_pad, float* normals, float* uni) {
  NoiseKey key{seed, chain};
  int J = blockIdx.x * blockDim.x + threadIdx.x;
  if (J < P) {
    const int k = flat_pad[J];
    float4 z = normal_quad(key, gstep, slot, (unsigned)(k >> 2));
    const int e = k & 3;
    normals[J] = e == 0 ? z.x : e == 1 ? z.y : e == 2 ? z.z : z.w;
  }
  if (J == 0) *uni = uniform_draw(key, gstep);
}

// ---- flat (ABI) <-> padded (engine) state conversion on the device --------------------------------------
__global__ void k_flat_to_padded(const float* __restrict__ flat, float* __restrict__ pad, const int* __restrict__ flat_pad,
                                 int P, int S, long long total) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long c = e / P;
    const int J = (int)(e - c * P);
    pad[c * S + flat_pad[J]] = flat[e];
  }
}
__global__ void k_padded_to_flat(const float* __restrict__ pad, float* __restrict__ flat, const int* __restrict__ flat_pad,
                                 int P, int S, long long total) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long c = e / P;
    const int J = (int)(e - c * P);
    flat[e] = pad[c * S + flat_pad[J]];
  }
}
// x (d x n, column-major) -> rows of [kpad/8 hi chunks | kpad/8 lo chunks], 8 bf16 per 16-byte chunk, zero beyond d:
// the staging format of the tensor-core path (tc::chain_eval_tc)
__global__ void k_split_x(const float* __restrict__ x, uint4* __restrict__ out, long long n, int d, int nch) {
  const long long total = n * nch;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / nch;
    const int ch = (int)(e - r * nch);
    float v[8];
#pragma unroll
    for (int f = 0; f < 8; ++f) {
      const int ff = 8 * ch + f;
      v[f] = ff < d ? x[r * d + ff] : 0.f;
    }
    uint4 hh, ll;
    tc::split2(v[0], v[1], hh.x, ll.x);
    tc::split2(v[2], v[3], hh.y, ll.y);
    tc::split2(v[4], v[5], hh.z, ll.z);
    tc::split2(v[6], v[7], hh.w, ll.w);
    out[r * (2 * nch) + ch] = hh;
    out[r * (2 * nch) + nch + ch] = ll;
  }
}
__global__ void k_fill(float* p, long long n, float v) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) p[e] = v;
}

cudaError_t launch_flat_to_padded(const float* flat, float* pad, const int* flat_pad, int P, int S, int C, float fill,
                                  cudaStream_t st) {
  k_fill<<<592, 256, 0, st>>>(pad, (long long)S * C, fill);
  k_flat_to_padded<<<1184, 256, 0, st>>>(flat, pad, flat_pad, P, S, (long long)P * C);
  return cudaGetLastError();
}
cudaError_t launch_split_x(const float* x, void* out, long long n, int d, int kpad, cudaStream_t st) {
  k_split_x<<<1184, 256, 0, st>>>(x, static_cast<uint4*>(out), n, d, kpad >> 3);
  return cudaGetLastError();
}
cudaError_t launch_padded_to_flat(const float* pad, float* flat, const int* flat_pad, int P, int S, int C, cudaStream_t st) {
  k_padded_to_flat<<<1184, 256, 0, st>>>(pad, flat, flat_pad, P, S, (long long)P * C);
  return cudaGetLastError();
}

size_t smallp_smem_bytes(const ModelDev& M, int bt) {
  return bt == 16 ? chain_smem_bytes<16>(M) : bt == 32 ? chain_smem_bytes<32>(M) : chain_smem_bytes<64>(M);
}

cudaError_t launch_eval(const EvalDev& E, const LaunchCfg& cfg, cudaStream_t st) {
  if (E.M.tc.enabled) return launch_eval_tc(E, st);
  if (cfg.bt == 16) return launch_eval_rec(E, st);
  if (cfg.nt == 32) return launch_eval_32(E, st);
  if (cfg.nt == 128) return launch_eval_128(E, st);
  return launch_eval_256(E, st);
}

cudaError_t launch_mcmc(const RunDev& R, const LaunchCfg& cfg, cudaStream_t st) {
  if (R.M.tc.enabled) return launch_mcmc_tc(R, st);
  if (cfg.bt == 16) return launch_mcmc_rec(R, st);
  if (cfg.nt == 32) return launch_mcmc_32(R, st);
  if (cfg.nt == 128) return launch_mcmc_128(R, st);
  return launch_mcmc_256(R, st);
}

cudaError_t launch_debug_batch(const RunDev& R, int c, unsigned long long gstep, long long* out, long long* Bout,
                               cudaStream_t st) {
  k_debug_batch<<<64, 256, 0, st>>>(with_fastdiv(R), c, gstep, out, Bout);
  return cudaGetLastError();
}

cudaError_t launch_debug_noise(unsigned long long seed, unsigned chain, unsigned long long gstep, unsigned slot, int P,
                               const int* flat_pad, float* normals, float* uni, cudaStream_t st) {
  k_debug_noise<<<(P + 127) / 128, 128, 0, st>>>(seed, chain, gstep, slot, P, flat_pad, normals, uni);
  return cudaGetLastError();
}

}  // namespace bfx
