// umma_probe2: issue/throughput characteristics of small tcgen05 MMAs (bf16, K-major operands).
//   variant a: NMMA dependent MMAs into one accumulator
//   variant b: round-robin over NACC accumulators (independent chains)
//   variant c: W warps issue concurrently, each its own accumulator
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_probe2.bin tools/umma_probe2.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(2);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma_bf16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory"); } while (!ok);
}
struct P { int M, N, nmma, nacc, nwarps; };
__global__ void __launch_bounds__(128) k(P p, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tb;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tb)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x < 4) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[threadIdx.x])), "r"(1u) : "memory");
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3f803f80u;  // bf16 1.0
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tb;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(p.M >> 4) << 24);
  // K = 64 tiles: A at 0 (M x 64 bf16 <= 16 KB), B at 16 KB (N x 64 bf16 <= 32 KB)
  const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 16384);
  long long t0 = 0, t1 = 0, t2 = 0;
  if (warp < p.nwarps && lane == 0) {
    const int ncols = ((p.N + 31) / 32) * 32;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < p.nmma; ++i) {
      const int kk = i & 3;
      const int acc = (i % p.nacc);
      uint64_t da = make_desc(a0 + kk * 256, 128, 1024);
      uint64_t db = make_desc(b0 + kk * 256, 128, 1024);
      mma_bf16(tmem + (uint32_t)((warp * p.nacc + acc) * ncols), da, db, idesc, i >= p.nacc ? 1u : 0u);
    }
    t1 = clock64();
    mma_commit(&bar[warp]);
    mbar_wait(&bar[warp], 0);
    t2 = clock64();
    if (blockIdx.x == 0) { out[warp * 2] = t1 - t0; out[warp * 2 + 1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
int main() {
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  long long* d; CK(cudaMalloc(&d, 64));
  const P ps[] = {
    {64, 64, 64, 1, 1}, {64, 64, 64, 2, 1}, {64, 64, 64, 4, 1}, {64, 64, 64, 8, 1}, {64, 64, 256, 1, 1}, {64, 64, 256, 4, 1},
    {64, 8, 64, 1, 1}, {64, 8, 64, 4, 1}, {64, 16, 64, 1, 1}, {64, 128, 64, 1, 1}, {64, 256, 64, 1, 1}, {64, 256, 64, 2, 1},
    {128, 64, 64, 1, 1}, {128, 64, 64, 4, 1}, {128, 128, 64, 1, 1}, {128, 256, 64, 1, 1}, {128, 256, 64, 2, 1},
    {64, 64, 64, 1, 2}, {64, 64, 64, 1, 4}, {64, 64, 64, 2, 2}, {128, 64, 64, 1, 4}, {64, 8, 64, 1, 4},
  };
  for (const P& p : ps) for (int blocks : {1, 148}) {
    long long h[8] = {0};
    CK(cudaMemset(d, 0, 64));
    k<<<blocks, 128, 56 * 1024>>>(p, d);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost));
    printf("M=%3d N=%3d nmma=%3d nacc=%d warps=%d blocks=%3d : issue %6lld total %6lld cyc -> %.1f cyc/MMA per issuing warp (floor %d)\n", p.M, p.N,
           p.nmma, p.nacc, p.nwarps, blocks, h[0], h[1], (double)h[1] / p.nmma, 128 * p.N / 256);
  }
  return 0;
}
