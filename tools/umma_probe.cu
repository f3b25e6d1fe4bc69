// umma_probe: pins the tcgen05 (UMMA) conventions the TC kernels of libbfx rely on, on real hardware:
//   * shared-memory matrix descriptors for the no-swizzle "core block" tiling, K-major and MN-major,
//     for kind::f16 (bf16 inputs) and kind::tf32
//   * the instruction descriptor (M, N, majorness, formats)
//   * where an M = 64 accumulator lands in TMEM and how tcgen05.ld 32x32b / 16x256b hand it to threads
//   * issue-to-completion time of back-to-back MMAs for the shapes the fused kernel uses
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_probe.bin tools/umma_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version 1 (Blackwell)
  return d;                // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}

// kind: 0 = f16 pipe with BF16 inputs, 1 = tf32
__device__ __forceinline__ uint32_t make_idesc(int kind, int M, int N, int a_mn, int b_mn) {
  uint32_t fmt = kind == 0 ? 1u /*BF16*/ : 2u /*TF32*/;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_16x256b_x1(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

struct Case {
  int kind;       // 0 bf16, 1 tf32
  int M, N, K;    // logical GEMM  D[M x N] = A[M x K] * B[N x K]^T
  int a_mn, b_mn; // 0: operand stored K-major, 1: MN-major
  int reps;       // timing: how many times the whole K loop is issued (accumulating)
  int mode;       // 0 = correctness dump, 1 = timing only
};

// core-block tiling: element (r, c) of a tile whose contiguous 16-byte runs are along c
//   off = (r/8)*S_r + (c/T)*S_c + (r%8)*16 + (c%T)*esz,   T = 16/esz, S_c = 128, S_r = (C/T)*128
__device__ __forceinline__ uint32_t tile_off(int r, int c, int Ccols, int esz) {
  const int T = 16 / esz;
  return (uint32_t)((r >> 3) * ((Ccols / T) * 128) + (c / T) * 128 + (r & 7) * 16 + (c % T) * esz);
}

// stores logical X[mn][k] into the tile, for the requested majorness
__device__ void fill_operand(uint8_t* tile, const float* X, int MN, int K, int mn_major, int kind) {
  const int esz = kind == 0 ? 2 : 4;
  for (int e = threadIdx.x; e < MN * K; e += blockDim.x) {
    int mn = e / K, k = e - mn * K;
    float v = X[e];
    uint32_t off = mn_major ? tile_off(k, mn, MN, esz) : tile_off(mn, k, K, esz);
    if (kind == 0)
      *reinterpret_cast<__nv_bfloat16*>(tile + off) = __float2bfloat16(v);
    else
      *reinterpret_cast<float*>(tile + off) = v;
  }
}

__global__ void __launch_bounds__(128) probe(Case cs, const float* A, const float* B, float* dump32 /*[128][N]*/,
                                             float* dump16 /*[4 warps][32 thr][8]*/, long long* cycles,
                                             uint32_t* tmem_base_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int esz = cs.kind == 0 ? 2 : 4;
  const int T = 16 / esz;
  const int KI = 32 / esz;  // K per instruction
  uint8_t* tA = smem;
  uint8_t* tB = smem + cs.M * cs.K * esz;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  fill_operand(tA, A, cs.M, cs.K, cs.a_mn, cs.kind);
  fill_operand(tB, B, cs.N, cs.K, cs.b_mn, cs.kind);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *tmem_base_out = tmem_base;
  // zero the accumulator region first so that untouched lanes are recognisable
  for (int c = 0; c < 128; ++c) tmem_st_32x32b_x1(tmem_base + ((uint32_t)(warp * 32) << 16) + c, __float_as_uint(-777.f));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(cs.kind, cs.M, cs.N, cs.a_mn, cs.b_mn);
    // K-major: SBO = stride between 8-row groups = (K/T)*128, LBO = 128 (next 16-byte K chunk); K step = 2 chunks
    // MN-major: SBO = 128 (next group of T MN-elements), LBO = (MN/T)*128 (next 8 k-rows)
    uint32_t a_lbo, a_sbo, a_step, b_lbo, b_sbo, b_step;
    if (!cs.a_mn) { a_sbo = (cs.K / T) * 128; a_lbo = 128; a_step = (KI / T) * 128; }
    else          { a_sbo = 128; a_lbo = (cs.M / T) * 128; a_step = (KI / 8) * (cs.M / T) * 128; }
    if (!cs.b_mn) { b_sbo = (cs.K / T) * 128; b_lbo = 128; b_step = (KI / T) * 128; }
    else          { b_sbo = 128; b_lbo = (cs.N / T) * 128; b_step = (KI / 8) * (cs.N / T) * 128; }
    const uint32_t a0 = smem_u32(tA), b0 = smem_u32(tB);
    long long t0 = clock64();
    for (int rep = 0; rep < cs.reps; ++rep)
      for (int kk = 0; kk < cs.K / KI; ++kk) {
        uint64_t da = make_desc(a0 + kk * a_step, a_lbo, a_sbo);
        uint64_t db = make_desc(b0 + kk * b_step, b_lbo, b_sbo);
        if (cs.kind == 0) mma_bf16(tmem_base, da, db, idesc, (rep | kk) ? 1u : 0u);
        else mma_tf32(tmem_base, da, db, idesc, (rep | kk) ? 1u : 0u);
      }
    long long t1 = clock64();
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) {
      cycles[0] = t1 - t0;
      cycles[1] = t2 - t0;
    }
  }
  __syncthreads();
  tc_fence_after();
  if (cs.mode == 0 && blockIdx.x == 0) {
    // (a) raw lane/column dump through 32x32b: thread t of warp w reads lane 32w+t
    for (int c0 = 0; c0 < cs.N; c0 += 8) {
      uint32_t v[8];
      tmem_ld_32x32b_x8(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
      for (int j = 0; j < 8; ++j) dump32[(size_t)threadIdx.x * cs.N + c0 + j] = __uint_as_float(v[j]);
    }
    // (b) 16x256b.x2 fragment at column 0 (16 lanes of the warp's quadrant, 16 columns)
    uint32_t f[8];
    tmem_ld_16x256b_x2(tmem_base + ((uint32_t)(warp * 32) << 16), f);
    for (int j = 0; j < 8; ++j) dump16[(size_t)threadIdx.x * 8 + j] = __uint_as_float(f[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
}

static float bf16_round(float v) { return __bfloat162float(__float2bfloat16(v)); }

int main(int argc, char** argv) {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp pr;
  CK(cudaGetDeviceProperties(&pr, dev));
  printf("device %s sm_%d%d SMs %d\n", pr.name, pr.major, pr.minor, pr.multiProcessorCount);
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  float *dA, *dB, *d32, *d16;
  long long* dcyc;
  uint32_t* dtb;
  CK(cudaMalloc(&dA, 128 * 256 * 4));
  CK(cudaMalloc(&dB, 256 * 256 * 4));
  CK(cudaMalloc(&d32, 128 * 256 * 4));
  CK(cudaMalloc(&d16, 128 * 8 * 4));
  CK(cudaMalloc(&dcyc, 16));
  CK(cudaMalloc(&dtb, 4));
  int nfail = 0;
  // ---------------- correctness: every majorness combination, both kinds, the N values used ----------------
  const int Ns[] = {64, 8, 16, 24, 72};
  for (int kind = 0; kind < 2; ++kind)
    for (int M : {64, 128})
      for (int ni = 0; ni < 5; ++ni)
        for (int amn = 0; amn < 2; ++amn)
          for (int bmn = 0; bmn < 2; ++bmn) {
            const int N = Ns[ni], K = 64;
            if (M == 128 && (N % 16)) continue;
            if (bmn && kind == 0 && (N % 8)) continue;
            Case cs{kind, M, N, K, amn, bmn, 1, 0};
            std::vector<float> A(M * K), B(N * K), D(M * N, 0.f);
            srand(1234 + kind * 100 + ni * 10 + amn * 2 + bmn + M);
            for (auto& v : A) v = (float)(rand() % 9 - 4);
            for (auto& v : B) v = (float)(rand() % 9 - 4);
            for (int m = 0; m < M; ++m)
              for (int n = 0; n < N; ++n) {
                float s = 0;
                for (int k = 0; k < K; ++k) s += A[m * K + k] * B[n * K + k];
                D[m * N + n] = s;
              }
            CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
            CK(cudaMemset(d32, 0, 128 * 256 * 4));
            size_t smem = (size_t)(M + N) * K * 4 + 1024;
            probe<<<1, 128, smem>>>(cs, dA, dB, d32, d16, dcyc, dtb);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
              printf("CASE kind=%d M=%d N=%d amn=%d bmn=%d: CUDA ERROR %s\n", kind, M, N, amn, bmn, cudaGetErrorString(e));
              return 3;
            }
            std::vector<float> H(128 * N);
            CK(cudaMemcpy(H.data(), d32, H.size() * 4, cudaMemcpyDeviceToHost));
            // expected lane of row m: M=128 -> m; M=64 -> (m%16) + 32*(m/16)
            int bad = 0, bad_alt = 0;
            for (int m = 0; m < M; ++m) {
              int lane = M == 128 ? m : (m % 16) + 32 * (m / 16);
              for (int n = 0; n < N; ++n) {
                if (H[lane * N + n] != D[m * N + n]) ++bad;
                if (H[m * N + n] != D[m * N + n]) ++bad_alt;  // alternative hypothesis: lane == m
              }
            }
            uint32_t tb;
            CK(cudaMemcpy(&tb, dtb, 4, cudaMemcpyDeviceToHost));
            printf("CASE kind=%s M=%3d N=%3d A=%s B=%s : mismatches %d (lane==m hypothesis: %d) tmem_base=0x%x %s\n",
                   kind ? "tf32" : "bf16", M, N, amn ? "MN" : "K ", bmn ? "MN" : "K ", bad, bad_alt, tb,
                   bad == 0 ? "OK" : "FAIL");
            if (bad) {
              ++nfail;
              if (nfail <= 3) {
                printf("   first rows of dump (lane 0..3, col 0..7) vs expected row 0..3:\n");
                for (int l = 0; l < 4; ++l) {
                  printf("   lane %d:", l);
                  for (int n = 0; n < 8 && n < N; ++n) printf(" %6.0f", H[l * N + n]);
                  printf("   | exp:");
                  for (int n = 0; n < 8 && n < N; ++n) printf(" %6.0f", D[l * N + n]);
                  printf("\n");
                }
              }
            }
          }
  // ---------------- fragment maps: D[m][n] = 64 m + n ----------------
  for (int M : {64, 128}) {
    const int N = 64, K = 64;
    Case cs{0, M, N, K, 0, 0, 1, 0};
    std::vector<float> A(M * K, 0.f), B(N * K, 0.f);
    for (int m = 0; m < M; ++m) { A[m * K + 0] = (float)m; A[m * K + 1] = 1.f; }
    for (int n = 0; n < N; ++n) { B[n * K + 0] = 64.f; B[n * K + 1] = (float)n; }
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    probe<<<1, 128, (size_t)(M + N) * K * 4 + 1024>>>(cs, dA, dB, d32, d16, dcyc, dtb);
    CK(cudaDeviceSynchronize());
    std::vector<float> F(128 * 8);
    CK(cudaMemcpy(F.data(), d16, F.size() * 4, cudaMemcpyDeviceToHost));
    printf("FRAG 16x256b.x2 M=%d: thread -> (row,col) of its 8 registers (warp 0 and warp 1 shown)\n", M);
    for (int t = 0; t < 64; ++t) {
      if (t % 32 >= 10 && t % 32 < 28) continue;
      printf("  t%3d:", t);
      for (int j = 0; j < 8; ++j) {
        float v = F[t * 8 + j];
        if (v == -777.f) printf(" (--,--)");
        else printf(" (%2d,%2d)", (int)v / 64, (int)v % 64);
      }
      printf("\n");
    }
  }
  // ---------------- timing ----------------
  struct T { int kind, M, N, amn, bmn; };
  const T tcs[] = {{0, 64, 64, 0, 0}, {0, 64, 64, 1, 1}, {0, 64, 64, 1, 0}, {0, 64, 64, 0, 1}, {0, 64, 8, 0, 0},
                   {0, 64, 16, 0, 0}, {0, 64, 72, 0, 0}, {0, 128, 64, 0, 0}, {1, 64, 64, 0, 0}, {1, 64, 64, 1, 1},
                   {1, 128, 64, 0, 0}};
  for (const T& t : tcs) {
    for (int blocks : {1, 148, 296}) {
      Case cs{t.kind, t.M, t.N, 64, t.amn, t.bmn, 16, 1};
      size_t smem = (size_t)(t.M + t.N) * 64 * 4 + 1024;
      probe<<<blocks, 128, smem>>>(cs, dA, dB, d32, d16, dcyc, dtb);
      CK(cudaDeviceSynchronize());
      long long cyc[2];
      CK(cudaMemcpy(cyc, dcyc, 16, cudaMemcpyDeviceToHost));
      int nmma = 16 * 64 / (t.kind == 0 ? 16 : 8);
      printf("TIME kind=%s M=%3d N=%3d A=%s B=%s blocks=%3d: %d MMAs issue %lld cyc, issue+complete %lld cyc -> %.1f cyc/MMA\n",
             t.kind ? "tf32" : "bf16", t.M, t.N, t.amn ? "MN" : "K ", t.bmn ? "MN" : "K ", blocks, nmma, cyc[0], cyc[1],
             (double)cyc[1] / nmma);
    }
  }
  printf("SUMMARY: %d failing correctness cases\n", nfail);
  return nfail ? 1 : 0;
}
