// tcgen05 GEMM of the large-P ("stream") regime: C[M x N] = A * B with every operand stored in HBM as two
// BF16 arrays (hi, lo = the 3-term split of the FP32 value, see bfx_tc.cuh), column-major.  One CTA computes a
// 128 x 128 tile: K is consumed in chunks of 32 through a 3-stage cp.async ring (96 KB: two CTAs per SM) of no-swizzle core-block tiles,
// one thread issues tcgen05.mma.cta_group::1.kind::f16 (M = 128, N = 128, K = 16; hi*hi + hi*lo + lo*hi per
// k-step) into a 128-column FP32 accumulator in TMEM, and the epilogue reads it back with tcgen05.ld.32x32b
// (thread = accumulator row) fused with bias + activation / the act' multiply / the split-K partial store.
// Both operand majornesses are served from the natural column-major arrays through the descriptor, so the
// three GEMMs of a Dense layer (forward, input gradient, weight gradient) need no transposed copies.
//
// Reference arithmetic replaced: Flux Dense forward (called at src/likelihoods/feedforward.jl:35,91) and its
// Zygote pullbacks (src/model/BNN.jl:66) for layers wide enough to be real GEMMs (configs[2]: 512 x 512 x 8192).
#include <cstdlib>

#include <cuda_bf16.h>

#include "bfx_stream.cuh"
#include "bfx_tc.cuh"

namespace bfx {

using tc::commit;
using tc::fence_after;
using tc::fence_async_smem;
using tc::fence_before;
using tc::make_desc;
using tc::mbar_init;
using tc::mbar_wait;
using tc::smem_u32;

constexpr int TBM = 128, TBN = 128, TBK = 32, TSTAGES = 3;
constexpr int TILE_BYTES = 128 * TBK * 2;              // one operand half (hi or lo) of one stage (8 KB)
constexpr int STAGE_BYTES = 4 * TILE_BYTES;            // A hi, A lo, B hi, B lo

BFX_DEV void cp_async16(void* smem, const void* gmem, bool valid) {
  const unsigned n = valid ? 16u : 0u;  // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem)), "l"(gmem), "r"(n) : "memory");
}
BFX_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
BFX_DEV void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

BFX_DEV uint32_t idesc_128x128(int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(TBN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
}

template <class F>
BFX_DEV void with_act_tc(int act, F f) { tc::with_act(act, f); }

BFX_DEV void mma_issue(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  tc::mma(tmem_d, da, db, idesc, acc);
}

// 32 consecutive accumulator columns of this thread's row
BFX_DEV void ld_32x32b_x32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// load one 128 x 64 operand tile (both halves) of K-chunk [k0, k0+64) into its core-block layout
template <bool MN_MAJOR>
BFX_DEV void load_tile(unsigned char* s_hi, unsigned char* s_lo, const TcOperand& op, long long mn0, long long mn_lim,
                       long long k0, long long k_lim) {
  constexpr int KCH = TBK / 8;  // 16-byte chunks along k per row
  for (int q = threadIdx.x; q < 128 * KCH; q += 256) {
    long long gidx;
    int soff;
    bool ok;
    if (!MN_MAJOR) {  // rows = mn (128), 16-byte chunks along k
      const int r = q / KCH, kc = q % KCH;
      const long long mn = mn0 + r, k = k0 + kc * 8;
      ok = mn < mn_lim && k < k_lim;
      gidx = mn * op.ld + k;
      soff = (r >> 3) * (KCH * 128) + kc * 128 + (r & 7) * 16;
    } else {          // rows = k (TBK), 16-byte chunks along mn
      const int kk = q >> 4, mc = q & 15;
      const long long mn = mn0 + mc * 8, k = k0 + kk;
      ok = mn < mn_lim && k < k_lim;
      gidx = k * op.ld + mn;
      soff = (kk >> 3) * 2048 + mc * 128 + (kk & 7) * 16;
    }
    if (!ok) gidx = 0;
    cp_async16(s_hi + soff, op.hi + gidx, ok);
    cp_async16(s_lo + soff, op.lo + gidx, ok);
  }
}

template <bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(256, 2) k_tc_gemm(const TcGemmArgs g) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar[TSTAGES + 1];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long m0 = (long long)blockIdx.x * TBM, n0 = (long long)blockIdx.y * TBN;
  const long long kbeg = (long long)blockIdx.z * g.kchunk;
  long long kend = kbeg + g.kchunk;
  if (kend > g.K) kend = g.K;
  const int nk = (int)((kend - kbeg + TBK - 1) / TBK);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid < TSTAGES + 1) mbar_init(&bar[tid], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = tmem_slot;
  auto stage_ptr = [&](int s, int which) { return smem + s * STAGE_BYTES + which * TILE_BYTES; };
  auto issue_loads = [&](int kc) {
    if (kc < nk && !(g.debug & 4)) {
      const int s = kc % TSTAGES;
      const long long k0 = kbeg + (long long)kc * TBK;
      load_tile<A_MN>(stage_ptr(s, 0), stage_ptr(s, 1), g.A, m0, g.M, k0, kend);
      load_tile<B_MN>(stage_ptr(s, 2), stage_ptr(s, 3), g.B, n0, g.N, k0, kend);
    }
    cp_async_commit();
  };
  for (int kc = 0; kc < TSTAGES - 1; ++kc) issue_loads(kc);
  const uint32_t idesc = idesc_128x128(A_MN ? 1 : 0, B_MN ? 1 : 0);
  uint32_t parity = 0;  // bit s = phase parity of bar[s]
  for (int kc = 0; kc < nk; ++kc) {
    const int s = kc % TSTAGES;
    cp_async_wait<TSTAGES - 2>();  // chunk kc has landed (this thread's copies)
    fence_async_smem();
    fence_before();
    __syncthreads();
    if (tid == 0 && !(g.debug & 2)) {
      fence_after();
      const uint32_t ahi = smem_u32(stage_ptr(s, 0)), alo = smem_u32(stage_ptr(s, 1));
      const uint32_t bhi = smem_u32(stage_ptr(s, 2)), blo = smem_u32(stage_ptr(s, 3));
#pragma unroll
      for (int k = 0; k < TBK / 16; ++k) {
        // K-major tile: LBO = 128 (next 16-byte k chunk), SBO = (TBK/8)*128 (next 8 rows), 16 k = 256 bytes
        // MN-major tile: SBO = 128 (next 8 mn), LBO = 2048 (next 8 k rows), 16 k = 4096 bytes
        constexpr uint32_t KSBO = (TBK / 8) * 128;
        const uint64_t dah = A_MN ? make_desc(ahi + k * 4096, 2048, 128) : make_desc(ahi + k * 256, 128, KSBO);
        const uint64_t dal = A_MN ? make_desc(alo + k * 4096, 2048, 128) : make_desc(alo + k * 256, 128, KSBO);
        const uint64_t dbh = B_MN ? make_desc(bhi + k * 4096, 2048, 128) : make_desc(bhi + k * 256, 128, KSBO);
        const uint64_t dbl = B_MN ? make_desc(blo + k * 4096, 2048, 128) : make_desc(blo + k * 256, 128, KSBO);
        mma_issue(tmem, dah, dbh, idesc, (kc | k) ? 1u : 0u);
        mma_issue(tmem, dah, dbl, idesc, 1u);
        mma_issue(tmem, dal, dbh, idesc, 1u);
      }
      commit(&bar[s]);
    } else if (tid == 0) {
      commit(&bar[s]);
    }
    // the stage of chunk kc-1 is refilled with chunk kc + TSTAGES - 1 once its MMAs have completed
    if (kc >= 1) {
      const int sp = (kc - 1) % TSTAGES;
      mbar_wait(&bar[sp], (parity >> sp) & 1u);
      parity ^= 1u << sp;
    }
    issue_loads(kc + TSTAGES - 1);
  }
  // all MMAs done?
  {
    const int sl = (nk - 1) % TSTAGES;
    if (nk >= 1) {
      mbar_wait(&bar[sl], (parity >> sl) & 1u);
      parity ^= 1u << sl;
    }
  }
  cp_async_wait<0>();
  fence_after();
  // ---- epilogue: thread = accumulator row (lane quadrant of the warp), 64 columns per warp half ----------
  if (nk >= 1 && !(g.debug & 1)) {
    const int q = warp & 3, half = warp >> 2;
    const long long m = m0 + 32 * q + lane;
    float* C = g.C ? g.C + (long long)blockIdx.z * g.part_stride : nullptr;
    float bias = 0.f;
    if (EPI == EPI_BIAS_ACT_TC && m < g.M) bias = __ldg(g.bias + m);
#pragma unroll 1
    for (int cb = 0; cb < 2; ++cb) {
      float v[32];
      const int c0 = 64 * half + 32 * cb;
      ld_32x32b_x32(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)c0, v);
      const long long nb0 = n0 + c0;
      if (m < g.M && nb0 < g.N) {
        const int ncols = (int)((g.N - nb0) < 32 ? (g.N - nb0) : 32);
        // phase 1: everything the epilogue READS (all loads in flight before the first store: the output
        // arrays may alias nothing, but the compiler cannot know that across the stores)
        if (EPI == EPI_MUL_DACT_TC) {
          const float* ap = g.aux + m + g.ldaux * nb0;
          float a[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) a[j] = j < ncols ? __ldg(ap + (long long)j * g.ldaux) : 0.f;
          with_act_tc(g.act, [&](auto A) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= decltype(A)::d(a[j]);
          });
        } else if (EPI == EPI_BIAS_ACT_TC) {
          with_act_tc(g.act, [&](auto A) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = decltype(A)::f(v[j] + bias);
          });
        } else if (g.accumulate) {
          const float* cp = C + m + g.ldc * nb0;
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += j < ncols ? cp[(long long)j * g.ldc] : 0.f;
        }
        // phase 2: stores (lanes = consecutive rows m: every store instruction is one coalesced line)
        if (C) {
          float* cp = C + m + g.ldc * nb0;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < ncols) cp[(long long)j * g.ldc] = v[j];
        }
        if (EPI != EPI_PLAIN_TC && g.Chi) {
          __nv_bfloat16* hp = g.Chi + m + g.ldc * nb0;
          __nv_bfloat16* lp = g.Clo + m + g.ldc * nb0;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (j < ncols) {
              const __nv_bfloat16 h = __float2bfloat16_rn(v[j]);
              hp[(long long)j * g.ldc] = h;
              lp[(long long)j * g.ldc] = __float2bfloat16_rn(v[j] - __bfloat162float(h));
            }
          }
        }
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

template <bool A_MN, bool B_MN, int EPI>
static cudaError_t launch_tc(const TcGemmArgs& g0, int ksplit, cudaStream_t st) {
  TcGemmArgs g = g0;
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e = getenv("BFX_GEMM_DEBUG");
      dbg = e ? atoi(e) : 0;
    }
    g.debug = dbg;
  }
  auto kern = k_tc_gemm<A_MN, B_MN, EPI>;
  const size_t smem = (size_t)TSTAGES * STAGE_BYTES;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid((g.M + TBM - 1) / TBM, (g.N + TBN - 1) / TBN, ksplit);
  kern<<<grid, 256, smem, st>>>(g);
  return cudaGetLastError();
}

cudaError_t tc_gemm_forward(const TcGemmArgs& g, cudaStream_t st) { return launch_tc<true, false, EPI_BIAS_ACT_TC>(g, 1, st); }
cudaError_t tc_gemm_dx(const TcGemmArgs& g, cudaStream_t st) { return launch_tc<false, false, EPI_MUL_DACT_TC>(g, 1, st); }
cudaError_t tc_gemm_dw(const TcGemmArgs& g, int ksplit, cudaStream_t st) { return launch_tc<true, true, EPI_PLAIN_TC>(g, ksplit, st); }

}  // namespace bfx
