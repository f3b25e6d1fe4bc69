// tcgen05 GEMM of the large-P ("stream") regime, second generation: C[M x N] = A * B with every operand stored in
// HBM as two BF16 arrays (hi, lo = the 3-term split of the FP32 value, see bfx_tc.cuh), column-major.
//
//   * a CTA PAIR (cluster 2 x 1 x 1, tcgen05 cta_group::2) owns a 256 x 256 output tile: each CTA holds 128 rows of A
//     and 128 of the 256 rows of B in its own shared memory, the leader CTA issues M = 256, N = 256, K = 16 MMAs that
//     read both CTAs' operands, every CTA keeps its 128 x 256 FP32 accumulator in its own TMEM (256 columns);
//   * operands arrive by TMA (cp.async.bulk.tensor.2d, 128-byte swizzle) into a 3-stage ring of 64 KB per CTA
//     (K = 64 per stage: A hi, A lo, B hi, B lo), producer = one lane of warp 0, full / empty mbarriers, the empty
//     barriers are released by tcgen05.commit (multicast to both CTAs);
//   * one lane of warp 1 of the leader issues hi*hi + hi*lo + lo*hi per k-step into the same accumulator;
//   * all eight warps run the epilogue from tcgen05.ld.32x32b (thread = accumulator row): bias + activation with
//     BF16 hi / lo stores, the act' multiply with the bias-gradient row sums, or the split-K partial store.
//
// Both operand majornesses are read straight from the natural column-major arrays (K-major: box 64 k x 128 rows;
// MN-major: two boxes of 64 rows x 64 k), so the three GEMMs of a Dense layer (forward, input gradient, weight
// gradient) need no transposed copies.  The single-CTA variant (CG = 1, 128 x 256 tile, 2 stages of 96 KB) is kept
// for bring-up and for devices that cannot co-schedule the pair.
//
// Reference arithmetic replaced: Flux Dense forward (called at src/likelihoods/feedforward.jl:35,91) and its
// Zygote pullbacks (src/model/BNN.jl:66) for layers wide enough to be real GEMMs (configs[2]: 512 x 512 x 8192).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstring>

#include "bfx_tc.cuh"
#include "bfx_tc_gemm_args.cuh"

namespace bfx {
namespace tcg {

template <int CG>
struct Cfg {
  static constexpr int BROWS = BN / CG;               // rows of B staged by one CTA
  static constexpr int TILE_B = BROWS * BK * 2;
  static constexpr int STAGE = 2 * TILE_A + 2 * TILE_B;
  static constexpr int STAGES = CG == 2 ? 3 : 2;
  static constexpr int SMEM = STAGES * STAGE + 1024;  // + slack for the 1024-byte alignment of the swizzle atoms
};

BFX_DEV uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
BFX_DEV void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
BFX_DEV void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}
// 2-D tiled TMA load, completion on `bar` (for CG = 2 the barrier of the pair's leader CTA)
template <int CG>
BFX_DEV void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  const uint32_t d = tc::smem_u32(dst);
  if (CG == 2) {
    const uint32_t b = tc::smem_u32(bar) & 0xFEFFFFFFu;  // peer bit cleared: the transaction bytes land on CTA 0's barrier
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(d), "l"(map), "r"(b), "r"(c0), "r"(c1)
        : "memory");
  } else {
    const uint32_t b = tc::smem_u32(bar);
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(d), "l"(map), "r"(b), "r"(c0), "r"(c1)
        : "memory");
  }
}
// shared-memory matrix descriptor, SWIZZLE_128B (layout type 2), descriptor version 1
BFX_DEV uint64_t desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
template <int CG>
BFX_DEV void mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  if (CG == 2) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
  }
}
// all MMAs issued so far by this thread arrive on `bar` when they complete (CG = 2: on both CTAs' copies of it)
template <int CG>
BFX_DEV void mma_commit(uint64_t* bar) {
  if (CG == 2) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            tc::smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
  } else {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc::smem_u32(bar))
                 : "memory");
  }
}
template <int CG>
BFX_DEV void tmem_alloc(uint32_t* slot, uint32_t cols) {
  if (CG == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(slot)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(slot)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
}
template <int CG>
BFX_DEV void tmem_free(uint32_t addr, uint32_t cols) {
  if (CG == 2)
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
  else
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}

// 32 consecutive accumulator columns of this thread's row: asynchronous issue, then a wait that owns the registers
// (so that no use of them can be scheduled above it)
BFX_DEV void ld_row32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
BFX_DEV void ld_row32_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

BFX_DEV void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(tc::smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}

#ifdef BFX_GEMM_TIMING
#define BFX_GT(i)                                                                                                       \
  do {                                                                                                                  \
    if (g.timing && threadIdx.x == 32 * ((i) == 1 || (i) == 2 ? 1 : 0))                                                 \
      g.timing[((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 + (i)] = clock64();                \
  } while (0)
#else
#define BFX_GT(i) do { } while (0)
#endif

// D = F32, A = B = BF16, M = 128 * CG, N = bn
template <int CG>
BFX_DEV uint32_t idesc_of(int a_mn, int b_mn, int bn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(bn >> 3) << 17) | ((uint32_t)((BM * CG) >> 4) << 24);
}

// stage `rows` rows (mn) x 64 k of one operand half at `dst`
template <int CG, bool MN_MAJOR>
BFX_DEV void load_operand(unsigned char* dst, const CUtensorMap* map, uint64_t* bar, int mn0, int k0, int rows) {
  if (!MN_MAJOR) {
    tma_load_2d<CG>(dst, map, bar, k0, mn0);  // box {64 k, rows}: row r at r * 128 bytes, 16-byte chunks XOR (r % 8)
  } else {
    for (int j = 0; j < rows / 64; ++j)  // boxes {64 mn, 64 k}: k-row r at r * 128 bytes inside its 8 KB box
      tma_load_2d<CG>(dst + j * 8192, map, bar, mn0 + 64 * j, k0);
  }
}

template <int CG, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(NTHREADS, 1)
k_tc_gemm2(const __grid_constant__ CUtensorMap mAhi, const __grid_constant__ CUtensorMap mAlo,
           const __grid_constant__ CUtensorMap mBhi, const __grid_constant__ CUtensorMap mBlo,
           const __grid_constant__ CUtensorMap mChi, const __grid_constant__ CUtensorMap mClo,
           const __grid_constant__ CUtensorMap mChi2, const __grid_constant__ CUtensorMap mClo2, const Args g) {
  using C = Cfg<CG>;
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[C::STAGES], bar_empty[C::STAGES], bar_acc;
  __shared__ uint32_t tmem_slot;
  unsigned char* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  const int mtile = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int m0 = mtile * (BM * CG) + (int)rank * BM;         // first accumulator row of this CTA
  const int bn = g.bn ? g.bn : BN;                            // accumulator columns of this launch (256 or 224)
  const int brows = bn / CG;                                  // rows of B staged by one CTA
  const int n0 = (int)blockIdx.y * bn;                        // first accumulator column
  const int nb0 = n0 + (int)rank * brows;                     // first B row staged by this CTA
  const int kbeg = (int)blockIdx.z * g.kchunk;
  int kend = kbeg + g.kchunk;
  if (kend > g.K) kend = g.K;
  const int nkb = (kend - kbeg + BK - 1) / BK;
  BFX_GT(0);

  if (tid == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      tc::mbar_init(&bar_full[s], 1);
      tc::mbar_init(&bar_empty[s], 1);
    }
    tc::mbar_init(&bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mAhi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mAlo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mBhi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mBlo) : "memory");
  }
  if (CG == 2) cluster_sync_all(); else __syncthreads();   // barrier inits visible (to the peer) before anything arrives
  // The producer starts fetching at once; the TMEM allocation (warp 2) is awaited by warps 1-7 only (named barrier),
  // warp 0 picks the address up after the accumulator barrier, which orders it behind the allocation.
#ifdef BFX_GEMM_NO_EARLY
  if (warp == 2) tmem_alloc<CG>(&tmem_slot, (uint32_t)BN);
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
#else
  if (warp != 0) {
    if (warp == 2) tmem_alloc<CG>(&tmem_slot, (uint32_t)BN);
    tc::fence_before();
    asm volatile("bar.sync 3, %0;" ::"n"(NTHREADS - 32) : "memory");
    tc::fence_after();
  }
#endif
  BFX_GT(7);

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (tc::elect_one()) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % C::STAGES;
        const uint32_t ph = (uint32_t)(kb / C::STAGES) & 1u;
        tc::mbar_wait(&bar_empty[s], ph ^ 1u);
        if (rank == 0) mbar_expect_tx(&bar_full[s], (uint32_t)((2 * TILE_A + 2 * brows * BK * 2) * CG));
        unsigned char* st = smem + s * C::STAGE;
        const int k0 = kbeg + kb * BK;
        load_operand<CG, A_MN>(st, &mAhi, &bar_full[s], m0, k0, BM);
        load_operand<CG, A_MN>(st + TILE_A, &mAlo, &bar_full[s], m0, k0, BM);
        load_operand<CG, B_MN>(st + 2 * TILE_A, &mBhi, &bar_full[s], nb0, k0, brows);
        load_operand<CG, B_MN>(st + 2 * TILE_A + C::TILE_B, &mBlo, &bar_full[s], nb0, k0, brows);
      }
    }
    __syncwarp();
  } else if (warp == 1 && rank == 0) {
    // ---------------- MMA issuer (leader CTA) ----------------
    const uint32_t idesc = idesc_of<CG>(A_MN ? 1 : 0, B_MN ? 1 : 0, bn);
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&tmem_slot);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % C::STAGES;
      const uint32_t ph = (uint32_t)(kb / C::STAGES) & 1u;
      tc::mbar_wait(&bar_full[s], ph);
      tc::fence_after();
      if (kb == 0) BFX_GT(1);
      if (tc::elect_one()) {
        const uint32_t ahi = tc::smem_u32(smem + s * C::STAGE), alo = ahi + TILE_A;
        const uint32_t bhi = ahi + 2 * TILE_A, blo = bhi + C::TILE_B;
#pragma unroll
        for (int ks = 0; ks < BK / 16; ++ks) {
          // K-major: 16 k = 32 bytes inside the 128-byte row, 8-row groups 1024 bytes apart (SBO), LBO unused
          // MN-major: 16 k = two 8-row groups of 1024 bytes (SBO), 64-row (mn) groups 8192 bytes apart (LBO)
          const uint64_t dah = A_MN ? desc_sw128(ahi + ks * 2048, 8192, 1024) : desc_sw128(ahi + ks * 32, 16, 1024);
          const uint64_t dal = A_MN ? desc_sw128(alo + ks * 2048, 8192, 1024) : desc_sw128(alo + ks * 32, 16, 1024);
          const uint64_t dbh = B_MN ? desc_sw128(bhi + ks * 2048, 8192, 1024) : desc_sw128(bhi + ks * 32, 16, 1024);
          const uint64_t dbl = B_MN ? desc_sw128(blo + ks * 2048, 8192, 1024) : desc_sw128(blo + ks * 32, 16, 1024);
          mma_bf16<CG>(tmem, dah, dbh, idesc, (kb | ks) ? 1u : 0u);
          mma_bf16<CG>(tmem, dah, dbl, idesc, 1u);
          mma_bf16<CG>(tmem, dal, dbh, idesc, 1u);
        }
        mma_commit<CG>(&bar_empty[s]);               // stage s may be refilled once these MMAs have read it
        if (kb == nkb - 1) mma_commit<CG>(&bar_acc); // accumulator complete
      }
      __syncwarp();
    }
    BFX_GT(2);
  }

  // ---------------- epilogue: thread = accumulator row (lane quadrant of the warp), 128 columns per warp half ------
  if (nkb >= 1) {
    tc::mbar_wait(&bar_acc, 0u);
    tc::fence_after();
    BFX_GT(3);
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&tmem_slot);
    // sixteen warps: lane quadrant q, column group grp = 64 columns (two 32-column blocks), two groups per 128-column
    // half (the unit of the TMA stores).  The epilogue is instruction-bound (about 11 k warp instructions per CTA):
    // eight warps ran it at IPC 1.2, sixteen hide each other's latencies.
    const int q = warp & 3, grp = warp >> 2, half = grp >> 1, sub = grp & 1;
    const int nblk = half == 0 ? 4 : (bn - 128) / 32;   // 32-column blocks of this 128-column half (bn = 224: 4 + 3)
    const int cb0 = 2 * sub, cb1 = (cb0 + 2 < nblk) ? cb0 + 2 : nblk;  // this warp's blocks [cb0, cb1)
    const int mloc = 32 * q + lane;
    const long long m = (long long)m0 + mloc;
    float* Cp = g.C ? g.C + (long long)blockIdx.z * g.part_stride : nullptr;
    float bias = 0.f;
    if (EPI == EPI_BIAS_ACT && m < g.M) bias = __ldg(g.bias + m);
    float rsum = 0.f;
    // the operand ring is idle now: the BF16 halves of this half's 128 x 128 block are staged there as [n][m]
    // (m contiguous, 256 bytes per column) and leave through one TMA store each
    unsigned char* stage_hi = smem + half * 65536;
    unsigned char* stage_lo = stage_hi + 32768;
    const bool split = EPI != EPI_PLAIN && g.store_split;
    auto process = [&](int cb, const uint32_t (&r)[32]) {
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      const int c0 = 128 * half + 32 * cb;
      const long long nb = (long long)n0 + c0;
      const bool inb = m < g.M && nb < g.N;
      const int ncols = inb ? (int)((g.N - nb) < 32 ? (g.N - nb) : 32) : 0;
      if (EPI == EPI_MUL_DACT) {
        if (g.act == A_RELU && g.mask_in) {
          const uint32_t w = inb ? __ldg(g.mask_in + (nb >> 5) * g.M + m) : 0u;
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = (w >> j) & 1u ? v[j] : 0.f;
        } else {
          float a[32];
          const __nv_bfloat16* ah = g.aux_hi + m + g.ldaux * nb;
          const __nv_bfloat16* al = g.aux_lo + m + g.ldaux * nb;
          if (g.act == A_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) a[j] = j < ncols ? __bfloat162float(ah[(long long)j * g.ldaux]) : 0.f;
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              a[j] = j < ncols ? __bfloat162float(ah[(long long)j * g.ldaux]) + __bfloat162float(al[(long long)j * g.ldaux]) : 0.f;
          }
          tc::with_act(g.act, [&](auto A) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = j < ncols ? v[j] * decltype(A)::d(a[j]) : 0.f;
          });
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) rsum += v[j];
      } else if (EPI == EPI_BIAS_ACT) {
        tc::with_act(g.act, [&](auto A) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = decltype(A)::f(v[j] + bias);
        });
        if (g.act == A_RELU && g.mask_out && inb) {
          uint32_t w = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) w |= (v[j] > 0.f ? 1u : 0u) << j;
          g.mask_out[(nb >> 5) * g.M + m] = w;
        }
      } else if (g.accumulate && inb) {
        const float* cp = Cp + m + g.ldc * nb;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += j < ncols ? cp[(long long)j * g.ldc] : 0.f;
      }
      if (Cp && inb) {  // lanes = consecutive rows m: every store instruction is one full line
        float* cp = Cp + m + g.ldc * nb;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < ncols) cp[(long long)j * g.ldc] = v[j];
      }
      if (split) {
        __nv_bfloat16* sh = reinterpret_cast<__nv_bfloat16*>(stage_hi) + (32 * cb) * 128 + mloc;
        __nv_bfloat16* sl = reinterpret_cast<__nv_bfloat16*>(stage_lo) + (32 * cb) * 128 + mloc;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const __nv_bfloat16 h = __float2bfloat16_rn(v[j]);
          sh[j * 128] = h;
          sl[j * 128] = __float2bfloat16_rn(v[j] - __bfloat162float(h));
        }
      }
    };
    const uint32_t tbase = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(128 * half);
    if (EPI == EPI_PLAIN) {  // FP32 stores only: the plain loop keeps 32 stores in flight per load
#pragma unroll 1
      for (int cb = cb0; cb < cb1; ++cb) {
        uint32_t ra[32];
        ld_row32_issue(tbase + 32 * cb, ra);
        ld_row32_wait(ra);
        process(cb, ra);
      }
    } else if (cb0 < cb1) {
      uint32_t ra[32], rb[32];
      const bool two = cb0 + 1 < cb1;
      ld_row32_issue(tbase + 32 * cb0, ra);
      ld_row32_wait(ra);
      if (two) ld_row32_issue(tbase + 32 * (cb0 + 1), rb);   // the second block streams out of TMEM while the first is processed
      process(cb0, ra);
      if (two) {
        ld_row32_wait(rb);
        process(cb0 + 1, rb);
      }
    }
    if (EPI == EPI_MUL_DACT && g.rowpart && m < g.M) g.rowpart[(long long)(ROWSLOTS * blockIdx.y + grp) * g.M + m] = rsum;
    BFX_GT(4);
    if (split) {
      tc::fence_async_smem();  // generic-proxy stores -> visible to the TMA engine
      asm volatile("bar.sync %0, 256;" ::"r"(1 + half) : "memory");
      if (q == 0 && sub == 0 && lane == 0 && m0 < g.M && n0 + 128 * half < g.N) {  // rows / columns beyond M / N are clipped by the TMA
        // (the second half's box is bn - 128 columns wide: it must not reach into the next tile)
        tma_store_2d(half ? &mChi2 : &mChi, stage_hi, m0, n0 + 128 * half);
        tma_store_2d(half ? &mClo2 : &mClo, stage_lo, m0, n0 + 128 * half);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
    }
    BFX_GT(5);
  }
  tc::fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) tmem_free<CG>(*reinterpret_cast<volatile uint32_t*>(&tmem_slot), (uint32_t)BN);
  BFX_GT(6);
}

// ---------------- host side ----------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && p) fn = (EncodeTiledFn)p;
  }
  return fn;
}

// one operand half: element (mn, k) at base[mn * ld + k] (K-major) or base[k * ld + mn] (MN-major), extents MN x K
inline cudaError_t make_map(CUtensorMap* map, const __nv_bfloat16* base, bool mn_major, long long MN, long long K, long long ld,
                            int rows_per_box) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return cudaErrorNotSupported;
  cuuint64_t dims[2], strides[1];
  cuuint32_t box[2], estr[2] = {1, 1};
  if (!mn_major) {
    dims[0] = (cuuint64_t)K; dims[1] = (cuuint64_t)MN;
    box[0] = BK; box[1] = (cuuint32_t)rows_per_box;
  } else {
    dims[0] = (cuuint64_t)MN; dims[1] = (cuuint64_t)K;
    box[0] = 64; box[1] = BK;
  }
  strides[0] = (cuuint64_t)ld * 2;
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// output halves: element (m, n) at base[m + ldc * n], box 128 rows x 128 columns, no swizzle
inline cudaError_t make_store_map(CUtensorMap* map, __nv_bfloat16* base, long long M, long long N, long long ldc,
                                  int box_cols = 128) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return cudaErrorNotSupported;
  cuuint64_t dims[2] = {(cuuint64_t)M, (cuuint64_t)N}, strides[1] = {(cuuint64_t)ldc * 2};
  cuuint32_t box[2] = {128, (cuuint32_t)box_cols}, estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

template <int CG, bool A_MN, bool B_MN, int EPI>
inline cudaError_t launch(const Args& g0, const Operand& A, const Operand& B, const Output& O, int ksplit, cudaStream_t st) {
  using C = Cfg<CG>;
  Args g = g0;
  CUtensorMap mAhi, mAlo, mBhi, mBlo, mChi, mClo, mChi2, mClo2;
  cudaError_t e;
  memset(&mChi, 0, sizeof mChi);
  memset(&mClo, 0, sizeof mClo);
  memset(&mChi2, 0, sizeof mChi2);
  memset(&mClo2, 0, sizeof mClo2);
  // tile width: 224 columns when that fills the machine better (B K-major only: its rows come as one box)
  const int bn = (CG == 2 && !B_MN && g.bn != BN) ? tile_cols_kmajor_b(g.M, g.N) : BN;
  g.bn = bn;
  const int brows = bn / CG;
  g.store_split = 0;
  if (EPI != EPI_PLAIN && O.hi) {
    if ((e = make_store_map(&mChi, O.hi, g.M, g.N, g.ldc)) != cudaSuccess) return e;
    if ((e = make_store_map(&mClo, O.lo, g.M, g.N, g.ldc)) != cudaSuccess) return e;
    if ((e = make_store_map(&mChi2, O.hi, g.M, g.N, g.ldc, bn - 128)) != cudaSuccess) return e;
    if ((e = make_store_map(&mClo2, O.lo, g.M, g.N, g.ldc, bn - 128)) != cudaSuccess) return e;
    g.store_split = 1;
  }
  if ((e = make_map(&mAhi, A.hi, A_MN, g.M, g.K, A.ld, BM)) != cudaSuccess) return e;
  if ((e = make_map(&mAlo, A.lo, A_MN, g.M, g.K, A.ld, BM)) != cudaSuccess) return e;
  if ((e = make_map(&mBhi, B.hi, B_MN, g.N, g.K, B.ld, brows)) != cudaSuccess) return e;
  if ((e = make_map(&mBlo, B.lo, B_MN, g.N, g.K, B.ld, brows)) != cudaSuccess) return e;
  auto kern = k_tc_gemm2<CG, A_MN, B_MN, EPI>;
  {  // once per device and template instance (function attributes are per context)
    static unsigned long long done = 0ull;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!((done >> (dev & 63)) & 1ull)) {
      if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM)) != cudaSuccess) return e;
      done |= 1ull << (dev & 63);
    }
  }
  cudaLaunchConfig_t cfg{};
  const int mt = (g.M + BM * CG - 1) / (BM * CG);
  cfg.gridDim = dim3((unsigned)(mt * CG), (unsigned)((g.N + bn - 1) / bn), (unsigned)ksplit);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = C::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, mAhi, mAlo, mBhi, mBlo, mChi, mClo, mChi2, mClo2, g);
}

}  // namespace tcg
}  // namespace bfx
