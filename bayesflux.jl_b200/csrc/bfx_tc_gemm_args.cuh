// Plain argument structs of the stream-regime tcgen05 GEMM (bfx_tc_gemm.cuh), shared with the host orchestration
// (bfx_stream.cu) without pulling the kernel templates in.
#pragma once
#include <cstdint>

#include <cuda_bf16.h>

namespace bfx {
namespace tcg {

constexpr int BM = 128;   // accumulator rows per CTA
constexpr int BN = 256;   // accumulator columns (TMEM allocation; a launch may use 224 of them, Args::bn)
constexpr int BK = 64;    // K per stage = one 128-byte swizzle span of BF16
constexpr int TILE_A = BM * BK * 2;  // one half (hi or lo) of the A stage, 16 KB
constexpr int NTHREADS = 512;        // producer warp, MMA warp, allocator warp; all sixteen warps run the epilogue
constexpr int ROWSLOTS = 4;          // bias-gradient row-sum partials per tile (one per 64-column group)

enum { EPI_BIAS_ACT = 0, EPI_MUL_DACT = 1, EPI_PLAIN = 2 };

struct Args {
  int M, N, K;
  float* C;              // FP32 output or nullptr.  EPI_PLAIN: the split-K partial / accumulate target
  long long ldc;
  int store_split;       // 1: the BF16 hi / lo halves of the output go out through the tensor maps mChi / mClo
  const float* bias;     // EPI_BIAS_ACT
  int act;
  const __nv_bfloat16* aux_hi;  // EPI_MUL_DACT: multiply by act'(aux[m + ldaux * n]), aux = hi + lo
  const __nv_bfloat16* aux_lo;
  long long ldaux;
  // ReLU fast path: the forward epilogue leaves one sign bit per output (word (n / 32) * M + m, bit n % 32), the
  // input-gradient epilogue reads it back instead of 32 strided BF16 loads per word
  uint32_t* mask_out;        // EPI_BIAS_ACT with act == A_RELU (nullptr: not written)
  const uint32_t* mask_in;   // EPI_MUL_DACT with act == A_RELU (nullptr: use aux_hi)
  float* rowpart;        // EPI_MUL_DACT: per-row sums of the stored tile, slot (2 * blockIdx.y + half) * M + m
  int accumulate;        // EPI_PLAIN, single split: C += acc
  int kchunk;            // split-K: blockIdx.z handles k in [z * kchunk, ...), multiple of BK; writes C + z * part_stride
  long long part_stride;
  int bn;                // accumulator columns per tile: 256, or 224 when B is K-major (8 192 columns = 37 tiles x 2 row
                         // pairs = 148 CTAs, one full wave, instead of 128); 0 = 256
  long long* timing;     // BFX_GEMM_TIMING builds: [CTA][8] clock64 marks
};

// accumulator columns per tile of a cta_group::2 launch whose B operand is K-major (forward, input gradient): 224
// when that needs fewer waves of the 74 CTA pairs than 256 (every consumer of per-tile partials must use this too)
inline int tile_cols_kmajor_b(long long M, long long N) {
#ifdef BFX_GEMM_NO224
  return BN;
#endif
  const long long mt2 = (M + 2 * BM - 1) / (2 * BM);
  auto cost = [&](long long w) { const long long cl = mt2 * ((N + w - 1) / w); return ((cl + 73) / 74) * w; };
  return cost(224) < cost(BN) ? 224 : BN;
}

struct Operand {
  const __nv_bfloat16* hi;
  const __nv_bfloat16* lo;
  long long ld;
};


struct Output {
  __nv_bfloat16* hi;  // nullptr: the BF16 halves are not written
  __nv_bfloat16* lo;
};


}  // namespace tcg
}  // namespace bfx
