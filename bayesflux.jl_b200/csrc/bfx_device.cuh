// Device building blocks of the small-P ("theta resident in shared memory") regime:
// Philox noise, keyed batch permutation, block reductions, the three register-tiled
// shared-memory GEMM shapes, the fused MLP forward / likelihood / backward pass over one
// minibatch tile, and the element-wise parameter passes the samplers are made of.
//
// One CTA owns one chain.  theta and the gradient stay in shared memory for the whole
// launch; only minibatch rows (L2-resident data set) and recorded samples touch HBM.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "bfx_common.cuh"

namespace bfx {

#define BFX_DEV __device__ __forceinline__

// Phase timing of the fused kernels (build with -DBFX_TIMING; off by default: TICK compiles to nothing).
// Thread 0 of every CTA adds the cycles since its previous TICK to counter i.
#ifdef BFX_TIMING
static __device__ unsigned long long g_bfx_timing[24];
__shared__ long long s_bfx_tprev;
#define BFX_TICK(i)                                                        \
  do {                                                                     \
    if (threadIdx.x == 0) {                                                \
      const long long now__ = clock64();                                   \
      atomicAdd(&g_bfx_timing[i], (unsigned long long)(now__ - s_bfx_tprev)); \
      s_bfx_tprev = now__;                                                 \
    }                                                                      \
  } while (0)
#else
#define BFX_TICK(i) do { } while (0)
#endif

// ---------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011).  counter = (quad index, step, chain, draw slot)
// ---------------------------------------------------------------------------------
BFX_DEV uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    unsigned hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    unsigned hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

// the same rounds with the key schedule given (rk[2r], rk[2r+1] = key + r * (W0, W1))
BFX_DEV uint4 philox4x32_10_rk(uint4 ctr, const unsigned (&rk)[20]) {
  const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    unsigned hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    unsigned hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ rk[2 * r], lo1, hi0 ^ ctr.w ^ rk[2 * r + 1], lo0);
  }
  return ctr;
}

// two uint32 -> two independent N(0,1) (Box-Muller)
BFX_DEV float2 box_muller(unsigned a, unsigned b) {
  float u1 = ((float)a + 1.0f) * 2.3283064365386963e-10f;  // (0,1]
  float u2 = (float)b * 2.3283064365386963e-10f;           // [0,1)
  // straight-line MUFU code (no denormal / slow-path branches: u1 >= 2^-32)
  float l2, r, s, c;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u1));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-1.3862943611198906f * l2));  // sqrt(-2 ln u1)
  const float ang = 6.283185307179586f * u2;
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(ang));
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(ang));
  return make_float2(r * c, r * s);
}

BFX_DEV void fdivmod(const FastDiv& f, unsigned long long n, unsigned long long& q, unsigned long long& r) {
  if (f.d == 1ull) {
    q = n;
    r = 0ull;
  } else if (f.m != 0ull && n < (1ull << 32)) {
    q = __umul64hi(n, f.m);
    r = n - q * f.d;
  } else {
    q = n / f.d;
    r = n - q * f.d;
  }
}

struct NoiseKey {
  unsigned long long seed;
  unsigned chain;  // global chain id
};

// the 4 standard normals of flat parameters 4q..4q+3 for (chain, step, slot)
BFX_DEV float4 normal_quad(const NoiseKey& k, unsigned long long step, unsigned slot, unsigned q) {
  uint4 ctr = make_uint4(q, (unsigned)step, k.chain, (slot << 24) | (unsigned)((step >> 32) & 0xFFFFFFu));
  uint2 key = make_uint2((unsigned)k.seed, (unsigned)(k.seed >> 32));
  uint4 r = philox4x32_10(ctr, key);
  float2 a = box_muller(r.x, r.y), b = box_muller(r.z, r.w);
  return make_float4(a.x, a.y, b.x, b.y);
}

// the uniform(0,1) of (chain, step) used by the MH accept tests
BFX_DEV float uniform_draw(const NoiseKey& k, unsigned long long step) {
  uint4 ctr = make_uint4(0xFFFFFFFFu, (unsigned)step, k.chain, 0xFF000000u | (unsigned)((step >> 32) & 0xFFFFFFu));
  uint2 key = make_uint2((unsigned)k.seed, (unsigned)(k.seed >> 32));
  uint4 r = philox4x32_10(ctr, key);
  return ((float)(r.x >> 8) + 0.5f) * 5.9604644775390625e-08f;  // (0,1)
}

// ---------------------------------------------------------------------------------
// per-epoch pseudo-random permutation of [0,n): keyed Feistel network on the next
// power of two, cycle-walked into range.  Replaces MLUtils' randperm(n) per epoch
// (src/inference/mcmc/abstract.jl:102-103): every observation exactly once per epoch.
// ---------------------------------------------------------------------------------
BFX_DEV unsigned mix32(unsigned x) {
  x ^= x >> 16;
  x *= 0x7FEB352Du;
  x ^= x >> 15;
  x *= 0x846CA68Bu;
  x ^= x >> 16;
  return x;
}

BFX_DEV unsigned long long perm_index(unsigned long long pos, unsigned long long n, unsigned k0, unsigned k1) {
  if (n <= 1) return 0;
  if (n <= 0x80000000ull) {
    // every data set the API accepts (n < 2^31): the same network in 32-bit arithmetic (bit-identical results,
    // half the instructions of the 64-bit form below)
    const unsigned nn = (unsigned)n;
    int bits = 32 - __clz((int)(nn - 1u));
    if (bits < 2) bits = 2;
    const int lb = bits >> 1, rb = bits - lb;  // left (high) / right (low) widths
    const unsigned lmask = (1u << lb) - 1u, rmask = (1u << rb) - 1u;
    unsigned v = (unsigned)pos;
    do {
      unsigned L = (v >> rb) & lmask, R = v & rmask;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if ((r & 1) == 0) L = (L ^ mix32(R * 0x9E3779B1u + k0 + (unsigned)r * 0x85EBCA6Bu)) & lmask;
        else R = (R ^ mix32(L * 0xC2B2AE35u + k1 + (unsigned)r * 0x27D4EB2Fu)) & rmask;
      }
      v = (L << rb) | R;
    } while (v >= nn);
    return v;
  }
  int bits = 64 - __clzll((long long)(n - 1));
  const int lb = bits >> 1, rb = bits - lb;
  const unsigned long long lmask = (1ull << lb) - 1, rmask = (1ull << rb) - 1;
  unsigned long long v = pos;
  do {
    unsigned long long L = (v >> rb) & lmask, R = v & rmask;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      // alternate which half is updated so that widths stay fixed (unbalanced Feistel)
      if ((r & 1) == 0)
        L = (L ^ (unsigned long long)mix32((unsigned)R * 0x9E3779B1u + k0 + (unsigned)r * 0x85EBCA6Bu + (unsigned)(R >> 32))) & lmask;
      else
        R = (R ^ (unsigned long long)mix32((unsigned)L * 0xC2B2AE35u + k1 + (unsigned)r * 0x27D4EB2Fu)) & rmask;
    }
    v = (L << rb) | R;
  } while (v >= n);
  return v;
}

// ---------------------------------------------------------------------------------
// block reductions (warp shuffle + one smem hop).  `red` holds >= 32*N floats.
// ---------------------------------------------------------------------------------
template <int N>
BFX_DEV void warp_sum(float (&v)[N]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
}

template <int NT, int N>
BFX_DEV void block_sum(float (&v)[N], float* red) {
  warp_sum<N>(v);
  if (NT > 32) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();  // protect `red` from a previous use
    if (lane == 0)
#pragma unroll
      for (int i = 0; i < N; ++i) red[w * N + i] = v[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < N; ++i) {
      float s = 0.f;
#pragma unroll
      for (int ww = 0; ww < NT / 32; ++ww) s += red[ww * N + i];
      v[i] = s;
    }
  }
}

// ---------------------------------------------------------------------------------
// activations.  Accurate tanhf/expf forms: Flux's tanh_fast/sigmoid_fast are within a
// few ulp of these, far inside the 1e-4 parity tolerance.
// ---------------------------------------------------------------------------------
BFX_DEV float act_apply(int a, float z) {
  switch (a) {
    case A_RELU: return fmaxf(z, 0.f);
    case A_TANH: return tanhf(z);
    case A_SIGMOID: return 1.f / (1.f + expf(-z));
    default: return z;
  }
}
BFX_DEV float act_deriv_from_out(int a, float o) {
  switch (a) {
    case A_RELU: return o > 0.f ? 1.f : 0.f;
    case A_TANH: return 1.f - o * o;
    case A_SIGMOID: return o * (1.f - o);
    default: return 1.f;
  }
}

BFX_DEV float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
// per-chain state in HBM: L2-coherent load (a chain may be continued by a CTA on another SM, whose L1 could
// still hold the lines of an earlier chunk)
BFX_DEV float4 gld4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
BFX_DEV void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

#define BFX_FMA4x4(ACC, WV, AV) \
  ACC[0][0] = fmaf(WV.x, AV.x, ACC[0][0]); \
  ACC[0][1] = fmaf(WV.x, AV.y, ACC[0][1]); \
  ACC[0][2] = fmaf(WV.x, AV.z, ACC[0][2]); \
  ACC[0][3] = fmaf(WV.x, AV.w, ACC[0][3]); \
  ACC[1][0] = fmaf(WV.y, AV.x, ACC[1][0]); \
  ACC[1][1] = fmaf(WV.y, AV.y, ACC[1][1]); \
  ACC[1][2] = fmaf(WV.y, AV.z, ACC[1][2]); \
  ACC[1][3] = fmaf(WV.y, AV.w, ACC[1][3]); \
  ACC[2][0] = fmaf(WV.z, AV.x, ACC[2][0]); \
  ACC[2][1] = fmaf(WV.z, AV.y, ACC[2][1]); \
  ACC[2][2] = fmaf(WV.z, AV.z, ACC[2][2]); \
  ACC[2][3] = fmaf(WV.z, AV.w, ACC[2][3]); \
  ACC[3][0] = fmaf(WV.w, AV.x, ACC[3][0]); \
  ACC[3][1] = fmaf(WV.w, AV.y, ACC[3][1]); \
  ACC[3][2] = fmaf(WV.w, AV.z, ACC[3][2]); \
  ACC[3][3] = fmaf(WV.w, AV.w, ACC[3][3]);

// ---------------------------------------------------------------------------------
// GEMM shape 1 (forward):  Z[m][n] = act( sum_k W[k*ldw + m] * A[k*lda + n] + bias[m] )
//   M = Mpad rows (multiple of 4), N = BT batch columns, K = fan-in.
//   W is the reference's column-major weight (m contiguous), A is feature-major.
//   If `accumulate`, Z is read first (used to add the recurrent term) and no bias is added.
// ---------------------------------------------------------------------------------
template <int NT, int BT>
BFX_DEV void gemm_fwd(const float* __restrict__ W, int ldw, const float* __restrict__ A, int lda,
                      const float* __restrict__ bias, float* __restrict__ Z, int ldz, int Mpad, int K, int act,
                      bool accumulate = false) {
  constexpr int NTN = BT / 4;
  const int ntiles = (Mpad >> 2) * NTN;
  for (int t = threadIdx.x; t < ntiles; t += NT) {
    const int tm = t / NTN, tn = t - tm * NTN;
    const float* wp = W + 4 * tm;
    const float* ap = A + 4 * tn;
    float* zp = Z + (4 * tm) * ldz + 4 * tn;
    float acc[4][4];
    if (accumulate) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 z = ld4(zp + i * ldz);
        acc[i][0] = z.x; acc[i][1] = z.y; acc[i][2] = z.z; acc[i][3] = z.w;
      }
    } else {
      float4 b = ld4(bias + 4 * tm);
      float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = bb[i];
    }
    int k = 0;
    for (; k + 4 <= K; k += 4) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float4 w = ld4(wp + (k + kk) * ldw);
        float4 a = ld4(ap + (k + kk) * lda);
        BFX_FMA4x4(acc, w, a);
      }
    }
    for (; k < K; ++k) {
      float4 w = ld4(wp + k * ldw);
      float4 a = ld4(ap + k * lda);
      BFX_FMA4x4(acc, w, a);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 o;
      o.x = act_apply(act, acc[i][0]);
      o.y = act_apply(act, acc[i][1]);
      o.z = act_apply(act, acc[i][2]);
      o.w = act_apply(act, acc[i][3]);
      st4(zp + i * ldz, o);
    }
  }
}

// ---------------------------------------------------------------------------------
// GEMM shape 2 (input gradient):  C[m][n] = sum_k W[m*ldw + k] * D[k*ldd + n]
//   m = fan-in index (Mpad rows), k = fan-out index (Kpad, multiple of 4, zero padded),
//   result is multiplied by act'(Aprev[m][n]) and written IN PLACE over Aprev, or
//   (add_to) accumulated into `Out` without masking (used for recurrent dh).
// ---------------------------------------------------------------------------------
template <int NT, int BT>
BFX_DEV void gemm_dx(const float* __restrict__ W, int ldw, const float* __restrict__ D, int ldd, float* Out,
                     int ldo, int Mpad, int Kpad, int act_prev, bool add_to = false) {
  constexpr int NTN = BT / 4;
  const int ntiles = (Mpad >> 2) * NTN;
  for (int t = threadIdx.x; t < ntiles; t += NT) {
    const int tm = t / NTN, tn = t - tm * NTN;
    const float* wp = W + (4 * tm) * ldw;
    const float* dp = D + 4 * tn;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k = 0; k < Kpad; k += 4) {
      float4 w0 = ld4(wp + k), w1 = ld4(wp + ldw + k), w2 = ld4(wp + 2 * ldw + k), w3 = ld4(wp + 3 * ldw + k);
      float4 d;
      float4 wc;
      d = ld4(dp + (k + 0) * ldd); wc = make_float4(w0.x, w1.x, w2.x, w3.x); BFX_FMA4x4(acc, wc, d);
      d = ld4(dp + (k + 1) * ldd); wc = make_float4(w0.y, w1.y, w2.y, w3.y); BFX_FMA4x4(acc, wc, d);
      d = ld4(dp + (k + 2) * ldd); wc = make_float4(w0.z, w1.z, w2.z, w3.z); BFX_FMA4x4(acc, wc, d);
      d = ld4(dp + (k + 3) * ldd); wc = make_float4(w0.w, w1.w, w2.w, w3.w); BFX_FMA4x4(acc, wc, d);
    }
    float* op = Out + (4 * tm) * ldo + 4 * tn;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 a = ld4(op + i * ldo);
      float4 o;
      if (add_to) {
        o = make_float4(a.x + acc[i][0], a.y + acc[i][1], a.z + acc[i][2], a.w + acc[i][3]);
      } else {
        o.x = acc[i][0] * act_deriv_from_out(act_prev, a.x);
        o.y = acc[i][1] * act_deriv_from_out(act_prev, a.y);
        o.z = acc[i][2] * act_deriv_from_out(act_prev, a.z);
        o.w = acc[i][3] * act_deriv_from_out(act_prev, a.w);
      }
      st4(op + i * ldo, o);
    }
  }
}

// ---------------------------------------------------------------------------------
// GEMM shape 3 (weight gradient):  G[n*ldg + m] += sum_k D[m*ldd + k] * A[n*lda + k]
//   m = fan-out (Mpad), n = fan-in (Npad, multiple of 4), k = batch column (BT).
//   The n rows of a thread's tile are strided by Npad/4 so the 128-bit loads of A are
//   bank-conflict free.  G uses the weight layout (m contiguous).
// ---------------------------------------------------------------------------------
template <int NT, int BT>
BFX_DEV void gemm_dw(const float* __restrict__ D, int ldd, const float* __restrict__ A, int lda, float* __restrict__ G,
                     int ldg, int Mpad, int Npad) {
  const int ntn = Npad >> 2;
  const int ntiles = (Mpad >> 2) * ntn;
  for (int t = threadIdx.x; t < ntiles; t += NT) {
    const int tm = t / ntn, tn = t - tm * ntn;
    const float* dp = D + (4 * tm) * ldd;
    const float* ap = A + tn * lda;
    const int as = ntn * lda;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 2
    for (int k = 0; k < BT; k += 4) {
      float4 d0 = ld4(dp + k), d1 = ld4(dp + ldd + k), d2 = ld4(dp + 2 * ldd + k), d3 = ld4(dp + 3 * ldd + k);
      float4 a0 = ld4(ap + k), a1 = ld4(ap + as + k), a2 = ld4(ap + 2 * as + k), a3 = ld4(ap + 3 * as + k);
      float4 wc, ac;
      wc = make_float4(d0.x, d1.x, d2.x, d3.x); ac = make_float4(a0.x, a1.x, a2.x, a3.x); BFX_FMA4x4(acc, wc, ac);
      wc = make_float4(d0.y, d1.y, d2.y, d3.y); ac = make_float4(a0.y, a1.y, a2.y, a3.y); BFX_FMA4x4(acc, wc, ac);
      wc = make_float4(d0.z, d1.z, d2.z, d3.z); ac = make_float4(a0.z, a1.z, a2.z, a3.z); BFX_FMA4x4(acc, wc, ac);
      wc = make_float4(d0.w, d1.w, d2.w, d3.w); ac = make_float4(a0.w, a1.w, a2.w, a3.w); BFX_FMA4x4(acc, wc, ac);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float* gp = G + (tn + j * ntn) * ldg + 4 * tm;
      float4 g = ld4(gp);
      g.x += acc[0][j]; g.y += acc[1][j]; g.z += acc[2][j]; g.w += acc[3][j];
      st4(gp, g);
    }
  }
}

// bias gradient: G[m] += sum_k D[m*ldd + k], one warp per row
template <int NT, int BT>
BFX_DEV void rowsum_acc(const float* __restrict__ D, int ldd, float* __restrict__ G, int M) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int m = w; m < M; m += NT / 32) {
    float s = 0.f;
#pragma unroll
    for (int k = lane; k < BT; k += 32) s += D[m * ldd + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) G[m] += s;
  }
}

// ---------------------------------------------------------------------------------
// Parameter passes.  Inside the engine every per-parameter vector (theta, g, momentum, Minv, ...)
// uses the PADDED layout of ModelDev (S floats, a multiple of 4), in shared memory and in HBM
// alike, so that every element-wise pass is a float4 stream over k = 0, 4, 8, ... < S.
// Padding slots hold 0 and are kept at 0 through the 4-bit validity mask of each quad
// (M.pad_mask).  The reference's flat order (the ABI layout) is only touched where values
// enter or leave the engine, through the tables M.flat_pad / M.pad_flat.
// ---------------------------------------------------------------------------------
template <int NT, class F>
BFX_DEV void for_each_quad(const ModelDev& M, F f) {
  for (int k = 4 * threadIdx.x; k < M.S; k += 4 * NT) f(k);
}

// the same over S floats given directly; UNROLL: S is a compile-time constant at the call site (shape-specialised
// kernels, TcFixed), so the loop has a known trip count
template <int NT, bool UNROLL, class F>
BFX_DEV void for_each_quad_n(int S, F f) {
  if (UNROLL) {
#pragma unroll
    for (int k0 = 0; k0 < S; k0 += 4 * NT) {
      const int k = k0 + 4 * (int)threadIdx.x;
      if (k0 + 4 * NT <= S || k < S) f(k);
    }
  } else {
    for (int k = 4 * threadIdx.x; k < S; k += 4 * NT) f(k);
  }
}

// call f(flat index J, padded index k) for every parameter (boundary code only: one table load each)
template <int NT, class F>
BFX_DEV void for_each_flat(const ModelDev& M, F f) {
  // four table entries in flight per thread: the loop is otherwise a chain of dependent L2 round trips
  int J = threadIdx.x;
  for (; J + 3 * NT < M.P; J += 4 * NT) {
    const int k0 = __ldg(M.flat_pad + J), k1 = __ldg(M.flat_pad + J + NT);
    const int k2 = __ldg(M.flat_pad + J + 2 * NT), k3 = __ldg(M.flat_pad + J + 3 * NT);
    f(J, k0);
    f(J + NT, k1);
    f(J + 2 * NT, k2);
    f(J + 3 * NT, k3);
  }
  for (; J < M.P; J += NT) f(J, __ldg(M.flat_pad + J));
}

BFX_DEV float4 mask4(float4 v, unsigned m) {
  return make_float4((m & 1u) ? v.x : 0.f, (m & 2u) ? v.y : 0.f, (m & 4u) ? v.z : 0.f, (m & 8u) ? v.w : 0.f);
}

// the 4 standard normals of padded quad k: Philox draws of (key, step, slot, k/4), or the injected
// flat-indexed normals inj[J] (parity entry)
BFX_DEV float4 noise_quad(const ModelDev& M, const NoiseKey& key, unsigned long long step, unsigned slot,
                          const float* __restrict__ inj, int k) {
  if (inj) {
    const int4 J = __ldg(reinterpret_cast<const int4*>(M.pad_flat + k));
    return make_float4(J.x >= 0 ? inj[J.x] : 0.f, J.y >= 0 ? inj[J.y] : 0.f, J.z >= 0 ? inj[J.z] : 0.f,
                       J.w >= 0 ? inj[J.w] : 0.f);
  }
  if (M.philox_rk_set) {
    const uint4 ctr = make_uint4((unsigned)(k >> 2), (unsigned)step, key.chain,
                                 (slot << 24) | (unsigned)((step >> 32) & 0xFFFFFFu));
    const uint4 r = philox4x32_10_rk(ctr, M.philox_rk);
    const float2 a = box_muller(r.x, r.y), b = box_muller(r.z, r.w);
    return make_float4(a.x, a.y, b.x, b.y);
  }
  return normal_quad(key, step, slot, (unsigned)(k >> 2));
}

// ---------------------------------------------------------------------------------
// per-CTA working set
// ---------------------------------------------------------------------------------
template <int BT>
struct TileCfg {
  static constexpr int LDB = BT + 4;  // (BT+4)/4 odd for BT % 8 == 0
};

struct ChainSmem {
  float* th;    // [S]   padded parameters
  float* g;     // [S]   padded gradient
  float* act;   // activation buffers of one batch tile (FP32 path only)
  float* ys;    // [BT]  targets of the tile
  int* idx;     // [BT]  observation indices of the tile (-1 = masked)
  float* red;   // [32*4] reduction scratch
  double* dacc; // [4]   double accumulators (thread 0)
  unsigned char* mask;  // [S/4] validity bits of each padded quad (copy of M.pad_mask)
  unsigned short* qflat;  // [S/4] flat index of element 0 of each padded quad (P < 65536 in this regime)
  float* scr;   // per-CTA HBM scratch of the recurrent path (per-step records) or nullptr
  unsigned char* tc;  // tensor-core region (128-byte aligned) or nullptr
};

// floats of the CTA working set that precede the tensor-core region
template <int BT>
__host__ __device__ inline size_t chain_smem_floats(const ModelDev& M) {
  return (size_t)2 * M.S + (M.tc.enabled ? 0 : M.act_total) + BT /*ys*/ + BT /*idx*/ + 32 * 4 /*red*/ +
         (size_t)((M.S / 4 + 15) & ~15) / 4 /*mask*/ + (size_t)((M.S / 4 + 15) & ~15) / 2 /*qflat*/;
}

template <int BT>
__host__ __device__ inline size_t chain_smem_bytes(const ModelDev& M) {
  size_t b = chain_smem_floats<BT>(M) * 4 + 4 * sizeof(double) + 16;
  if (M.tc.enabled) b = ((b + 127) & ~(size_t)127) + (size_t)M.tc.total_bytes + 128;
  return b;
}

// (SP: a TcFixed shape makes every offset below a literal, so the pointers cost no registers of their own)
template <int BT, class SP = TcDyn>
BFX_DEV ChainSmem carve_smem(const ModelDev& M, unsigned char* base) {
  ChainSmem s;
  double* d = reinterpret_cast<double*>(base);
  s.dacc = d;
  const int S = SP::S(M);
  float* f = reinterpret_cast<float*>(d + 4);
  s.th = f; f += S;
  s.g = f; f += S;
  s.act = f; f += (SP::fixed || M.tc.enabled ? 0 : M.act_total);
  s.ys = f; f += BT;
  s.idx = reinterpret_cast<int*>(f); f += BT;
  s.red = f; f += 32 * 4;
  s.mask = reinterpret_cast<unsigned char*>(f); f += ((S / 4 + 15) & ~15) / 4;
  s.qflat = reinterpret_cast<unsigned short*>(f); f += ((S / 4 + 15) & ~15) / 2;
  s.tc = nullptr;
  s.scr = nullptr;
  if (M.tc.enabled) {
    // aligned by pointer arithmetic (not through an integer): the compiler keeps the shared address space and
    // emits LDS / STS with 32-bit addresses for every tile access instead of generic LD / ST
    const unsigned mis = (unsigned)__cvta_generic_to_shared(f) & 127u;
    s.tc = reinterpret_cast<unsigned char*>(f) + ((128u - mis) & 127u);
    // (opaque to the optimiser: with literal offsets everywhere else it would otherwise RECOMPUTE this alignment at
    // every use of a tile pointer, 800 instructions per chain-step, instead of holding the one pointer)
    unsigned tc_off = (unsigned)__cvta_generic_to_shared(s.tc);
    asm volatile("" : "+r"(tc_off));
    s.tc = static_cast<unsigned char*>(__cvta_shared_to_generic(tc_off));
  }
  return s;
}

// where a chain's minibatch comes from
struct BatchSrc {
  const long long* idx;  // explicit indices (B per chain) or nullptr
  long long n;           // data set size
  long long first;       // explicit==nullptr: position of the first element inside the epoch permutation
  unsigned k0, k1;       // permutation key of this (chain, epoch)
  int shuffle;           // 0: identity permutation
  int full;              // 1: batch = [0, n) in order (full-data evaluations)
};

BFX_DEV long long batch_obs(const BatchSrc& b, long long j) {
  if (b.full) return j;
  if (b.idx) return b.idx[j];
  long long pos = b.first + j;
  return b.shuffle ? (long long)perm_index((unsigned long long)pos, (unsigned long long)b.n, b.k0, b.k1) : pos;
}

// ---------------------------------------------------------------------------------
// network pass over one batch tile (MLP).  On return (want_grad) the gradient of
//   sum_b dl[b] * yhat[b]   has been ACCUMULATED into s.g.  Returns per-thread partial
//   likelihood sums in `sums` (to be block-reduced by the caller).
// Likelihood arithmetic: src/likelihoods/feedforward.jl:30-43 (Normal), :86-97 (TDist);
// closed-form deltas: SURVEY.md section 8 rows a8/a9.
// ---------------------------------------------------------------------------------
struct LikeCoef {
  float inv_sigma;
  float nb;  // num_batches (scales the delta; src/model/BNN.jl:55)
};

// indices and targets of the tile
template <int NT, int BT>
BFX_DEV void stage_targets(const ChainSmem& s, const float* __restrict__ y, const BatchSrc& bs, long long tile0,
                           int nvalid) {
  for (int b = threadIdx.x; b < BT; b += NT) {
    long long o = (b < nvalid) ? batch_obs(bs, tile0 + b) : -1;
    s.idx[b] = (int)o;
    s.ys[b] = (o >= 0) ? __ldg(y + o) : 0.f;
  }
}

// Dense layers l0 .. nlayers-1 forward; the input of layer l0 is already in its activation buffer
template <int NT, int BT>
BFX_DEV void dense_forward(const ModelDev& M, const ChainSmem& s, int l0) {
  constexpr int LDB = TileCfg<BT>::LDB;
  for (int l = l0; l < M.nlayers; ++l) {
    const LayerDev& L = M.L[l];
    gemm_fwd<NT, BT>(s.th + L.w_s, L.ldw, s.act + M.act_off[l], LDB, s.th + L.b_s, s.act + M.act_off[l + 1], LDB,
                     L.nout_pad, L.nin, L.act);
    __syncthreads();
  }
}

// likelihood epilogue on the network output; leaves dl[b] (scaled by num_batches) in the output buffer
template <int NT, int BT>
BFX_DEV void like_epilogue(const ModelDev& M, const ChainSmem& s, int nvalid, const LikeCoef& lc, bool want_grad,
                           float (&sums)[2], float* __restrict__ yhat_tile = nullptr) {
  constexpr int LDB = TileCfg<BT>::LDB;
  float* out = s.act + M.act_off[M.nlayers];
  const int out_act = M.L[M.nlayers - 1].act;
  for (int b = threadIdx.x; b < BT; b += NT) {
    float dl = 0.f;
    if (b < nvalid) {
      float yh = out[b];
      if (yhat_tile) yhat_tile[b] = yh;  // predictive helpers: the network output itself
      float z = (s.ys[b] - yh) * lc.inv_sigma;
      if (M.like_kind == LK_NORMAL) {
        sums[0] += z * z;
        dl = lc.nb * z * lc.inv_sigma;
      } else {
        float zz = z * z;
        float w = (M.nu + 1.f) * z / (M.nu + zz);
        sums[0] += log1pf(zz / M.nu);
        sums[1] += w * z;
        dl = lc.nb * w * lc.inv_sigma;
      }
      dl *= act_deriv_from_out(out_act, yh);
    }
    if (want_grad) {
      out[b] = dl;
      out[LDB + b] = 0.f;
      out[2 * LDB + b] = 0.f;
      out[3 * LDB + b] = 0.f;
    }
  }
}

// Dense layers nlayers-1 .. l0 backward.  For l0 > 0 the raw gradient w.r.t. the input of layer l0 (not
// multiplied by any activation derivative) is left in that layer's input buffer.
template <int NT, int BT>
BFX_DEV void dense_backward(const ModelDev& M, const ChainSmem& s, int l0) {
  constexpr int LDB = TileCfg<BT>::LDB;
  for (int l = M.nlayers - 1; l >= l0; --l) {
    const LayerDev& L = M.L[l];
    const float* D = s.act + M.act_off[l + 1];
    float* A = s.act + M.act_off[l];
    gemm_dw<NT, BT>(D, LDB, A, LDB, s.g + L.w_s, L.ldw, L.nout_pad, L.nin_pad);
    rowsum_acc<NT, BT>(D, LDB, s.g + L.b_s, L.nout);
    if (l > 0) {
      __syncthreads();
      gemm_dx<NT, BT>(s.th + L.w_s, L.ldw, D, LDB, A, LDB, L.nin_pad, L.nout_pad,
                      l > l0 ? M.L[l - 1].act : A_IDENTITY);
      __syncthreads();
    }
  }
}

template <int NT, int BT>
BFX_DEV void mlp_tile(const ModelDev& M, const ChainSmem& s, const float* __restrict__ x,
                      const float* __restrict__ y, const BatchSrc& bs, long long tile0, int nvalid,
                      const LikeCoef& lc, bool want_grad, float (&sums)[2], float* yhat_tile = nullptr) {
  constexpr int LDB = TileCfg<BT>::LDB;
  const int d = M.d_in;
  stage_targets<NT, BT>(s, y, bs, tile0, nvalid);
  __syncthreads();
  {
    float* X = s.act + M.act_off[0];
    const int dpad = M.act_rows[0];
    for (int e = threadIdx.x; e < dpad * BT; e += NT) {
      int b = e / dpad, f = e - b * dpad;
      int o = s.idx[b];
      X[f * LDB + b] = (o >= 0 && f < d) ? __ldg(x + (long long)o * d + f) : 0.f;
    }
  }
  __syncthreads();
  dense_forward<NT, BT>(M, s, 0);
  like_epilogue<NT, BT>(M, s, nvalid, lc, want_grad, sums, yhat_tile);
  if (!want_grad) return;
  __syncthreads();
  dense_backward<NT, BT>(M, s, 0);
}

// ---------------------------------------------------------------------------------
// Seq-to-one networks: Chain(RNN | LSTM, Dense..., Dense(., 1)) on x of shape T x d x n.
// src/likelihoods/seq_to_one.jl:31-44,89-100: the chain is applied to every time slice with the
// recurrent state carried over (zero initial state, src/layers/recurrent.jl:30,60,64) and only the
// last output is used -- the Dense head therefore only matters at t = T.  Flux cells (third party):
//   RNNCell   h' = act(Wi x + Wh h + b)
//   LSTMCell  g = Wi x + Wh h + b, rows [i; f; c~; o];  c' = sig(f) c + sig(i) tanh(c~);  h' = sig(o) tanh(c')
// Backward = BPTT, hand-derived (oracle/bfx_oracle.py net_backward).  The per-step records the backward
// needs (gate outputs, c_t, h_t) go to a per-CTA scratch in HBM (L2 resident); weights, gradient and the
// working tiles stay in shared memory.
// ---------------------------------------------------------------------------------
BFX_DEV float sigm(float z) { return 1.f / (1.f + expf(-z)); }

template <int NT, int BT>
BFX_DEV void rec_stage_x(const ModelDev& M, const ChainSmem& s, const float* __restrict__ x, int t) {
  constexpr int LDB = TileCfg<BT>::LDB;
  float* X = s.act + M.act_off[0];
  const int dpad = M.act_rows[0], d = M.d_in, T = M.T;
  for (int e = threadIdx.x; e < dpad * BT; e += NT) {
    int f = e / BT, b = e - f * BT;
    int o = s.idx[b];
    X[f * LDB + b] = (o >= 0 && f < d) ? __ldg(x + ((long long)o * d + f) * T + t) : 0.f;
  }
}

// ---------------------------------------------------------------------------------
// Forward-only pass of an LSTM(d, H <= 64) + Dense(H, 1) chain over a tile of 64 sequences: the full-data log
// posteriors of the GGMC / HMC acceptance tests (ggmc.jl:147,176; hmc.jl:135-136) and bfx_loglikeprior, which need no
// gradient and no per-step records.  A thread owns two units x eight sequences and computes ALL FOUR gate rows of its
// units (64 accumulators: every weight pair and every h value loaded from shared memory feeds 8 / 8 FMAs instead of
// the 4 / 4 of the 16-sequence gradient tile), so the gate nonlinearities and the cell update are thread-local: c stays
// in registers over the whole sequence, h ping-pongs between two buffers, no gate buffer and one barrier per time step.
// Same accumulation order as rec_tile (bias, input features, recurrent features), so the values agree with it.
// Workspace: the gradient buffer s.g (unused by a value-only evaluation).
// ---------------------------------------------------------------------------------
template <int NT>
BFX_DEV bool lstm_value_ok(const ModelDev& M) {
  if (M.rec_layer != 0 || M.nlayers != 2) return false;
  const LayerDev& L = M.L[0];
  const LayerDev& LO = M.L[1];
  if (L.kind != L_LSTM || (L.H & 1) || L.H > 64 || (L.H / 2) * 8 > NT || LO.kind != L_DENSE || LO.nout != 1) return false;
  return 2 * 64 * 68 + L.nin_pad * 68 + 128 <= M.S;
}

template <int NT>
BFX_DEV void lstm_value_tile64(const ModelDev& M, const ChainSmem& s, const float* __restrict__ x,
                               const float* __restrict__ y, const BatchSrc& bs, long long tile0, int nvalid,
                               const LikeCoef& lc, float (&sums)[2]) {
  constexpr int LDH = 68;
  const LayerDev& L = M.L[0];
  const LayerDev& LO = M.L[1];
  const int H = L.H, T = M.T, d = M.d_in, dpad = L.nin_pad;
  float* ws = s.g;
  float* Hb0 = ws;
  float* Hb1 = ws + 64 * LDH;
  float* X = ws + 2 * 64 * LDH;  // [dpad][LDH] features of the current time step
  float* ys = X + dpad * LDH;    // [64]
  int* idx = reinterpret_cast<int*>(ys + 64);
  const int tid = threadIdx.x;
  for (int b = tid; b < 64; b += NT) {
    const long long o = (b < nvalid) ? batch_obs(bs, tile0 + b) : -1;
    idx[b] = (int)o;
    ys[b] = (o >= 0) ? __ldg(y + o) : 0.f;
  }
  for (int e = tid; e < 64 * LDH; e += NT) Hb0[e] = 0.f;
  __syncthreads();
  const bool active = tid < (H / 2) * 8;
  const int r = 2 * (tid >> 3), c0 = 8 * (tid & 7);  // first unit, first sequence of this thread
  float c[2][8];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int j = 0; j < 8; ++j) c[u][j] = 0.f;
  const float* Wi = s.th + L.w_s + r;
  const float* Wh = s.th + L.wh_s + r;
  for (int t = 0; t < T; ++t) {
    for (int e = tid; e < dpad * 64; e += NT) {
      const int f = e >> 6, b = e & 63;
      const int o = idx[b];
      X[f * LDH + b] = (o >= 0 && f < d) ? __ldg(x + ((long long)o * d + f) * T + t) : 0.f;
    }
    __syncthreads();
    const float* hp = (t & 1) ? Hb1 : Hb0;
    float* hn = (t & 1) ? Hb0 : Hb1;
    if (active) {
      float acc[4][2][8];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float2 bb = *reinterpret_cast<const float2*>(s.th + L.b_s + g * H + r);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[g][0][j] = bb.x;
          acc[g][1][j] = bb.y;
        }
      }
      auto kstep = [&](const float* wcol, const float* arow) {
        const float4 a0 = ld4(arow), a1 = ld4(arow + 4);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float2 w = *reinterpret_cast<const float2*>(wcol + g * H);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            acc[g][0][j] = fmaf(w.x, a[j], acc[g][0][j]);
            acc[g][1][j] = fmaf(w.y, a[j], acc[g][1][j]);
          }
        }
      };
      for (int k = 0; k < d; ++k) kstep(Wi + k * L.ldw, X + k * LDH + c0);
#pragma unroll 4
      for (int k = 0; k < H; ++k) kstep(Wh + k * L.ldh, hp + k * LDH + c0);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        float hv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float ig = sigm(acc[0][u][j]), fg = sigm(acc[1][u][j]);
          const float cg = tanhf(acc[2][u][j]), og = sigm(acc[3][u][j]);
          c[u][j] = fg * c[u][j] + ig * cg;
          hv[j] = og * tanhf(c[u][j]);
        }
        st4(hn + (r + u) * LDH + c0, make_float4(hv[0], hv[1], hv[2], hv[3]));
        st4(hn + (r + u) * LDH + c0 + 4, make_float4(hv[4], hv[5], hv[6], hv[7]));
      }
    }
    __syncthreads();
  }
  // Dense(H, 1) head on h_T and the likelihood sums (like_epilogue, value part)
  const float* hT = (T & 1) ? Hb1 : Hb0;
  for (int b = tid; b < nvalid; b += NT) {
    float yh = s.th[LO.b_s];
    for (int k = 0; k < H; ++k) yh = fmaf(s.th[LO.w_s + k * LO.ldw], hT[k * LDH + b], yh);
    yh = act_apply(LO.act, yh);
    const float z = (ys[b] - yh) * lc.inv_sigma;
    if (M.like_kind == LK_NORMAL) {
      sums[0] += z * z;
    } else {
      const float zz = z * z;
      sums[0] += log1pf(zz / M.nu);
      sums[1] += (M.nu + 1.f) * z / (M.nu + zz) * z;
    }
  }
}

template <int NT, int BT>
BFX_DEV void rec_tile(const ModelDev& M, const ChainSmem& s, const float* __restrict__ x,
                      const float* __restrict__ y, const BatchSrc& bs, long long tile0, int nvalid,
                      const LikeCoef& lc, bool want_grad, float (&sums)[2], float* yhat_tile = nullptr) {
  constexpr int LDB = TileCfg<BT>::LDB;
  const LayerDev& L = M.L[0];
  const int H = L.H, T = M.T;
  const bool lstm = L.kind == L_LSTM;
  const int GR = lstm ? 4 * H : H;   // gate rows
  const int GRp = L.nout_pad, Hp4 = M.act_rows[1];
  const int RR = lstm ? 6 * H : H;   // rows of one per-step record in the scratch
  float* Hc = s.act + M.act_off[1];  // h_t (input of the Dense head); dh during the backward sweep
  float* Hp = s.act + M.rec_hp;      // h_{t-1}
  float* Cc = s.act + M.rec_c;       // c_t
  float* DC = s.act + M.rec_dc;      // dL/dc flowing to the past
  float* G = s.act + M.rec_g;        // gate pre-activations / dz
  float* scr = s.scr;
  stage_targets<NT, BT>(s, y, bs, tile0, nvalid);
  for (int e = threadIdx.x; e < Hp4 * LDB; e += NT) {
    Hc[e] = 0.f;
    Cc[e] = 0.f;
  }
  __syncthreads();
  // ---- forward over time -----------------------------------------------------------------
  for (int t = 0; t < T; ++t) {
    for (int e = threadIdx.x; e < Hp4 * LDB; e += NT) Hp[e] = Hc[e];
    rec_stage_x<NT, BT>(M, s, x, t);
    __syncthreads();
    gemm_fwd<NT, BT>(s.th + L.w_s, L.ldw, s.act + M.act_off[0], LDB, s.th + L.b_s, G, LDB, GRp, L.nin, A_IDENTITY);
    __syncthreads();
    gemm_fwd<NT, BT>(s.th + L.wh_s, L.ldh, Hp, LDB, nullptr, G, LDB, GRp, H, lstm ? A_IDENTITY : L.act, true);
    __syncthreads();
    float* rec = scr + (long long)t * RR * BT;
    for (int e = threadIdx.x; e < H * BT; e += NT) {
      const int r = e / BT, b = e - r * BT;
      if (lstm) {
        const float ig = sigm(G[r * LDB + b]), fg = sigm(G[(H + r) * LDB + b]);
        const float cg = tanhf(G[(2 * H + r) * LDB + b]), og = sigm(G[(3 * H + r) * LDB + b]);
        const float c = fg * Cc[r * LDB + b] + ig * cg;
        const float h = og * tanhf(c);
        Cc[r * LDB + b] = c;
        Hc[r * LDB + b] = h;
        if (want_grad) {
          rec[r * BT + b] = ig;
          rec[(H + r) * BT + b] = fg;
          rec[(2 * H + r) * BT + b] = cg;
          rec[(3 * H + r) * BT + b] = og;
          rec[(4 * H + r) * BT + b] = c;
          rec[(5 * H + r) * BT + b] = h;
        }
      } else {
        const float h = G[r * LDB + b];
        Hc[r * LDB + b] = h;
        if (want_grad) rec[r * BT + b] = h;
      }
    }
    __syncthreads();
  }
  // ---- Dense head on h_T, likelihood, head backward (leaves dL/dh_T in Hc) ---------------------
  dense_forward<NT, BT>(M, s, 1);
  like_epilogue<NT, BT>(M, s, nvalid, lc, want_grad, sums, yhat_tile);
  if (!want_grad) return;
  __syncthreads();
  dense_backward<NT, BT>(M, s, 1);
  for (int e = threadIdx.x; e < Hp4 * LDB; e += NT) DC[e] = 0.f;
  for (int e = threadIdx.x; e < GRp * LDB; e += NT) G[e] = 0.f;
  __syncthreads();
  // ---- backward through time -------------------------------------------------------------------
  for (int t = T - 1; t >= 0; --t) {
    const float* rec = scr + (long long)t * RR * BT;
    const float* recp = scr + (long long)(t - 1) * RR * BT;
    for (int e = threadIdx.x; e < H * BT; e += NT) {
      const int r = e / BT, b = e - r * BT;
      const float dh = Hc[r * LDB + b];
      if (lstm) {
        const float ig = rec[r * BT + b], fg = rec[(H + r) * BT + b], cg = rec[(2 * H + r) * BT + b];
        const float og = rec[(3 * H + r) * BT + b], c = rec[(4 * H + r) * BT + b];
        const float cprev = t > 0 ? recp[(4 * H + r) * BT + b] : 0.f;
        Hp[r * LDB + b] = t > 0 ? recp[(5 * H + r) * BT + b] : 0.f;
        const float tc = tanhf(c);
        const float dgo = dh * tc;
        const float dc = dh * og * (1.f - tc * tc) + DC[r * LDB + b];
        G[r * LDB + b] = dc * cg * ig * (1.f - ig);
        G[(H + r) * LDB + b] = dc * cprev * fg * (1.f - fg);
        G[(2 * H + r) * LDB + b] = dc * ig * (1.f - cg * cg);
        G[(3 * H + r) * LDB + b] = dgo * og * (1.f - og);
        DC[r * LDB + b] = dc * fg;
      } else {
        const float h = rec[r * BT + b];
        Hp[r * LDB + b] = t > 0 ? recp[r * BT + b] : 0.f;
        G[r * LDB + b] = dh * act_deriv_from_out(L.act, h);
      }
    }
    rec_stage_x<NT, BT>(M, s, x, t);
    __syncthreads();
    gemm_dw<NT, BT>(G, LDB, s.act + M.act_off[0], LDB, s.g + L.w_s, L.ldw, GRp, L.nin_pad);
    gemm_dw<NT, BT>(G, LDB, Hp, LDB, s.g + L.wh_s, L.ldh, GRp, Hp4);
    rowsum_acc<NT, BT>(G, LDB, s.g + L.b_s, GR);
    if (t > 0) gemm_dx<NT, BT>(s.th + L.wh_s, L.ldh, G, LDB, Hc, LDB, Hp4, GRp, A_IDENTITY);
    __syncthreads();
  }
}

}  // namespace bfx
