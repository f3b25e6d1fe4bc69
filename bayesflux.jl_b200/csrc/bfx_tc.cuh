// Tensor-core (tcgen05 / TMEM) evaluation of the log posterior and its gradient for MLP chains.
//
// Same contract as chain_eval (bfx_chain.cuh): one CTA = one chain, on return s.g holds
// grad[num_batches*loglike + logprior] in the padded parameter layout.  What changes is where
// the contractions run: every hidden Dense layer is a 64 x Hout x Hin GEMM per batch tile
// (batch rows x fan-out x fan-in), issued as tcgen05.mma.cta_group::1.kind::f16 with M = 64 by
// one thread, operands in shared memory (no-swizzle core-block tiles, hi/lo BF16 halves),
// accumulators in TMEM.  FP32 accuracy is kept by the 3-term split  a*b ~ ah*bh + ah*bl + al*bh
// (|error| ~ 2^-17 per product, FP32 accumulate), which holds the 1e-4 gradient bar of the
// FP32 gate path.  The 1-wide output layer, the likelihood and all activation work run in the
// epilogues on the TMEM fragments.
//
// Reference arithmetic: Flux Dense forward called at src/likelihoods/feedforward.jl:35,91, the
// Zygote pullbacks of it (src/model/BNN.jl:66), likelihood deltas of SURVEY.md rows a8/a9.
// Descriptor / fragment conventions were pinned on a B200 by tools/umma_probe*.cu
// (profiles/r01_umma_probe_conventions.log).
#pragma once
#include <cuda_bf16.h>

#include <type_traits>

#include "bfx_chain.cuh"

namespace bfx {
namespace tc {

BFX_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, SWIZZLE_NONE, descriptor version 1
BFX_DEV uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
// instruction descriptor: D = F32, A = B = BF16, M = 64
BFX_DEV uint32_t make_idesc(int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
}
BFX_DEV void mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
// one lane of a converged warp.  MMA issue sits in warp-uniform code behind this predicate: inside a divergent
// `tid == 0` region the compiler wraps every tcgen05.mma in an elect/broadcast loop (~100 cycles per MMA).
BFX_DEV bool elect_one() {
  uint32_t e;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(e));
  return e != 0;
}
BFX_DEV void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
BFX_DEV void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
BFX_DEV void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
#ifdef BFX_WATCHDOG  // bring-up builds (tools/gemm_probe.cu): a barrier that never completes traps instead of hanging the GPU
  const long long t0 = clock64();
#endif
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
#ifdef BFX_WATCHDOG
    if (!ok && clock64() - t0 > 4000000000LL) {
      printf("mbar_wait watchdog: block (%d,%d,%d) thread %d bar %x parity %u\n", blockIdx.x, blockIdx.y, blockIdx.z,
             threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
#endif
  } while (!ok);
}
BFX_DEV void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
BFX_DEV void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
BFX_DEV void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 16-byte asynchronous copy global -> shared; src_bytes = 0 writes zeros
BFX_DEV void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes)
               : "memory");
}
BFX_DEV void cp_async4(void* smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes)
               : "memory");
}
BFX_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
BFX_DEV void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 16 lanes x 32 columns of a warp's lane quadrant: thread t gets, for j = 0..3,
//   v[4j+0], v[4j+1] = (row t/4,     cols 8j + 2(t%4) + {0,1})
//   v[4j+2], v[4j+3] = (row t/4 + 8, same columns)
BFX_DEV void ld_16x256b_x4(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
BFX_DEV void ld_16x256b_x1(uint32_t taddr, float (&v)[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];\n\ttcgen05.wait::ld.sync.aligned;"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}

// BF16 split of two adjacent values: hi word = {bf16(a), bf16(b)}, lo word = the residuals
BFX_DEV float bf16lo_to_f(uint32_t w) { return __uint_as_float(w << 16); }
BFX_DEV float bf16hi_to_f(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }
BFX_DEV uint32_t pack_bf16x2(float lo_elem, float hi_elem) {  // one cvt.rn.bf16x2.f32
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}
BFX_DEV void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16x2(a, b);
  lo = pack_bf16x2(a - bf16lo_to_f(hi), b - bf16hi_to_f(hi));
}

// activation dispatch hoisted out of the element loops
template <int A>
struct ActT {
  static BFX_DEV float f(float z) {
    if (A == A_RELU) return fmaxf(z, 0.f);
    if (A == A_TANH) return tanhf(z);
    if (A == A_SIGMOID) return 1.f / (1.f + expf(-z));
    return z;
  }
  static BFX_DEV float d(float o) {  // derivative from the output
    if (A == A_RELU) return o > 0.f ? 1.f : 0.f;
    if (A == A_TANH) return 1.f - o * o;
    if (A == A_SIGMOID) return o * (1.f - o);
    return 1.f;
  }
};
template <class F>
BFX_DEV void with_act(int act, F f) {
  switch (act) {
    case A_RELU: f(ActT<A_RELU>{}); break;
    case A_TANH: f(ActT<A_TANH>{}); break;
    case A_SIGMOID: f(ActT<A_SIGMOID>{}); break;
    default: f(ActT<A_IDENTITY>{}); break;
  }
}

// Everything warp 0 needs to issue the MMAs of one hidden layer, built once per kernel (setup) so that an issue
// block starts with three 16-byte shared-memory loads instead of a chain of dynamically indexed constant loads and
// descriptor arithmetic (400 cycles per block, measured).  Descriptors point at the first k-step.
struct alignas(16) IssueDesc {
  uint64_t f_ahi, f_alo, f_bhi, f_blo;  // forward: A = input activations (K-major), B = weights (MN-major)
  uint64_t a_dhi, a_dlo, a_whi, a_wlo;  // input gradient: A = delta (K-major), B = weights (K-major)
  uint64_t w_dhi, w_dlo, w_xhi, w_xlo;  // weight gradient: A = delta (MN-major), B = input activations (MN-major)
  uint32_t idesc_f, idesc_a, idesc_w, nk_f, nk_a, bstep_f, xstep_w, dw_col;
};
static_assert(sizeof(IssueDesc) == 128, "one IssueDesc = 128 bytes of the misc area");

struct Region {
  unsigned char* base;  // 128-byte aligned tensor-core region
  const IssueDesc* idesc;  // [nh]
  float* shadow;           // [8] results of the caller's shadow job (see chain_eval_tc)
  uint32_t tmem;        // TMEM base address (lane 0, column 0)
  float* yh;            // [2][64] partial network outputs of the two column halves
  float* glast;         // [72]    gradient of the output layer (weights, then bias at [64])
  float* gpart;         // [4][65] its per-lane-quadrant partials of the current tile
  uint64_t* bar;        // [4]
  uint32_t* tmem_slot;
  uint32_t par[3];      // phase parity of bar[0..2]
};

template <class SP = TcDyn>
BFX_DEV Region region_of(const ModelDev& M, const ChainSmem& s) {
  Region R;
  R.base = s.tc;
  float* f = reinterpret_cast<float*>(s.tc + SP::misc_off(M));
  R.yh = f;
  R.glast = f + 128;
  R.bar = reinterpret_cast<uint64_t*>(f + 200);
  R.tmem_slot = reinterpret_cast<uint32_t*>(f + 208);
  R.gpart = f + 256;
  R.idesc = reinterpret_cast<const IssueDesc*>(s.tc + SP::misc_off(M) + 2304);
  R.shadow = f + 520;
  R.tmem = 0;
  R.par[0] = R.par[1] = R.par[2] = 0;
  return R;
}

// kernel prologue: TMEM allocation, barriers, constant parts of the tiles.  All threads call it.
template <int NT>
BFX_DEV void setup(const ModelDev& M, Region& R) {
  const TcDesc& T = M.tc;
  if ((threadIdx.x >> 5) == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(R.tmem_slot)),
                 "r"((uint32_t)T.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x < 3) mbar_init(&R.bar[threadIdx.x], 1);
  // zero every tile, then the constant-one column of the activation tiles
  uint4* z = reinterpret_cast<uint4*>(R.base);
  for (int i = threadIdx.x; i < T.misc_off / 16; i += NT) z[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (int l = 0; l < T.nh; ++l) {
    const TcLayer& L = T.L[l];
    const int sbo = (L.ccols >> 3) * 128;
    for (int b = threadIdx.x; b < 64; b += NT) {
      const int off = (b >> 3) * sbo + (L.kpad >> 3) * 128 + (b & 7) * 16;
      *reinterpret_cast<unsigned short*>(R.base + L.a_off + off) = 0x3F80;  // bf16(1), hi tile only
    }
  }
  if ((int)threadIdx.x < T.nh) {
    const int l = threadIdx.x;
    const TcLayer& L = T.L[l];
    IssueDesc D;
    const uint32_t ahi = smem_u32(R.base + L.a_off), whi = smem_u32(R.base + L.w_off);
    const uint32_t dhi = smem_u32(R.base + T.d_off);
    const uint32_t a_sbo = (L.ccols >> 3) * 128, b_lbo = (L.hout >> 3) * 128, w_sbo = (L.hout >> 3) * 128;
    const uint32_t x_lbo = (L.ccols >> 3) * 128;
    // forward.  A: K-major (rows = batch): LBO = next 16-byte K chunk, SBO = next 8 rows; a k-step = 2 chunks.
    //           B: MN-major (rows = fan-in k, cols = fan-out n): SBO = next 8 n, LBO = next 8 k rows
    D.f_ahi = make_desc(ahi, 128, a_sbo);
    D.f_alo = make_desc(ahi + L.a_bytes, 128, a_sbo);
    D.f_bhi = make_desc(whi, b_lbo, 128);
    D.f_blo = make_desc(whi + L.w_bytes, b_lbo, 128);
    // dA[b][i] = sum_o D[b][o] W[o][i]:  A = delta tile K-major (K = fan-out), B = W tile K-major (rows n = fan-in)
    D.a_dhi = make_desc(dhi, 128, 1024);
    D.a_dlo = make_desc(dhi + T.d_bytes, 128, 1024);
    D.a_whi = make_desc(whi, 128, w_sbo);
    D.a_wlo = make_desc(whi + L.w_bytes, 128, w_sbo);
    // dW[o][c] += sum_b D[b][o] Ain[b][c]:  A = delta tile MN-major (M = fan-out), B = Ain tile MN-major, K = batch
    D.w_dhi = make_desc(dhi, 1024, 128);
    D.w_dlo = make_desc(dhi + T.d_bytes, 1024, 128);
    D.w_xhi = make_desc(ahi, x_lbo, 128);
    D.w_xlo = make_desc(ahi + L.a_bytes, x_lbo, 128);
    D.idesc_f = make_idesc(L.hout, 0, 1);
    D.idesc_a = make_idesc(L.kpad, 0, 0);
    D.idesc_w = make_idesc(L.ccols, 1, 1);
    D.nk_f = (uint32_t)(L.kpad >> 4);
    D.nk_a = (uint32_t)(L.hout >> 4);
    D.bstep_f = (2 * b_lbo) >> 4;
    D.xstep_w = (2 * x_lbo) >> 4;
    D.dw_col = (uint32_t)L.dw_col;
    *const_cast<IssueDesc*>(R.idesc + l) = D;
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  R.tmem = *R.tmem_slot;
}

template <int NT>
BFX_DEV void teardown(const ModelDev& M, Region& R) {
  fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(R.tmem), "r"((uint32_t)M.tc.tmem_cols)
                 : "memory");
}

// theta (FP32 master copy in s.th) -> hi/lo BF16 weight tiles.  Tile element (r = fan-in i, c = fan-out o).
template <int NT, class SP = TcDyn>
BFX_DEV void pack_weights(const ModelDev& M, const ChainSmem& s, const Region& R) {
#pragma unroll(SP::fixed ? 2 : 1)
  for (int l = 0; l < SP::nh(M); ++l) {
    const TcLayer& L = SP::tl(M, l);
    const LayerDev& D = SP::ld(M, l);
    const int noc = L.hout >> 3;           // 8-wide fan-out chunks
    const int nin8 = (D.nin + 7) >> 3;     // 8-row fan-in blocks (rows i >= nin stay zero)
    unsigned char* hi = R.base + L.w_off;
    unsigned char* lo = hi + L.w_bytes;
    // work item of a warp = (8-row fan-in block, group of 4 fan-out chunks): lanes walk i%8 fastest (8 lanes x 16 B =
    // one conflict-free 128-byte row of the core block).  No integer division per element.
    const int lane = threadIdx.x & 31, il = lane & 7;
    const int two = noc > 4 ? 1 : 0, items = nin8 << two;  // hout <= 64: one or two groups of 4 chunks per block
    for (int it = threadIdx.x >> 5; it < items; it += NT / 32) {
      const int ih = it >> two, oc = 4 * (it & two) + (lane >> 3);
      const int i = ih * 8 + il;
      if (i >= D.nin || oc >= noc) continue;
      const float* src = s.th + D.w_s + i * D.ldw + 8 * oc;
      const float4 a = ld4(src), b = ld4(src + 4);
      uint4 h, q;
      split2(a.x, a.y, h.x, q.x);
      split2(a.z, a.w, h.y, q.y);
      split2(b.x, b.y, h.z, q.z);
      split2(b.z, b.w, h.w, q.w);
      const int off = ih * (noc * 128) + oc * 128 + il * 16;
      *reinterpret_cast<uint4*>(hi + off) = h;
      *reinterpret_cast<uint4*>(lo + off) = q;
    }
  }
}

// one k-step = the three MMAs of the split product
BFX_DEV void mma3(uint32_t d, uint64_t ahi, uint64_t alo, uint64_t bhi, uint64_t blo, uint32_t idesc, uint32_t acc) {
  mma(d, ahi, bhi, idesc, acc);
  mma(d, ahi, blo, idesc, 1u);
  mma(d, alo, bhi, idesc, 1u);
}

// ---------------------------------------------------------------------------------------------
// value / gradient over a minibatch (loop over 64-row tiles).  NT must be 256.
// ---------------------------------------------------------------------------------------------
// `shadow`: a functor run once per evaluation by warp 1 while warp 0 issues the first MMAs (warps 1-7 only wait
// there): the caller's scalar bookkeeping that does not depend on the gradient (step size, sample slot) leaves the
// serial path this way; results travel through Region::shadow (shared memory, visible after the next barrier).
struct NoShadow {
  BFX_DEV void operator()() const {}
};
template <int NT, class Shadow = NoShadow, class SP = TcDyn, bool VAL = true>
BFX_DEV double chain_eval_tc(const ModelDev& M, const ChainSmem& s, Region& R, const float* __restrict__ x,
                             const uint4* __restrict__ xs, const float* __restrict__ y, const BatchSrc& bs, long long B,
                             float nb, bool want_grad, float& gn2, Shadow shadow = Shadow{}) {
  static_assert(NT == 256, "the tensor-core path uses 8 warps: 4 lane quadrants x 2 column halves");
  const int nh = SP::nh(M);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // warp index / TMEM base broadcast from lane 0: tells the compiler they are warp-uniform, so the MMA-issue code
  // (descriptor arithmetic included) runs on the uniform datapath instead of per-thread registers + R2UR moves
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, R.tmem, 0);
  const int q = warp & 3, h = warp >> 2;
  const int r0 = 16 * q + (lane >> 2), r1 = r0 + 8;  // the two batch rows of this thread's fragment
  const uint32_t zaddr = R.tmem + ((uint32_t)(32 * q) << 16);
  const LayerDev& LO = SP::ld(M, nh);  // the 1-wide output layer
  const float sl = s.th[SP::like_s(M)];
  const float inv_sigma = 1.f / expf(sl);
  double ds0 = 0.0, ds1 = 0.0;

  BFX_TICK(0);  // everything since the previous evaluation ended (update, record, bookkeeping)
  if (want_grad)
    for (int k = tid; k < 72; k += NT) R.glast[k] = 0.f;
  if (!xs) pack_weights<NT, SP>(M, s, R);  // (with xs: packed while the first tile's copies are in flight)
  // (made visible to the async proxy by the fence + barrier after the first tile is staged)

  // ---- dW accumulators (TMEM) -> s.g in the padded parameter layout.  ADD = std::true_type: add to what an earlier
  // dump of the same evaluation left there.  The tensor core's FP32 accumulation truncates, and the bias grows with
  // the number of MMAs chained into one accumulator (1.3e-3 relative over the 15 625 tiles of a 10^6-observation
  // batch, tests/test_gpu_fullsize.py): accumulators are therefore flushed every kAccTiles tiles and the partial
  // sums added on the CUDA cores (round to nearest).  (Shorter chains were tried against the direct full-size oracle
  // comparison, tests/test_gpu_fullsize.py: 8 instead of 32 tiles changes the error from 9.5e-6 to 8.6e-6 on n = 10^6
  // and nothing on a cancelling full batch, whose 1.4e-4 comes from ReLU units on the other side of their kink:
  // 1.4e-6 on the observations that keep a margin.  32 keeps the flush below 2 % of a full-batch evaluation.)
  constexpr int kAccTiles = 32;
  // LAST = std::true_type: the dump that completes the gradient also folds in the prior term and accumulates the
  // two squared norms (theta_net, finished gradient) per thread, so that eval_finish needs no pass of its own over
  // the parameters (the element of theta that pairs with a gradient slot sits S floats below it)
  float fa0 = 0.f, fa1 = 0.f;
  const float c2 = M.prior_c2;
  const long long th_off = s.th - s.g;
  auto dump = [&](auto ADD, auto LAST) {
    constexpr bool add = decltype(ADD)::value;
    constexpr bool last = decltype(LAST)::value;
    auto put = [&](float* p, float v) {
      if (add) v += *p;
      if (last) {
        const float t = p[th_off];
        v = fmaf(-c2, t, v);
        if (VAL) fa0 = fmaf(t, t, fa0);
        fa1 = fmaf(v, v, fa1);
      }
      *p = v;
    };
    fence_after();
#pragma unroll(SP::fixed ? 2 : 1)
    for (int l = 0; l < nh; ++l) {
      const TcLayer& L = SP::tl(M, l);
      const LayerDev& D = SP::ld(M, l);
      const int ngroups = L.ccols >> 3;
      // 64 columns that lie fully inside the weight matrix: one 32-column fragment per column half
      int g0 = 0;
      if (D.nout == 64) {
        for (; 8 * (g0 + 8) <= D.nin; g0 += 8) {
          float v[16];
          ld_16x256b_x4(zaddr + (uint32_t)(L.dw_col + 8 * (g0 + 4 * h)), v);
          float* gp = s.g + D.w_s + (8 * (g0 + 4 * h) + 2 * (lane & 3)) * D.ldw + r0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            put(gp, v[4 * j + 0]);
            put(gp + D.ldw, v[4 * j + 1]);
            put(gp + 8, v[4 * j + 2]);
            put(gp + D.ldw + 8, v[4 * j + 3]);
            gp += 8 * D.ldw;
          }
        }
      }
      for (int gq = g0 + h; gq < ngroups; gq += 2) {
        float v[4];
        ld_16x256b_x1(zaddr + (uint32_t)(L.dw_col + 8 * gq), v);
        const int c = 8 * gq + 2 * (lane & 3);
        if (8 * gq + 8 <= D.nin && D.nout == 64) {  // group fully inside the weight matrix: no bounds checks
          float* gp = s.g + D.w_s + c * D.ldw + r0;
          put(gp, v[0]);
          put(gp + D.ldw, v[1]);
          put(gp + 8, v[2]);
          put(gp + D.ldw + 8, v[3]);
        } else if (8 * gq == L.kpad) {  // the constant-1 column: the bias gradient sits in column kpad only
          if ((lane & 3) == 0) {
            if (r0 < D.nout) put(s.g + D.b_s + r0, v[0]);
            if (r1 < D.nout) put(s.g + D.b_s + r1, v[2]);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int o = (e & 2) ? r1 : r0, cc = c + (e & 1);
            if (o < D.nout) {
              if (cc < D.nin) put(s.g + D.w_s + cc * D.ldw + o, v[e]);
              else if (cc == L.kpad) put(s.g + D.b_s + o, v[e]);
            }
          }
        }
      }
    }
  };
  int acc_tiles = 0;
  bool dumped = false;
  const int d = SP::d_in(M);
  for (long long t0 = 0; t0 < B; t0 += 64) {
    const int nvalid = (int)((B - t0) < 64 ? (B - t0) : 64);
    const bool first_tile = t0 == 0;
    // ---- stage the tile: targets and X as hi/lo BF16 (rows >= nvalid are zero) ------------------------
    if (xs) {
      // x was split once per data set into rows of [hi chunks | lo chunks] (16 bytes = 8 bf16 each): staging is one
      // asynchronous 16-byte copy per chunk straight into the core-block tile.  Four adjacent threads own one
      // batch row (each derives the row's observation index itself: no index exchange, no barrier).
      const TcLayer& L0 = SP::tl(M, 0);
      const int nch = L0.kpad >> 3, cpr = 2 * nch;
      const int sbo = (L0.ccols >> 3) * 128;
      const int b = tid >> 2;
      const long long o = (b < nvalid) ? batch_obs(bs, t0 + b) : -1;
      unsigned char* row = R.base + L0.a_off + (b >> 3) * sbo + (b & 7) * 16;
      const uint4* src = xs + (o >= 0 ? o : 0) * cpr;
      for (int piece = tid & 3; piece < cpr; piece += 4) {
        const int half = piece >= nch ? 1 : 0, ch = piece - half * nch;
        cp_async16(row + half * L0.a_bytes + ch * 128, src + piece, o >= 0 ? 16 : 0);
      }
      if ((tid & 3) == 0) cp_async4(s.ys + b, y + (o >= 0 ? o : 0), o >= 0 ? 4 : 0);
      BFX_TICK(1);  // stage: indices + copy issue
      if (first_tile) pack_weights<NT, SP>(M, s, R);  // theta -> bf16 weight tiles, in the shadow of the copies
      cp_async_wait_all();
    } else {
      if (tid < 64) {
        long long o = (tid < nvalid) ? batch_obs(bs, t0 + tid) : -1;
        s.idx[tid] = (int)o;
        s.ys[tid] = (o >= 0) ? __ldg(y + o) : 0.f;
      }
      __syncthreads();
      const TcLayer& L0 = SP::tl(M, 0);
      const int nch = L0.kpad >> 3;
      const int sbo = (L0.ccols >> 3) * 128;
      unsigned char* hi = R.base + L0.a_off;
      unsigned char* lo = hi + L0.a_bytes;
      for (int e = tid; e < 64 * nch; e += NT) {
        const int bl = e & 7, rest = e >> 3, ch = rest % nch, bh = rest / nch;
        const int b = bh * 8 + bl;
        const int o = s.idx[b];
        float v[8];
#pragma unroll
        for (int f = 0; f < 8; ++f) {
          const int ff = 8 * ch + f;
          v[f] = (o >= 0 && ff < d) ? __ldg(x + (long long)o * d + ff) : 0.f;
        }
        uint4 hh, ll;
        split2(v[0], v[1], hh.x, ll.x);
        split2(v[2], v[3], hh.y, ll.y);
        split2(v[4], v[5], hh.z, ll.z);
        split2(v[6], v[7], hh.w, ll.w);
        const int off = bh * sbo + ch * 128 + bl * 16;
        *reinterpret_cast<uint4*>(hi + off) = hh;
        *reinterpret_cast<uint4*>(lo + off) = ll;
      }
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    BFX_TICK(2);  // stage indices, targets, X

    // ---- forward ---------------------------------------------------------------------------------
    float a[16];       // activations of the last hidden layer (this thread's fragment)
    float dl0 = 0.f, dl1 = 0.f;
#pragma unroll(SP::fixed ? 2 : 1)
    for (int l = 0; l < nh; ++l) {
      const TcLayer& L = SP::tl(M, l);
      const LayerDev& D = SP::ld(M, l);
      if (warp_u == 0) {
        fence_after();
        const bool lead = elect_one();
        const IssueDesc& D = R.idesc[SP::fixed ? l : __shfl_sync(0xffffffffu, l, 0)];  // (uniform layer index: see warp_u)
        uint64_t dahi = D.f_ahi, dalo = D.f_alo, dbhi = D.f_bhi, dblo = D.f_blo;
        const uint32_t idesc = D.idesc_f;
        const uint64_t astep = 256 >> 4, bstep = D.bstep_f;  // a k-step only moves the start-address field
        const int nk = (int)D.nk_f;
        BFX_TICK(18);  // (forward issue block: fence + elect + descriptor fetch)
        if (lead) {  // (the whole loop under the predicate: one branch, MMAs back to back)
          for (int k = 0; k < nk; ++k) {
            mma3(tmem_u, dahi, dalo, dbhi, dblo, idesc, k ? 1u : 0u);
            dahi += astep;
            dalo += astep;
            dbhi += bstep;
            dblo += bstep;
          }
          commit(&R.bar[0]);
        }
        __syncwarp();
      }
      BFX_TICK(16);  // (forward MMA issue, warp 0)
      if (warp_u == 1 && first_tile && l == 0) shadow();
      mbar_wait(&R.bar[0], R.par[0]);
      R.par[0] ^= 1;
      fence_after();
      BFX_TICK(3);  // forward MMA issue + wait
      // epilogue of layer l on the fragment: rows r0/r1, columns 32h + 8j + 2(lane%4) + {0,1}
      const bool active = 32 * h < L.hout;
      if (active) ld_16x256b_x4(zaddr + (uint32_t)(32 * h), a);
      const int cbase = 32 * h + 2 * (lane & 3);
      with_act(D.act, [&](auto A) {
        using AT = decltype(A);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = cbase + 8 * j;
          if (active && c < L.hout) {
            const float2 bb = *reinterpret_cast<const float2*>(s.th + D.b_s + c);
            a[4 * j + 0] = AT::f(a[4 * j + 0] + bb.x);
            a[4 * j + 1] = AT::f(a[4 * j + 1] + bb.y);
            a[4 * j + 2] = AT::f(a[4 * j + 2] + bb.x);
            a[4 * j + 3] = AT::f(a[4 * j + 3] + bb.y);
          } else {
            a[4 * j + 0] = a[4 * j + 1] = a[4 * j + 2] = a[4 * j + 3] = 0.f;
          }
        }
      });
      if (l + 1 < nh) {
        // next layer's input tile
        const TcLayer& Ln = SP::tl(M, l + 1);
        const int sbo = (Ln.ccols >> 3) * 128;
        unsigned char* hi = R.base + Ln.a_off;
        unsigned char* lo = hi + Ln.a_bytes;
        if (active) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = cbase + 8 * j;
            if (c < L.hout) {
              uint32_t h0, l0, h1, l1;
              split2(a[4 * j + 0], a[4 * j + 1], h0, l0);
              split2(a[4 * j + 2], a[4 * j + 3], h1, l1);
              const int off0 = (r0 >> 3) * sbo + (c >> 3) * 128 + (r0 & 7) * 16 + (c & 7) * 2;
              const int off1 = off0 + sbo;  // r1 = r0 + 8
              *reinterpret_cast<uint32_t*>(hi + off0) = h0;
              *reinterpret_cast<uint32_t*>(lo + off0) = l0;
              *reinterpret_cast<uint32_t*>(hi + off1) = h1;
              *reinterpret_cast<uint32_t*>(lo + off1) = l1;
            }
          }
        }
        fence_async_smem();
        fence_before();
        __syncthreads();
        BFX_TICK(4);  // hidden-layer epilogue
      }
    }
    // ---- output layer (fan-out 1) + likelihood on the fragment of the last hidden layer -----------
    const int hl = SP::tl(M, nh - 1).hout;
    const int cbase = 32 * h + 2 * (lane & 3);
    float wl[8];
    {
      float p0 = 0.f, p1 = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = cbase + 8 * j;
        wl[2 * j] = (c < hl) ? s.th[LO.w_s + c * LO.ldw] : 0.f;
        wl[2 * j + 1] = (c + 1 < hl) ? s.th[LO.w_s + (c + 1) * LO.ldw] : 0.f;
        p0 = fmaf(a[4 * j + 0], wl[2 * j], fmaf(a[4 * j + 1], wl[2 * j + 1], p0));
        p1 = fmaf(a[4 * j + 2], wl[2 * j], fmaf(a[4 * j + 3], wl[2 * j + 1], p1));
      }
      p0 += __shfl_xor_sync(0xffffffffu, p0, 1);
      p0 += __shfl_xor_sync(0xffffffffu, p0, 2);
      p1 += __shfl_xor_sync(0xffffffffu, p1, 1);
      p1 += __shfl_xor_sync(0xffffffffu, p1, 2);
      if ((lane & 3) == 0) {
        R.yh[64 * h + r0] = p0;
        R.yh[64 * h + r1] = p1;
      }
    }
    __syncthreads();
    {
      const float bl = s.th[LO.b_s];
      const bool owner = (lane & 3) == 0 && h == 0;
      float sm0 = 0.f, sm1 = 0.f;
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int r = rr ? r1 : r0;
        float dl = 0.f;
        if (r < nvalid) {
          const float yh = act_apply(LO.act, R.yh[r] + R.yh[64 + r] + bl);
          const float z = (s.ys[r] - yh) * inv_sigma;
          if (M.like_kind == LK_NORMAL) {
            sm0 += z * z;
            dl = nb * z * inv_sigma;
          } else {
            const float zz = z * z;
            const float w = (M.nu + 1.f) * z / (M.nu + zz);
            sm0 += log1pf(zz / M.nu);
            sm1 += w * z;
            dl = nb * w * inv_sigma;
          }
          dl *= act_deriv_from_out(LO.act, yh);
        }
        if (rr) dl1 = dl; else dl0 = dl;
      }
      if (owner) {
        ds0 += (double)sm0;
        ds1 += (double)sm1;
      }
    }
    if (want_grad) {
      // output-layer gradient: dW[c] = sum_b dl[b] a[b][c], db = sum_b dl[b]
      {
        float pc[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          pc[2 * j] = dl0 * a[4 * j + 0] + dl1 * a[4 * j + 2];
          pc[2 * j + 1] = dl0 * a[4 * j + 1] + dl1 * a[4 * j + 3];
        }
        float pb = dl0 + dl1;
#pragma unroll
        for (int o = 4; o <= 16; o <<= 1) {
#pragma unroll
          for (int i = 0; i < 8; ++i) pc[i] += __shfl_xor_sync(0xffffffffu, pc[i], o);
          pb += __shfl_xor_sync(0xffffffffu, pb, o);
        }
        // per-quadrant partials, then a fixed-order sum: the result must not depend on warp timing (resume
        // determinism: a continued run reproduces the uninterrupted one bit for bit)
        if (lane < 4) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = cbase + 8 * j;
            R.gpart[q * 65 + c] = pc[2 * j];
            R.gpart[q * 65 + c + 1] = pc[2 * j + 1];
          }
          if (lane == 0 && h == 0) R.gpart[q * 65 + 64] = pb;
        }
      }
      // delta of the last hidden layer -> D tile (64 x 64, zero beyond hl)
      {
        const int act_l = SP::ld(M, nh - 1).act;
        unsigned char* hi = R.base + SP::d_off(M);
        unsigned char* lo = hi + SP::d_bytes(M);
        const int offb = (r0 >> 3) * 1024 + (r0 & 7) * 16 + (lane & 3) * 4 + h * 512;
        with_act(act_l, [&](auto A) {
          using AT = decltype(A);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float d00 = dl0 * wl[2 * j] * AT::d(a[4 * j + 0]);
            const float d01 = dl0 * wl[2 * j + 1] * AT::d(a[4 * j + 1]);
            const float d10 = dl1 * wl[2 * j] * AT::d(a[4 * j + 2]);
            const float d11 = dl1 * wl[2 * j + 1] * AT::d(a[4 * j + 3]);
            uint32_t h0, l0, h1, l1;
            split2(d00, d01, h0, l0);
            split2(d10, d11, h1, l1);
            const int off0 = offb + j * 128;  // column 32h + 8j + 2(lane%4): chunk 4h + j, 2 bytes per column
            *reinterpret_cast<uint32_t*>(hi + off0) = h0;
            *reinterpret_cast<uint32_t*>(lo + off0) = l0;
            *reinterpret_cast<uint32_t*>(hi + off0 + 1024) = h1;
            *reinterpret_cast<uint32_t*>(lo + off0 + 1024) = l1;
          }
        });
      }
      fence_async_smem();
      fence_before();
      __syncthreads();
      if (tid <= 64 && (tid < hl || tid == 64))
        R.glast[tid] += (R.gpart[tid] + R.gpart[65 + tid]) + (R.gpart[130 + tid] + R.gpart[195 + tid]);
      BFX_TICK(5);  // output layer + likelihood + delta tile
      // ---- backward ------------------------------------------------------------------------------
#pragma unroll(SP::fixed ? 2 : 1)
      for (int lv = nh - 1; lv >= 0; --lv) {
        const int l = lv;
        const TcLayer& L = SP::tl(M, l);
        if (warp_u == 0) {
          fence_after();
          const bool lead = elect_one();
          const int l = SP::fixed ? lv : __shfl_sync(0xffffffffu, lv, 0);  // (uniform layer index: see warp_u)
          const IssueDesc& D = R.idesc[l];
          if (lead) {
            if (l > 0) {
              uint64_t ddhi = D.a_dhi, ddlo = D.a_dlo, dwhi = D.a_whi, dwlo = D.a_wlo;
              const uint32_t idesc = D.idesc_a;
              const int nk = (int)D.nk_a;
              for (int k = 0; k < nk; ++k) {
                mma3(tmem_u, ddhi, ddlo, dwhi, dwlo, idesc, k ? 1u : 0u);
                ddhi += 16;  // 256 bytes
                ddlo += 16;
                dwhi += 16;
                dwlo += 16;
              }
              commit(&R.bar[1]);
            }
            uint64_t ddhi = D.w_dhi, ddlo = D.w_dlo, dxhi = D.w_xhi, dxlo = D.w_xlo;
            const uint32_t idesc = D.idesc_w;
            const uint64_t xstep = D.xstep_w;
            const uint32_t dwacc = tmem_u + D.dw_col;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              mma3(dwacc, ddhi, ddlo, dxhi, dxlo, idesc, (k || acc_tiles > 0) ? 1u : 0u);
              ddhi += 128;  // 2048 bytes
              ddlo += 128;
              dxhi += xstep;
              dxlo += xstep;
            }
            commit(&R.bar[2]);
          }
          __syncwarp();
        }
        BFX_TICK(17);  // (backward MMA issue, warp 0)
        if (l > 0) {
          mbar_wait(&R.bar[1], R.par[1]);
          R.par[1] ^= 1;
          fence_after();
          BFX_TICK(6);  // backward MMA issue + wait (dA)
          // delta of layer l-1 = dA * act'(A_{l-1}); A_{l-1} is the input tile of layer l
          const int hp = L.kpad;  // = fan-out of layer l-1
          const int act_p = SP::ld(M, l - 1).act;
          const bool active = 32 * h < hp;
          float da[16];
          if (active) ld_16x256b_x4(zaddr + (uint32_t)(32 * h), da);
          const int sbo = (L.ccols >> 3) * 128;
          const unsigned char* ahi = R.base + L.a_off;
          const unsigned char* alo = ahi + L.a_bytes;
          uint32_t oh[8], ol[8];
          const int offa = (r0 >> 3) * sbo + (r0 & 7) * 16 + (lane & 3) * 4 + h * 512;
          with_act(act_p, [&](auto A) {
            using AT = decltype(A);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int c = cbase + 8 * j;
              oh[2 * j] = ol[2 * j] = oh[2 * j + 1] = ol[2 * j + 1] = 0u;
              if (active && c < hp) {
                const int off0 = offa + j * 128;
                const uint32_t h0 = *reinterpret_cast<const uint32_t*>(ahi + off0);
                const uint32_t h1 = *reinterpret_cast<const uint32_t*>(ahi + off0 + sbo);
                float a00 = bf16lo_to_f(h0), a01 = bf16hi_to_f(h0), a10 = bf16lo_to_f(h1), a11 = bf16hi_to_f(h1);
                if (!std::is_same<AT, ActT<A_RELU>>::value && !std::is_same<AT, ActT<A_IDENTITY>>::value) {
                  // tanh / sigmoid derivatives need the full-precision output: add the lo half
                  const uint32_t l0 = *reinterpret_cast<const uint32_t*>(alo + off0);
                  const uint32_t l1 = *reinterpret_cast<const uint32_t*>(alo + off0 + sbo);
                  a00 += bf16lo_to_f(l0); a01 += bf16hi_to_f(l0); a10 += bf16lo_to_f(l1); a11 += bf16hi_to_f(l1);
                }
                split2(da[4 * j + 0] * AT::d(a00), da[4 * j + 1] * AT::d(a01), oh[2 * j], ol[2 * j]);
                split2(da[4 * j + 2] * AT::d(a10), da[4 * j + 3] * AT::d(a11), oh[2 * j + 1], ol[2 * j + 1]);
              }
            }
          });
          // the dW MMAs of layer l still read the D tile: wait for them before overwriting it
          mbar_wait(&R.bar[2], R.par[2]);
          R.par[2] ^= 1;
          unsigned char* hi = R.base + SP::d_off(M);
          unsigned char* lo = hi + SP::d_bytes(M);
          const int offd = (r0 >> 3) * 1024 + (r0 & 7) * 16 + (lane & 3) * 4 + h * 512;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int off0 = offd + j * 128;
            *reinterpret_cast<uint32_t*>(hi + off0) = oh[2 * j];
            *reinterpret_cast<uint32_t*>(lo + off0) = ol[2 * j];
            *reinterpret_cast<uint32_t*>(hi + off0 + 1024) = oh[2 * j + 1];
            *reinterpret_cast<uint32_t*>(lo + off0 + 1024) = ol[2 * j + 1];
          }
          fence_async_smem();
          fence_before();
          __syncthreads();
          BFX_TICK(7);  // backward epilogue (delta of the layer below)
        } else {
          mbar_wait(&R.bar[2], R.par[2]);
          R.par[2] ^= 1;
          fence_after();
          BFX_TICK(8);  // last dW MMA issue + wait
        }
      }
    }
    fence_before();
    __syncthreads();  // tile buffers, yh and the Z accumulator are free again
    if (want_grad && ++acc_tiles == kAccTiles && t0 + 64 < B) {
      if (dumped) dump(std::true_type{}, std::false_type{});
      else dump(std::false_type{}, std::false_type{});
      dumped = true;
      acc_tiles = 0;
      fence_before();
      __syncthreads();
    }
  }

  if (want_grad) {
    if (dumped) dump(std::true_type{}, std::true_type{});
    else dump(std::false_type{}, std::true_type{});
    const int hl = SP::tl(M, nh - 1).hout;
    for (int k = tid; k <= hl; k += NT) {  // output layer: weights, then its bias
      const int kk = k < hl ? LO.w_s + k * LO.ldw : LO.b_s;
      const float t = s.th[kk];
      const float v = fmaf(-c2, t, R.glast[k < hl ? k : 64]);
      fa0 = fmaf(t, t, fa0);
      fa1 = fmaf(v, v, fa1);
      s.g[kk] = v;
    }
    fence_before();
  }
  __syncthreads();
  BFX_TICK(9);  // gradient dump
  const double v__ = want_grad ? eval_finish<NT, SP, false, VAL>(M, s, B, nb, ds0, ds1, true, gn2, fa0, fa1)
                               : eval_finish<NT, SP, true, VAL>(M, s, B, nb, ds0, ds1, false, gn2);
  BFX_TICK(10);  // prior + norms + value
  return v__;
}

}  // namespace tc
}  // namespace bfx
