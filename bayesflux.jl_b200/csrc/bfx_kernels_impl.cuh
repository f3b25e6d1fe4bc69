// Kernel templates of the small-P regime and their templated launchers (included by the bfx_k_*.cu
// translation units, each of which instantiates one launch shape).
//   k_eval  : loglikeprior / ∇loglikeprior for C chains on one (explicit or full) batch
//   k_mcmc  : nsteps update! calls per chain, theta resident in shared memory
//   k_debug_*: expose the engine's batch schedule / noise streams to the tests
#pragma once
#include <cstdio>
#include <cstdlib>

#include "bfx_kernels.h"
#include "bfx_samplers.cuh"

namespace bfx {

// permutation key of (seed, chain-or-shared, epoch)
BFX_DEV void perm_keys(unsigned long long seed, unsigned chainkey, unsigned long long epoch, unsigned& k0,
                       unsigned& k1) {
  uint4 r = philox4x32_10(make_uint4((unsigned)epoch, (unsigned)(epoch >> 32), chainkey, 0x5eedba7cu),
                          make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
  k0 = r.x;
  k1 = r.y;
}

// minibatch of global step `gstep` (0-based): epoch = gstep / nb, batch k = gstep % nb covers
// positions [k*B, min(n, (k+1)*B)) of that epoch's permutation (abstract.jl:102-105,117-118)
// `fresh` = false: bs still holds the keys of the previous step of the same chain, which stay valid inside an epoch
BFX_DEV void make_batch(const RunDev& R, int c, unsigned long long gstep, BatchSrc& bs, long long& Bk,
                        bool fresh = true) {
  bs.n = R.n;
  bs.full = 0;
  bs.shuffle = R.shuffle;
  if (R.idx) {
    bs.idx = R.idx + (long long)c * R.idx_chain_stride;
    bs.first = 0;
    bs.k0 = bs.k1 = 0;
    Bk = R.B;
    return;
  }
  bs.idx = nullptr;
  unsigned long long epoch, k;
  fdivmod(R.fd_nb, gstep, epoch, k);
  bs.first = (long long)k * R.B;
  long long rem = R.n - bs.first;
  Bk = rem < R.B ? rem : R.B;
  if (fresh || k == 0ull) {
    unsigned chainkey = R.batch_shared ? 0xFFFFFFFFu : (R.chain0 + (unsigned)c);
    perm_keys(R.seed, chainkey, epoch, bs.k0, bs.k1);
  }
}

template <int NT>
BFX_DEV void load_mask(const ModelDev& M, const ChainSmem& s) {
  for (int i = threadIdx.x; i < M.S / 4; i += NT) {
    s.mask[i] = __ldg(M.pad_mask + i);
    const int J0 = __ldg(M.pad_flat + 4 * i);
    s.qflat[i] = (unsigned short)(J0 < 0 ? 0 : J0);
  }
}

template <int NT, int BT, bool TC>
__global__ void __launch_bounds__(NT) k_eval(const __grid_constant__ EvalDev E) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const ModelDev& M = E.M;
  ChainSmem s = carve_smem<BT>(M, smem_raw);
  const int c = blockIdx.x;
  s.scr = E.scratch ? E.scratch + (long long)c * E.scratch_stride : nullptr;
  tc::Region tr{};
  if constexpr (TC) {
    tr = tc::region_of(M, s);
    tc::setup<NT>(M, tr);
  }
  load_mask<NT>(M, s);
  for (int k = threadIdx.x; k < M.S; k += NT) s.th[k] = 0.f;
  if constexpr (TC)  // (the tensor-core evaluation writes only the valid gradient slots: padding stays zero from here)
    for (int k = threadIdx.x; k < M.S; k += NT) s.g[k] = 0.f;
  __syncthreads();
  const float* th_in = E.theta + (long long)c * M.P;
  for_each_flat<NT>(M, [&](int J, int k) { s.th[k] = th_in[J]; });
  __syncthreads();
  BatchSrc bs;
  bs.n = E.n;
  bs.first = 0;
  bs.k0 = bs.k1 = 0;
  bs.shuffle = 0;
  bs.full = E.idx ? 0 : 1;
  bs.idx = E.idx ? E.idx + (long long)c * E.idx_chain_stride : nullptr;
  float gn2;
  double v;
  if constexpr (TC) v = tc::chain_eval_tc<NT>(M, s, tr, E.x, E.xs, E.y, bs, E.B, E.num_batches, E.want_grad != 0, gn2);
  else v = chain_eval<NT, BT>(M, s, E.x, E.y, bs, E.B, E.num_batches, E.want_grad != 0, gn2,
                              E.yhat_out ? E.yhat_out + (long long)c * E.B : nullptr);
  if (threadIdx.x == 0) E.v_out[c] = v;
  if (E.want_grad) {
    float* g_out = E.g_out + (long long)c * M.P;
    for_each_flat<NT>(M, [&](int J, int k) { g_out[J] = s.g[k]; });
  }
  if constexpr (TC) tc::teardown<NT>(M, tr);
}

// L2-coherent copy of a chain's scalar state (see gld4)
BFX_DEV ChainScalars load_scalars(const ChainScalars* p) {
  static_assert(sizeof(ChainScalars) % 8 == 0, "ChainScalars is copied in 8-byte words");
  ChainScalars out;
  const long long* src = reinterpret_cast<const long long*>(p);
  long long* dst = reinterpret_cast<long long*>(&out);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(ChainScalars) / 8); ++i) dst[i] = __ldcg(src + i);
  return out;
}

// nsteps update! calls for every chain.  The launch is a pool of persistent CTAs (one or two per SM) that pull
// (chain, chunk) work items from a counter: item i = chunk i / C of chain i % C, chunk_len updates each.  Chunks
// of one chain run in order (the per-chain counter work[1 + c] is the hand-over), so a launch is balanced over
// the SMs for any number of chains instead of running ceil(C / resident CTAs) full waves.
template <int NT, int BT, int KIND, bool TC, class SP = TcDyn>
// (at least 512 resident threads per SM for every launch shape = a 128-register cap, and 24 one-warp CTAs = an
// 80-register cap for the 32-thread shape: without it the small shapes took 200+ registers and ran 10 CTAs per SM;
// cfg1 127 M -> 200 M grad-evals/s)
// (the recurrent shape <256,16> runs one CTA per SM at configs[3]'s size, 207 KB: no register cap there, the 64
// accumulators of the forward-only LSTM tile spilled under the 128-register one)
__global__ void __launch_bounds__(NT, (NT == 32 ? 24 : (BT == 16 ? 1 : 512 / NT))) k_mcmc(const __grid_constant__ RunDev R) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ int s_item;
  const ModelDev& M = R.M;
  ChainSmem s = carve_smem<BT, SP>(M, smem_raw);
  tc::Region tr{};
  if constexpr (TC) {
    tr = tc::region_of<SP>(M, s);
    tc::setup<NT>(M, tr);
  }
  s.scr = R.scratch ? R.scratch + (long long)blockIdx.x * R.scratch_stride : nullptr;
  load_mask<NT>(M, s);
  if constexpr (TC)  // (the tensor-core evaluation writes only the valid gradient slots: padding stays zero from here)
    for (int k = threadIdx.x; k < M.S; k += NT) s.g[k] = 0.f;
  const int nchunks = (R.nsteps + R.chunk_len - 1) / R.chunk_len;
  const int ccount = R.c_count > 0 ? R.c_count : R.C;
  const int total = ccount * nchunks;
  Ctx<NT, BT, TC, SP> X(R, s);
  X.tcr = &tr;
  X.key.seed = R.seed;
  while (true) {
    if (threadIdx.x == 0) s_item = atomicAdd(R.work, 1);
    __syncthreads();
    const int item = s_item;
    __syncthreads();
    if (item >= total) break;
    const int chunk = item / ccount, c = R.c_begin + (item - chunk * ccount);
    if (threadIdx.x == 0) {
      volatile int* done = R.work + 1 + c;
      while (*done < chunk) __nanosleep(256);
      __threadfence();
    }
    __syncthreads();
    X.c = c;
    X.key.chain = R.chain0 + (unsigned)c;
    X.sc = load_scalars(R.sc + c);
    const long long row = (long long)c * R.Ps;
    X.th_g = R.theta + row;
    X.mom = R.mom ? R.mom + row : nullptr;
    X.minv = R.minv ? R.minv + row : nullptr;
    X.thold = R.theta_old ? R.theta_old + row : nullptr;
    X.rmsv = R.rms_v ? R.rms_v + row : nullptr;
    X.ring = R.ring ? R.ring + (long long)c * R.S.ma_windowlength * R.Ps : nullptr;
    X.window = R.mode_window ? R.mode_window + (long long)c * R.S.mode_windowlength : nullptr;
    for_each_quad<NT>(M, [&](int k) { st4(s.th + k, gld4(X.th_g + k)); });
    __syncthreads();
    BFX_TICK(13);  // work-item fetch + chain state load
    const int it0 = chunk * R.chunk_len;
    const int it1 = it0 + R.chunk_len < R.nsteps ? it0 + R.chunk_len : R.nsteps;
    for (int it = it0; it < it1; ++it) {
      if (X.sc.status != 0) break;  // a NaN guard fired: this chain is frozen
      X.gstep = (unsigned long long)(R.gstep0 + it);
      make_batch(R, c, X.gstep, X.bs, X.Bk, it == it0);
      if (KIND == S_SGLD) step_sgld(X);
      else if (KIND == S_SGNHT) step_sgnht(X);
      else if (KIND == S_SGNHTS) step_sgnhts(X);
      else if (KIND == S_GGMC) step_ggmc(X);
      else if (KIND == S_HMC) step_hmc(X);
      else step_mode(X);
      __syncthreads();
      BFX_TICK(11);  // sampler update pass (+ fused recording)
      if (R.progress && c == 0 && threadIdx.x == 0) *(volatile unsigned long long*)R.progress = X.gstep + 1ull;
    }
    __syncthreads();
    for_each_quad<NT>(M, [&](int k) { st4(X.th_g + k, ld4(s.th + k)); });
    if (threadIdx.x == 0) {
      if (KIND == S_SGLD) {
        // SGLD only moves its two counters: writing the whole struct back would keep all 24 words of it live in
        // registers over the chunk (the fused kernels sit at their register cap)
        R.sc[c].t = X.sc.t;
        R.sc[c].nsampled = X.sc.nsampled;
      } else {
        R.sc[c] = X.sc;
      }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) atomicExch(R.work + 1 + c, chunk + 1);
    BFX_TICK(12);  // chain state store + hand-over
  }
  if constexpr (TC) tc::teardown<NT>(M, tr);
}

// ---- launchers --------------------------------------------------------------------
inline RunDev with_fastdiv(const RunDev& R) {
  RunDev Rf = R;
  Rf.fd_nb = make_fastdiv(R.nb);
  Rf.fd_keep = make_fastdiv(R.keep_every);
  Rf.fd_slots = make_fastdiv(R.ring_slots);
  Rf.fd_window = make_fastdiv(R.S.ma_windowlength);
  for (int r = 0; r < 10; ++r) {
    Rf.M.philox_rk[2 * r] = (unsigned)R.seed + (unsigned)r * 0x9E3779B9u;
    Rf.M.philox_rk[2 * r + 1] = (unsigned)(R.seed >> 32) + (unsigned)r * 0xBB67AE85u;
  }
  Rf.M.philox_rk_set = 1;
  return Rf;
}

template <int NT, int BT, bool TC = false>
inline cudaError_t launch_eval_t(const EvalDev& E, cudaStream_t st) {
  size_t smem = chain_smem_bytes<BT>(E.M) + (size_t)E.M.tc.smem_pad;
  auto kern = k_eval<NT, BT, TC>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<E.C, NT, smem, st>>>(E);
  return cudaGetLastError();
}

// SP = a TcFixed shape: only the kinds in KINDS (bit mask) are compiled for it, the launcher of the generic instance
// handles the rest
template <int NT, int BT, bool TC = false, class SP = TcDyn, int KINDS = 63>
inline cudaError_t launch_mcmc_t(const RunDev& R, cudaStream_t st) {
  size_t smem = chain_smem_bytes<BT>(R.M) + (size_t)R.M.tc.smem_pad;
  cudaError_t e;
  const RunDev Rf = with_fastdiv(R);
#define BFX_LAUNCH(KIND)                                                                        \
  {                                                                                             \
    auto kern = k_mcmc<NT, BT, KIND, TC, SP>;                                                       \
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
    if (e != cudaSuccess) return e;                                                             \
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,              \
                             (int)cudaSharedmemCarveoutMaxShared);                              \
    if (e != cudaSuccess) return e;                                                             \
    int per_sm = 0, dev = 0, sms = 0;                                                           \
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem);                 \
    if (e != cudaSuccess) return e;                                                             \
    cudaGetDevice(&dev);                                                                        \
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);                          \
    {                                                                                           \
      /* the occupancy query can under-report before the carve-out is applied: bound it by hand. */ \
      /* Over-subscribing the pool is harmless (late CTAs find the work counter exhausted).     */ \
      cudaFuncAttributes fa;                                                                    \
      if (cudaFuncGetAttributes(&fa, kern) == cudaSuccess) {                                    \
        long long by_smem = 232448 / (long long)(smem + fa.sharedSizeBytes + 1024);  /* 227 KB usable per SM */             \
        long long by_regs = 65536 / ((long long)(fa.numRegs > 0 ? fa.numRegs : 1) * NT);        \
        long long by_thr = 2048 / NT;                                                           \
        long long mn = by_smem < by_regs ? by_smem : by_regs;                                   \
        mn = mn < by_thr ? mn : by_thr;                                                         \
        mn = mn < 32 ? mn : 32;                                                                 \
        if (mn > per_sm) per_sm = (int)mn;                                                      \
      }                                                                                         \
    }                                                                                           \
    const long long items = (long long)(R.c_count > 0 ? R.c_count : R.C) * ((R.nsteps + R.chunk_len - 1) / R.chunk_len); \
    long long grid = (long long)(per_sm > 0 ? per_sm : 1) * sms;                                \
    if (grid > items) grid = items;                                                             \
    if (grid > R.grid_cap) grid = R.grid_cap;                                                   \
    if (std::getenv("BFX_DEBUG_LAUNCH"))                                                        \
      std::fprintf(stderr, "[bfx] k_mcmc<%d,%d,kind %d,tc %d>: dyn smem %zu (S %d, tc %d + pad %d), CTAs/SM %d, grid %lld\n", \
                   NT, BT, (int)KIND, (int)TC, smem, R.M.S, R.M.tc.total_bytes, R.M.tc.smem_pad, per_sm, grid); \
    kern<<<(unsigned)grid, NT, smem, st>>>(Rf);                                                 \
  }
  e = cudaErrorInvalidValue;
  switch (R.S.kind) {
    case S_SGLD: if constexpr ((KINDS >> S_SGLD) & 1) BFX_LAUNCH(S_SGLD); break;
    case S_SGNHT: if constexpr ((KINDS >> S_SGNHT) & 1) BFX_LAUNCH(S_SGNHT); break;
    case S_SGNHTS: if constexpr ((KINDS >> S_SGNHTS) & 1) BFX_LAUNCH(S_SGNHTS); break;
    case S_GGMC: if constexpr ((KINDS >> S_GGMC) & 1) BFX_LAUNCH(S_GGMC); break;
    case S_HMC: if constexpr ((KINDS >> S_HMC) & 1) BFX_LAUNCH(S_HMC); break;
    default: if constexpr ((KINDS >> S_MODE) & 1) BFX_LAUNCH(S_MODE); break;
  }
  if (e != cudaSuccess) return e;
#undef BFX_LAUNCH
  return cudaGetLastError();
}


}  // namespace bfx
