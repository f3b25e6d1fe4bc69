// Chain-level device functions: log posterior value / gradient over a whole minibatch
// (loop over batch tiles), and the per-sampler update steps.  One CTA = one chain.
#pragma once
#include "bfx_device.cuh"

namespace bfx {

// ---------------------------------------------------------------------------------
// value (and gradient into s.g) of  num_batches*loglike + logprior  on one minibatch.
// src/model/BNN.jl:50-69.  Returns the value on every thread; gn2 = ||g||^2 over all P.
// ---------------------------------------------------------------------------------
// PASS = false: the caller has already folded the prior into s.g and passes its per-thread partial sums of
// theta_net^2 (a0) and of the squared finished gradient (a1): only the block reduction is left (tensor-core path)
// VAL = false: the caller does not use the value (every sampler's gradient call): theta_net^2 and the value
// arithmetic are skipped, and so is the second likelihood sum when the likelihood is Normal
template <int NT, class SP = TcDyn, bool PASS = true, bool VAL = true>
BFX_DEV double eval_finish(const ModelDev& M, const ChainSmem& s, long long B, float nb, double ds0, double ds1,
                           bool want_grad, float& gn2, float a0 = 0.f, float a1 = 0.f);

template <int NT, int BT>
BFX_DEV double chain_eval(const ModelDev& M, const ChainSmem& s, const float* __restrict__ x,
                          const float* __restrict__ y, const BatchSrc& bs, long long B, float nb, bool want_grad,
                          float& gn2, float* yhat = nullptr) {
  if (want_grad) for_each_quad<NT>(M, [&](int k) { st4(s.g + k, make_float4(0.f, 0.f, 0.f, 0.f)); });
  const float sl = s.th[M.like_s];
  const float sigma = expf(sl);
  LikeCoef lc;
  lc.inv_sigma = 1.f / sigma;
  lc.nb = nb;
  double ds0 = 0.0, ds1 = 0.0;
  __syncthreads();
  if (!want_grad && !yhat && lstm_value_ok<NT>(M)) {  // forward-only LSTM pass: 64 sequences per tile (bfx_device.cuh)
    for (long long t0 = 0; t0 < B; t0 += 64) {
      const int nvalid = (int)((B - t0) < 64 ? (B - t0) : 64);
      float sums[2] = {0.f, 0.f};
      lstm_value_tile64<NT>(M, s, x, y, bs, t0, nvalid, lc, sums);
      ds0 += (double)sums[0];
      ds1 += (double)sums[1];
      __syncthreads();
    }
    return eval_finish<NT>(M, s, B, nb, ds0, ds1, false, gn2);
  }
  for (long long t0 = 0; t0 < B; t0 += BT) {
    int nvalid = (int)((B - t0) < (long long)BT ? (B - t0) : (long long)BT);
    float sums[2] = {0.f, 0.f};
    float* yt = yhat ? yhat + t0 : nullptr;
    if (M.rec_layer >= 0) rec_tile<NT, BT>(M, s, x, y, bs, t0, nvalid, lc, want_grad, sums, yt);
    else mlp_tile<NT, BT>(M, s, x, y, bs, t0, nvalid, lc, want_grad, sums, yt);
    ds0 += (double)sums[0];
    ds1 += (double)sums[1];
    __syncthreads();
  }
  return eval_finish<NT>(M, s, B, nb, ds0, ds1, want_grad, gn2);
}

// likelihood sums -> value; sigma prior; network prior; finishes the gradient in s.g (which on entry
// holds the gradient of  sum_b dl[b]*yhat[b]  w.r.t. theta_net) and its squared norm.
template <int NT, class SP, bool PASS, bool VAL>
BFX_DEV double eval_finish(const ModelDev& M, const ChainSmem& s, long long B, float nb, double ds0, double ds1,
                           bool want_grad, float& gn2, float a0, float a1) {
  const float sl = s.th[SP::like_s(M)];
  const float sigma = expf(sl);
  // prior over theta_net + the network part of the gradient; theta_like (element 0 of its own quad) is handled
  // after the reduction because its gradient needs the reduced likelihood sums
  float acc[2] = {a0, a1};  // sum theta_net^2, sum g_net^2
  const float c2 = M.prior_c2;
  const int like_q = SP::like_s(M);
  if (PASS) for_each_quad_n<NT, SP::fixed>(SP::S(M), [&](int k) {
    const unsigned m = s.mask[k >> 2];
    float4 th = ld4(s.th + k);
    if (k == like_q) th.x = 0.f;
    acc[0] += th.x * th.x + th.y * th.y + th.z * th.z + th.w * th.w;
    if (want_grad) {
      float4 g = mask4(ld4(s.g + k), m);
      g.x -= c2 * th.x; g.y -= c2 * th.y; g.z -= c2 * th.z; g.w -= c2 * th.w;
      if (k == like_q) g.x = 0.f;
      st4(s.g + k, g);
      acc[1] += g.x * g.x + g.y * g.y + g.z * g.z + g.w * g.w;
    }
  });
  // ONE block reduction for everything: the likelihood sums in double, the two norms in float
  const bool two = VAL || M.like_kind != LK_NORMAL;  // (the second likelihood sum only feeds the TDist gradient)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ds0 += __shfl_xor_sync(0xffffffffu, ds0, o);
    if (two) ds1 += __shfl_xor_sync(0xffffffffu, ds1, o);
    if (VAL) acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], o);
    acc[1] += __shfl_xor_sync(0xffffffffu, acc[1], o);
  }
  if (NT > 32) {
    double* dred = reinterpret_cast<double*>(s.red);  // 32*4 floats = 64 doubles: [w][0..1] doubles, floats behind
    float* fred = s.red + 4 * (NT / 32);               // after 2 doubles per warp
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    // (`red` is free here: every caller arrives through a barrier and the pass above does not touch it)
    if (lane == 0) {
      dred[2 * w] = ds0;
      dred[2 * w + 1] = ds1;
      fred[2 * w] = acc[0];
      fred[2 * w + 1] = acc[1];
    }
    __syncthreads();
    ds0 = 0.0;
    ds1 = 0.0;
    acc[0] = 0.f;
    acc[1] = 0.f;
#pragma unroll
    for (int ww = 0; ww < NT / 32; ++ww) {
      ds0 += dred[2 * ww];
      if (two) ds1 += dred[2 * ww + 1];
      if (VAL) acc[0] += fred[2 * ww];
      acc[1] += fred[2 * ww + 1];
    }
  }
  // sigma prior (log link): logpdf(prior, e^s) + s and its derivative w.r.t. s
  double lps = 0.0, dlps;
  if (M.sprior_kind == SP_GAMMA) {
    if (VAL) lps = M.sprior_const + ((double)M.sprior_a - 1.0) * (double)sl - (double)sigma / (double)M.sprior_b + (double)sl;
    dlps = (double)M.sprior_a - (double)sigma * M.sprior_rb;
  } else {
    if (VAL) lps = M.sprior_const - ((double)M.sprior_a + 1.0) * (double)sl - (double)M.sprior_b / (double)sigma + (double)sl;
    dlps = -(double)M.sprior_a + (double)M.sprior_b * (double)(1.f / sigma);
  }
  double like, dlike_ds;
  if (M.like_kind == LK_NORMAL) {
    like = -0.5 * (double)B * 1.8378770664093453 - 0.5 * ds0;
    dlike_ds = ds0;
  } else {
    like = (double)B * M.tdist_const - 0.5 * ((double)M.nu + 1.0) * ds0;
    dlike_ds = ds1;
  }
  like += -(double)B * (double)sl + lps;
  dlike_ds += -(double)B + dlps;
  const float glike = (float)((double)nb * dlike_ds);
  if (want_grad && threadIdx.x == 0) s.g[like_q] = glike;  // made visible by the barrier every consumer pass starts after
  gn2 = acc[1] + glike * glike;
  __syncthreads();
  return (double)nb * like + (M.prior_c0n - 0.5 * (double)c2 * (double)acc[0]);
}

// src/utils/gradient_utils.jl:6-10: ng > maxnorm ? maxnorm / ng : 1, as min(1, maxnorm / sqrt(gn2)) with one rsqrt
// (same limits: gn2 = 0 or NaN -> 1, gn2 = inf -> 0)
BFX_DEV float clip_scale(float gn2, float maxnorm) { return fminf(1.f, maxnorm * rsqrtf(gn2)); }

}  // namespace bfx
