// The five SG-MCMC update! rules as fused passes over a chain's parameters.
// Each step follows the cited reference lines statement by statement; all element-wise
// work of one update (preconditioning, momentum / thermostat, noise, step size, sample
// write) is done in at most two passes over theta with the reductions folded in.
#pragma once
#include "bfx_tc.cuh"

namespace bfx {

template <int NT, int BT, bool TC = false>
struct Ctx {
  const RunDev& R;
  tc::Region* tcr = nullptr;  // tensor-core region of this CTA (TC builds only)
  ChainSmem s;
  int c;  // local chain
  NoiseKey key;
  ChainScalars sc;
  float* th_g;  // HBM rows of this chain
  float* mom;
  float* minv;
  float* thold;
  float* rmsv;
  float* ring;
  BatchSrc bs;   // minibatch of the current step
  long long Bk;  // its size
  unsigned long long gstep;

  BFX_DEV Ctx(const RunDev& r, const ChainSmem& sm) : R(r), s(sm) {}

  BFX_DEV const float* inj(int slot) const {
    return R.inj_normals ? R.inj_normals + ((long long)slot * R.C + c) * R.M.P : nullptr;
  }
  BFX_DEV float uniform() const { return R.inj_uniforms ? R.inj_uniforms[c] : uniform_draw(key, gstep); }

  BFX_DEV double eval(const BatchSrc& b, long long Bn, float nbs, bool want_grad, float& gn2) {
    if constexpr (TC) return tc::chain_eval_tc<NT>(R.M, s, *tcr, R.x, R.y, b, Bn, nbs, want_grad, gn2);
    else return chain_eval<NT, BT>(R.M, s, R.x, R.y, b, Bn, nbs, want_grad, gn2);
  }
  BFX_DEV double grad(float& gn2) { return eval(bs, Bk, R.num_batches, true, gn2); }
  // loglikeprior(bnn, θ, bnn.x, bnn.y): full data, num_batches = 1 (ggmc.jl:147,176; hmc.jl:135-136)
  BFX_DEV double full_lp() {
    BatchSrc f = bs;
    f.full = 1;
    f.idx = nullptr;
    float dummy;
    return eval(f, R.n, 1.f, false, dummy);
  }
  BFX_DEV bool uses_minv() const { return R.S.madapter_kind != MA_FIXED; }
};

// block-wide OR of a per-thread flag
template <int NT>
BFX_DEV bool block_any(bool f, float* red) {
  float v[1] = {f ? 1.f : 0.f};
  block_sum<NT, 1>(v, red);
  return v[0] > 0.f;
}

// s.samples[:, t] = copy(θ)  (+ DiagCov window ring).  ns = nsampled AFTER the increment.
template <int NT, int BT, bool TC>
BFX_DEV void record_sample(Ctx<NT, BT, TC>& X, long long ns, const float* src_global /* nullptr: smem theta */) {
  const RunDev& R = X.R;
  float* dst = nullptr;
  if (R.samples && (ns % R.keep_every) == 0) {
    long long slot = ns / R.keep_every - 1 - R.kept_origin;
    if (R.ring_slots > 0) slot %= R.ring_slots;
    dst = R.samples + (long long)X.c * R.samples_chain_stride + slot * R.M.P;
  }
  float* rdst = nullptr;
  if (X.ring) rdst = X.ring + ((ns - 1) % R.S.ma_windowlength) * (long long)R.Ps;
  if (!dst && !rdst) return;
  for_each_param<NT>(R.M, [&](int J, int k) {
    float v = src_global ? src_global[J] : X.s.th[k];
    if (dst) dst[J] = v;
    if (rdst) rdst[J] = v;
  });
}

// DualAveragingStepSize / ConstantStepsize  (adapters/stepsize/*.jl)
BFX_DEV float stepsize_adapt(ChainScalars& sc, const SamplerDev& S, float mh_prob) {
  if (S.sadapter_kind == SA_CONST) return S.l;
  if (sc.da_t > S.da_adapt_steps) return expf(sc.da_log_avg);
  sc.da_error_sum += S.da_target_accept - mh_prob;
  float t = (float)sc.da_t;
  float log_step = S.da_mu - sc.da_error_sum / sqrtf(t * S.da_gamma);
  float eta = powf(t, -S.da_kappa);
  sc.da_log_avg = eta * log_step + (1.f - eta) * sc.da_log_avg;
  sc.da_t += 1;
  return expf(log_step);
}

BFX_DEV float mh_prob_of(float lMH) { return (lMH != lMH) ? lMH : fminf(expf(lMH), 1.f); }

// madapter(s, θ, bnn, ∇θ) -> Minv (diagonal).  The gradient RMSProp asks for is the one
// of the current minibatch at the current θ, which is what s.g already holds.
template <int NT, int BT, bool TC>
BFX_DEV void mass_adapt(Ctx<NT, BT, TC>& X, float gn2) {
  const RunDev& R = X.R;
  const SamplerDev& S = R.S;
  if (S.madapter_kind == MA_RMSPROP) {
    // adapters/mass/rmspropmassadapter.jl:26-46
    if (X.sc.ma_t > S.ma_adapt_steps) return;
    const bool first = X.sc.ma_t == 1;
    const float inv_ng = 1.f / sqrtf(gn2);
    for_each_param<NT>(R.M, [&](int J, int k) {
      float g = X.s.g[k] * inv_ng;
      float V = first ? 1.f : X.rmsv[J];
      V = S.ma_alpha * V + (1.f - S.ma_alpha) * g * g;
      X.rmsv[J] = V;
      X.minv[J] = 1.f / (S.ma_lambda + sqrtf(V));
    });
    X.sc.ma_t += 1;
  } else if (S.madapter_kind == MA_DIAGCOV) {
    // adapters/mass/diagcovariancemassadapter.jl:35-50
    if (X.sc.ma_t > S.ma_adapt_steps) return;
    const int w = S.ma_windowlength;
    if ((long long)w > X.sc.nsampled) return;
    for (int J = threadIdx.x; J < R.M.P; J += NT) {
      float m = 0.f;
      for (int i = 0; i < w; ++i) m += X.ring[(long long)i * R.Ps + J];
      m /= (float)w;
      float v = 0.f;
      for (int i = 0; i < w; ++i) {
        float dlt = X.ring[(long long)i * R.Ps + J] - m;
        v += dlt * dlt;
      }
      v /= (float)(w - 1);
      X.minv[J] = (1.f - S.ma_kappa) * v + S.ma_kappa + S.ma_epsilon;
    }
    X.sc.ma_t += 1;
  }
  __syncthreads();  // make Minv visible to later passes of this CTA (same-thread mapping differs)
}

// --------------------------------------------------------------------------------
// SGLD  (src/inference/mcmc/sgld.jl:96-112)
// --------------------------------------------------------------------------------
template <int NT, int BT, bool TC>
BFX_DEV void step_sgld(Ctx<NT, BT, TC>& X) {
  const SamplerDev& S = X.R.S;
  float alpha = S.stepsize_a * powf(S.stepsize_b + (float)X.sc.t, -S.stepsize_gamma);  // sgld.jl:12
  alpha = fmaxf(alpha, S.min_stepsize);
  float gn2;
  X.grad(gn2);
  const float ha = 0.5f * alpha * clip_scale(gn2, S.maxnorm), sa = sqrtf(alpha);
  float* th = X.s.th;
  const float* g = X.s.g;
  for_each_param_noise<NT>(X.R.M, X.key, X.gstep, 0, X.inj(0),
                           [&](int J, int k, float z) { th[k] += ha * g[k] + sa * z; });
  __syncthreads();
  X.sc.nsampled += 1;
  record_sample(X, X.sc.t, nullptr);  // samples[:, t] = θ
  X.sc.t += 1;
}

// --------------------------------------------------------------------------------
// SGNHT  (src/inference/mcmc/sgnht.jl:64-90)
// --------------------------------------------------------------------------------
template <int NT, int BT, bool TC>
BFX_DEV void step_sgnht(Ctx<NT, BT, TC>& X) {
  const RunDev& R = X.R;
  const SamplerDev& S = R.S;
  float gn2;
  X.grad(gn2);
  if (gn2 != gn2) {  // any(isnan.(g))
    X.sc.status = 4;
    return;
  }
  const float l = S.l, xi = X.sc.xi, cn = sqrtf(l * 2.f * S.A);
  float* th = X.s.th;
  const float* g = X.s.g;
  float acc[2] = {0.f, 0.f};  // p'p, nan flag
  for_each_param_noise<NT>(R.M, X.key, X.gstep, 0, X.inj(0), [&](int J, int k, float z) {
    float p = X.mom[J];
    p = p - xi * l * p + l * g[k] + cn * z;  // g here is +∇, the reference's g is -∇
    X.mom[J] = p;
    float t = th[k] + l * p;
    th[k] = t;
    acc[0] += p * p;
    acc[1] += (t != t) ? 1.f : 0.f;
  });
  block_sum<NT, 2>(acc, X.s.red);
  const float kin = acc[0] / (float)R.M.P;
  if (R.kinetic_out) {
    if (threadIdx.x == 0) R.kinetic_out[(long long)X.c * R.hist_stride + (X.sc.t - 1 - R.hist_origin)] = kin;
  }
  X.sc.xi = xi + (kin - 1.f) * l;
  if (acc[1] > 0.f) {
    X.sc.status = 5;
    return;
  }
  X.sc.nsampled += 1;
  record_sample(X, X.sc.t, nullptr);
  X.sc.t += 1;
}

// --------------------------------------------------------------------------------
// SGNHTS  (src/inference/mcmc/sgnht-s.jl:81-116), BAOAB-like splitting, diagonal Minv
// --------------------------------------------------------------------------------
template <int NT, int BT, bool TC>
BFX_DEV void step_sgnhts(Ctx<NT, BT, TC>& X) {
  const RunDev& R = X.R;
  const SamplerDev& S = R.S;
  float gn2;
  X.grad(gn2);  // ∇θ(θ); also the gradient the RMSProp adapter evaluates (:83, same θ and batch)
  mass_adapt(X, gn2);
  if (gn2 != gn2) {
    X.sc.status = 4;
    return;
  }
  const bool um = X.uses_minv();
  const float l = S.l, hl = 0.5f * l, n = (float)R.M.P;
  float* th = X.s.th;
  const float* g = X.s.g;
  float acc[2] = {0.f, 0.f};
  // B half kick + A half drift, accumulate p' Minv p   (:95-97)
  for_each_param<NT>(R.M, [&](int J, int k) {
    float mi = um ? X.minv[J] : 1.f;
    float p = X.mom[J] + hl * g[k];
    X.mom[J] = p;
    th[k] += hl * mi * p;
    acc[0] += p * p * mi;
  });
  block_sum<NT, 2>(acc, X.s.red);
  float xi = X.sc.xi + l / (2.f * S.mu) * (acc[0] - n);
  float e1, sc;
  if (xi != 0.f) {
    e1 = expf(-xi * l);
    sc = S.sigmaA * sqrtf((1.f - expf(-2.f * xi * l)) / (2.f * xi));
  } else {
    e1 = 1.f;
    sc = sqrtf(l) * S.sigmaA;
  }
  acc[0] = 0.f;
  acc[1] = 0.f;
  // O step with noise ~ N(0, M), second A half drift  (:98-104)
  for_each_param_noise<NT>(R.M, X.key, X.gstep, 0, X.inj(0), [&](int J, int k, float z) {
    float mi = um ? X.minv[J] : 1.f;
    float p = e1 * X.mom[J] + sc * rsqrtf(mi) * z;
    X.mom[J] = p;
    float t = th[k] + hl * mi * p;
    th[k] = t;
    acc[0] += p * p * mi;
    acc[1] += (t != t) ? 1.f : 0.f;
  });
  block_sum<NT, 2>(acc, X.s.red);
  X.sc.xi = xi + l / (2.f * S.mu) * (acc[0] - n);
  if (acc[1] > 0.f) {
    X.sc.status = 5;
    return;
  }
  if (R.kinetic_out && threadIdx.x == 0)
    R.kinetic_out[(long long)X.c * R.hist_stride + (X.sc.t - 1 - R.hist_origin)] = acc[0] / n;
  X.sc.nsampled += 1;
  record_sample(X, X.sc.t, nullptr);
  X.sc.t += 1;
}

// --------------------------------------------------------------------------------
// GGMC  (src/inference/mcmc/ggmc.jl:140-198), diagonal mass
// --------------------------------------------------------------------------------
template <int NT, int BT, bool TC>
BFX_DEV void step_ggmc(Ctx<NT, BT, TC>& X) {
  const RunDev& R = X.R;
  const SamplerDev& S = R.S;
  const float N = (float)R.n;
  const float lcur = X.sc.l;
  const float gam = -sqrtf(N / lcur) * logf(S.beta);
  float h = sqrtf(lcur / N);
  const float a = expf(-gam * h);
  float* th = X.s.th;
  const float* g = X.s.g;
  if (X.sc.current_step == 1) {
    X.sc.lMH = (float)(-X.full_lp());  // :146-148
    // θ at the start of the window = samples[:, nsampled - steps] on a later reject
    for_each_param<NT>(R.M, [&](int J, int k) { X.thold[J] = th[k]; });
  }
  float gn2;
  X.grad(gn2);
  float cs = clip_scale(gn2, S.maxnorm);
  h = sqrtf(h);  // :154
  const float sa = sqrtf(a), s1a = sqrtf(1.f - a), hh = 0.5f * h;
  const bool um = X.uses_minv();
  float acc[2] = {0.f, 0.f};  // K(m14)*2, nan
  for_each_param_noise<NT>(R.M, X.key, X.gstep, 0, X.inj(0), [&](int J, int k, float z) {
    float mi = um ? X.minv[J] : 1.f;
    float m14 = sa * X.mom[J] + s1a * rsqrtf(mi) * z;  // Mhalf = chol(inv(Minv)) = 1/sqrt(Minv)
    float m12 = m14 + hh * cs * g[k];                  // reference g = -∇
    X.mom[J] = m12;
    float t = th[k] + h * mi * m12;
    th[k] = t;
    acc[0] += m14 * m14 * mi;
    acc[1] += (t != t) ? 1.f : 0.f;
  });
  block_sum<NT, 2>(acc, X.s.red);
  if (acc[1] > 0.f) {
    X.sc.status = 5;
    return;
  }
  const float K14 = 0.5f * acc[0];
  X.grad(gn2);  // same minibatch, new θ  (:164-166)
  cs = clip_scale(gn2, S.maxnorm);
  acc[0] = 0.f;
  for_each_param_noise<NT>(R.M, X.key, X.gstep, 1, X.inj(1), [&](int J, int k, float z) {
    float mi = um ? X.minv[J] : 1.f;
    float m34 = X.mom[J] + hh * cs * g[k];
    X.mom[J] = sa * m34 + s1a * rsqrtf(mi) * z;
    acc[0] += m34 * m34 * mi;
  });
  block_sum<NT, 2>(acc, X.s.red);
  X.sc.lMH += 0.5f * acc[0] - K14;  // :171
  X.sc.nsampled += 1;
  const long long ns = X.sc.nsampled;
  record_sample(X, ns, nullptr);  // samples[:, t] with t == nsampled for GGMC
  if (X.sc.current_step == S.steps && X.sc.t > S.steps) {
    __syncthreads();
    X.sc.lMH += (float)X.full_lp();  // :176
    const float pacc = mh_prob_of(X.sc.lMH);
    X.sc.l = stepsize_adapt(X.sc, S, pacc);
    mass_adapt(X, gn2);
    const float r = X.uniform();
    const bool accept = r < pacc;
    if (R.accepted_out && threadIdx.x == 0) {
      for (int j = 0; j < S.steps; ++j)
        R.accepted_out[(long long)X.c * R.hist_stride + (ns - 1 - j - R.hist_origin)] = accept ? 1 : 0;
    }
    if (accept) {
      X.sc.accepted_total += S.steps;
    } else {
      __syncthreads();
      for_each_param<NT>(R.M, [&](int J, int k) { th[k] = X.thold[J]; });
      __syncthreads();
      for (int j = 0; j < S.steps; ++j) record_sample(X, ns - j, nullptr);
    }
  }
  X.sc.t += 1;
  X.sc.current_step = (X.sc.current_step == S.steps) ? 1 : X.sc.current_step + 1;
}

// --------------------------------------------------------------------------------
// HMC  (src/inference/mcmc/hmc.jl:105-159): one call = one leapfrog stage
// --------------------------------------------------------------------------------
template <int NT, int BT, bool TC>
BFX_DEV void step_hmc(Ctx<NT, BT, TC>& X) {
  const RunDev& R = X.R;
  const SamplerDev& S = R.S;
  const float l = X.sc.l;
  float* th = X.s.th;
  const float* g = X.s.g;
  const bool um = X.uses_minv();
  const int stage = X.sc.current_step;
  if (stage == 1) {
    // new trajectory: p ~ N(0, Minv); θold, pold kept for the MH test (:117-121).
    // -loglikeprior(θold) and logpdf(moment_dist, pold) are evaluated here, while θ == θold.
    X.sc.start_lp = X.full_lp();
    float acc[1] = {0.f};
    for_each_param_noise<NT>(R.M, X.key, X.gstep, 0, X.inj(0), [&](int J, int k, float z) {
      float mi = um ? X.minv[J] : 1.f;
      float p = sqrtf(mi) * z;
      X.mom[J] = p;
      X.thold[J] = th[k];
      acc[0] += p * p / mi;
    });
    block_sum<NT, 1>(acc, X.s.red);
    X.sc.lMH = 0.5f * acc[0];  // -logpdf(moment_dist, pold) up to the constant that cancels
  } else {
    for_each_param<NT>(R.M, [&](int J, int k) { th[k] += l * X.mom[J]; });  // :124,128
  }
  __syncthreads();
  float gn2;
  X.grad(gn2);
  const float hk = 0.5f * l * clip_scale(gn2, S.maxnorm);
  const bool last = stage == S.path_len;
  float acc[1] = {0.f};
  for_each_param<NT>(R.M, [&](int J, int k) {
    float p = X.mom[J] + hk * g[k];  // momentum .-= l*(-∇)/2
    if (last) {
      p = -p;  // :133
      float mi = um ? X.minv[J] : 1.f;
      acc[0] += p * p / mi;
    }
    X.mom[J] = p;
  });
  if (last) {
    block_sum<NT, 1>(acc, X.s.red);
    const double start_mh = -X.sc.start_lp + (double)X.sc.lMH;
    const double end_mh = -X.full_lp() + 0.5 * (double)acc[0];
    const float lMH = (float)(start_mh - end_mh);
    const float lr = logf(X.uniform());
    X.sc.l = stepsize_adapt(X.sc, S, mh_prob_of(lMH));
    mass_adapt(X, gn2);
    const bool accept = lr < lMH;
    const long long ns = X.sc.nsampled + 1;
    if (R.accepted_out && threadIdx.x == 0)
      R.accepted_out[(long long)X.c * R.hist_stride + (ns - 1 - R.hist_origin)] = accept ? 1 : 0;
    if (accept) {
      X.sc.accepted_total += 1;
    } else {
      __syncthreads();
      for_each_param<NT>(R.M, [&](int J, int k) { th[k] = X.thold[J]; });
      __syncthreads();
    }
    X.sc.nsampled = ns;
    record_sample(X, ns, nullptr);
  }
  X.sc.current_step = last ? 1 : stage + 1;
  X.sc.t += 1;
}

}  // namespace bfx
