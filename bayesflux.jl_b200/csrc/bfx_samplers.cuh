// The five SG-MCMC update! rules as fused passes over a chain's parameters.
// Each step follows the cited reference lines statement by statement; all element-wise
// work of one update (preconditioning, momentum / thermostat, noise, step size, sample
// write) is done in at most two passes over theta with the reductions folded in.
#pragma once
#include "bfx_tc.cuh"

namespace bfx {

template <int NT, int BT, bool TC = false, class SP = TcDyn>
struct Ctx {
  using Shape = SP;
  const RunDev& R;
  tc::Region* tcr = nullptr;  // tensor-core region of this CTA (TC builds only)
  ChainSmem s;
  int c;  // local chain
  NoiseKey key;
  ChainScalars sc;
  float* th_g;  // HBM rows of this chain (padded layout, S floats each)
  float* mom;
  float* minv;
  float* thold;
  float* rmsv;
  float* ring;
  float* window;  // find_mode convergence window of this chain
  BatchSrc bs;   // minibatch of the current step
  long long Bk;  // its size
  unsigned long long gstep;

  BFX_DEV Ctx(const RunDev& r, const ChainSmem& sm) : R(r), s(sm) {}

  BFX_DEV const float* inj(int slot) const {
    return R.inj_normals ? R.inj_normals + ((long long)slot * R.C + c) * R.M.P : nullptr;
  }
  BFX_DEV float uniform() const { return R.inj_uniforms ? R.inj_uniforms[c] : uniform_draw(key, gstep); }

  template <bool VAL = true>
  BFX_DEV double eval(const BatchSrc& b, long long Bn, float nbs, bool want_grad, float& gn2) {
    if constexpr (TC)
      return tc::chain_eval_tc<NT, tc::NoShadow, SP, VAL>(R.M, s, *tcr, R.x, R.xs, R.y, b, Bn, nbs, want_grad, gn2);
    else return chain_eval<NT, BT>(R.M, s, R.x, R.y, b, Bn, nbs, want_grad, gn2);
  }
  // gradient into s.g and its squared norm (no sampler uses the value of a gradient call)
  BFX_DEV void grad(float& gn2) { eval<false>(bs, Bk, R.num_batches, true, gn2); }
  // tensor-core builds: the gradient with a shadow job (tc::chain_eval_tc)
  template <class Shadow>
  BFX_DEV void grad_shadow(float& gn2, Shadow job) {
    static_assert(TC, "shadow jobs exist on the tensor-core path only");
    tc::chain_eval_tc<NT, Shadow, SP, false>(R.M, s, *tcr, R.x, R.xs, R.y, bs, Bk, R.num_batches, true, gn2, job);
  }
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

// where the ns-th sample of this chain is kept (flat row of P floats), or nullptr when it is not kept
template <int NT, int BT, bool TC, class SP>
BFX_DEV float* sample_slot(const Ctx<NT, BT, TC, SP>& X, long long ns) {
  const RunDev& R = X.R;
  if (!R.samples) return nullptr;
  unsigned long long q, r;
  fdivmod(R.fd_keep, (unsigned long long)ns, q, r);
  if (r != 0ull) return nullptr;
  long long slot = (long long)q - 1 - R.kept_origin;
  // a slot outside the buffer of this launch is never written (stale host counters must not turn into stray stores)
  if (slot < 0) return nullptr;
  if (R.ring_slots <= 0 && (slot + 1) * (long long)R.M.P > R.samples_chain_stride) return nullptr;
  if (R.ring_slots > 0) {
    unsigned long long sq, sr;
    fdivmod(R.fd_slots, (unsigned long long)slot, sq, sr);
    slot = (long long)sr;
  }
  return R.samples + (long long)X.c * R.samples_chain_stride + slot * R.M.P;
}
// entry of the per-step history arrays (kinetic_out / accepted_out) for 1-based step i, or -1 outside this run
BFX_DEV long long hist_index(const RunDev& R, int c, long long i) {
  const long long k = i - 1 - R.hist_origin;
  return (k < 0 || k >= R.hist_stride) ? -1 : (long long)c * R.hist_stride + k;
}
// row of the DiagCov window ring (padded layout) that takes the ns-th sample
template <int NT, int BT, bool TC, class SP>
BFX_DEV float* window_slot(const Ctx<NT, BT, TC, SP>& X, long long ns) {
  unsigned long long q, r;
  fdivmod(X.R.fd_window, (unsigned long long)(ns - 1), q, r);
  return X.ring + (long long)r * (long long)X.R.Ps;
}

// The same recording fused into the pass that produces the new theta (SGLD / SGNHT / SGNHTS): the thread that
// updates padded quad k also writes it out, so the separate pass over theta and its barrier disappear.
struct Recorder {
  float* dst;   // flat sample slot of this chain for this step, or nullptr
  float* rdst;  // DiagCov window slot (padded), or nullptr
};

// J0 = flat index of element 0 of quad k (the valid elements of a quad are consecutive in the flat order: padding
// only ever follows the last row of a column)
BFX_DEV void rec_store(const Recorder& r, int k, unsigned m, const float4& t, int J0) {
  if (r.rdst) st4(r.rdst + k, t);
  if (r.dst) {
    float* p = r.dst + J0;
    if (m == 15u && (reinterpret_cast<unsigned long long>(p) & 15ull) == 0) {
      st4(p, t);
    } else if (m == 15u && (reinterpret_cast<unsigned long long>(p) & 7ull) == 0) {
      *reinterpret_cast<float2*>(p) = make_float2(t.x, t.y);
      *reinterpret_cast<float2*>(p + 2) = make_float2(t.z, t.w);
    } else {
      if (m & 1u) p[0] = t.x;
      if (m & 2u) p[1] = t.y;
      if (m & 4u) p[2] = t.z;
      if (m & 8u) p[3] = t.w;
    }
  }
}

// s.samples[:, t] = copy(θ)  (+ DiagCov window ring).  ns = nsampled AFTER the increment.
// Recorded samples leave the engine in the reference's flat order; the window ring is internal (padded).
template <int NT, int BT, bool TC, class SP>
BFX_DEV void record_sample(Ctx<NT, BT, TC, SP>& X, long long ns) {
  const RunDev& R = X.R;
  float* dst = sample_slot(X, ns);
  if (dst) {
    // quad order with the shared-memory quad -> flat table: vector stores, no global table lookups
    const float* th = X.s.th;
    const Recorder rec{dst, nullptr};
    for_each_quad<NT>(R.M, [&](int k) {
      const unsigned m = X.s.mask[k >> 2];
      if (m) rec_store(rec, k, m, ld4(th + k), X.s.qflat[k >> 2]);
    });
  }
  if (X.ring) {
    float* rdst = window_slot(X, ns);
    for_each_quad<NT>(R.M, [&](int k) { st4(rdst + k, ld4(X.s.th + k)); });
  }
}

template <int NT, int BT, bool TC, class SP>
BFX_DEV Recorder make_recorder(Ctx<NT, BT, TC, SP>& X, long long ns) {
  Recorder r{sample_slot(X, ns), nullptr};
  if (X.ring) r.rdst = window_slot(X, ns);
  return r;
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

#define BFX_Q4(EXPR_X, EXPR_Y, EXPR_Z, EXPR_W) make_float4(EXPR_X, EXPR_Y, EXPR_Z, EXPR_W)
// apply `expr` (written in terms of component c) to the four components
#define BFX_MAP4(OUT, expr)            \
  {                                    \
    { constexpr int c = 0; OUT.x = expr; } \
    { constexpr int c = 1; OUT.y = expr; } \
    { constexpr int c = 2; OUT.z = expr; } \
    { constexpr int c = 3; OUT.w = expr; } \
  }
BFX_DEV float comp(const float4& v, int c) { return c == 0 ? v.x : c == 1 ? v.y : c == 2 ? v.z : v.w; }
BFX_DEV float hsum(const float4& v) { return (v.x + v.y) + (v.z + v.w); }
BFX_DEV float4 ones4() { return make_float4(1.f, 1.f, 1.f, 1.f); }
// padding slots of an inverse mass vector hold 1 (never 0: they are divided by)
BFX_DEV float4 mask4_one(float4 v, unsigned m) {
  return make_float4((m & 1u) ? v.x : 1.f, (m & 2u) ? v.y : 1.f, (m & 4u) ? v.z : 1.f, (m & 8u) ? v.w : 1.f);
}
BFX_DEV bool any_nan(const float4& v) { return v.x != v.x || v.y != v.y || v.z != v.z || v.w != v.w; }

// madapter(s, θ, bnn, ∇θ) -> Minv (diagonal).  The gradient RMSProp asks for is the one
// of the current minibatch at the current θ, which is what s.g already holds.
template <int NT, int BT, bool TC, class SP>
BFX_DEV void mass_adapt(Ctx<NT, BT, TC, SP>& X, float gn2) {
  const RunDev& R = X.R;
  const SamplerDev& S = R.S;
  if (S.madapter_kind == MA_RMSPROP) {
    // adapters/mass/rmspropmassadapter.jl:26-46
    if (X.sc.ma_t > S.ma_adapt_steps) return;
    const bool first = X.sc.ma_t == 1;
    const float inv_ng = 1.f / sqrtf(gn2);
    for_each_quad<NT>(R.M, [&](int k) {
      const unsigned m = X.s.mask[k >> 2];
      const float4 g = ld4(X.s.g + k);
      const float4 V0 = first ? ones4() : gld4(X.rmsv + k);
      float4 V, mi;
      BFX_MAP4(V, S.ma_alpha * comp(V0, c) + (1.f - S.ma_alpha) * (comp(g, c) * inv_ng) * (comp(g, c) * inv_ng));
      BFX_MAP4(mi, 1.f / (S.ma_lambda + sqrtf(comp(V, c))));
      st4(X.rmsv + k, mask4(V, m));
      st4(X.minv + k, mask4_one(mi, m));
    });
    X.sc.ma_t += 1;
  } else if (S.madapter_kind == MA_DIAGCOV) {
    // adapters/mass/diagcovariancemassadapter.jl:35-50
    if (X.sc.ma_t > S.ma_adapt_steps) return;
    const int w = S.ma_windowlength;
    if ((long long)w > X.sc.nsampled) return;
    for_each_quad<NT>(R.M, [&](int k) {
      const unsigned m = X.s.mask[k >> 2];
      float4 mean = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int i = 0; i < w; ++i) {
        const float4 v = gld4(X.ring + (long long)i * R.Ps + k);
        mean.x += v.x; mean.y += v.y; mean.z += v.z; mean.w += v.w;
      }
      const float rw = 1.f / (float)w;
      mean.x *= rw; mean.y *= rw; mean.z *= rw; mean.w *= rw;
      float4 var = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int i = 0; i < w; ++i) {
        const float4 v = gld4(X.ring + (long long)i * R.Ps + k);
        var.x += (v.x - mean.x) * (v.x - mean.x); var.y += (v.y - mean.y) * (v.y - mean.y);
        var.z += (v.z - mean.z) * (v.z - mean.z); var.w += (v.w - mean.w) * (v.w - mean.w);
      }
      const float rv = 1.f / (float)(w - 1);
      float4 mi;
      BFX_MAP4(mi, (1.f - S.ma_kappa) * (comp(var, c) * rv) + S.ma_kappa + S.ma_epsilon);
      st4(X.minv + k, mask4_one(mi, m));
    });
    X.sc.ma_t += 1;
  }
  __syncthreads();  // make Minv visible to later passes of this CTA
}

// --------------------------------------------------------------------------------
// SGLD  (src/inference/mcmc/sgld.jl:96-112)
// --------------------------------------------------------------------------------
template <int NT, int BT, bool TC, class SP>
BFX_DEV void step_sgld(Ctx<NT, BT, TC, SP>& X) {
  const SamplerDev& S = X.R.S;
  // sgld.jl:12  a (b + t)^(-gamma): exp2(-gamma log2(b + t)) with the full-precision log2f / exp2f, within 2e-6
  // relative of powf for t < 2^24 and a third of its instructions
  auto stepsize = [&]() {
    return fmaxf(S.stepsize_a * exp2f(-S.stepsize_gamma * log2f(S.stepsize_b + (float)X.sc.t)), S.min_stepsize);
  };
  float gn2, alpha;
  Recorder rec;  // samples[:, t] = θ, written by the update pass itself
  if constexpr (TC) {
    // step size and sample slot do not depend on the gradient: warp 1 works them out while warp 0 issues the first
    // MMAs, everybody picks them up from shared memory afterwards
    float* sh = X.tcr->shadow;
    X.grad_shadow(gn2, [&]() {
      const float a = stepsize();
      const Recorder r = make_recorder(X, X.sc.t);
      if ((threadIdx.x & 31) == 0) {
        sh[0] = a;
        *reinterpret_cast<float**>(sh + 2) = r.dst;
        *reinterpret_cast<float**>(sh + 4) = r.rdst;
      }
    });
    alpha = sh[0];
    rec.dst = *reinterpret_cast<float* const*>(sh + 2);
    rec.rdst = *reinterpret_cast<float* const*>(sh + 4);
  } else {
    alpha = stepsize();
    X.grad(gn2);
    rec = make_recorder(X, X.sc.t);
  }
  const float ha = 0.5f * alpha * clip_scale(gn2, S.maxnorm), sa = sqrtf(alpha);
  float* th = X.s.th;
  const float* g = X.s.g;
  const float* inj = X.inj(0);
  X.sc.nsampled += 1;
  BFX_TICK(14);
  if (!inj && X.R.M.philox_rk_set && !rec.rdst) {
    // the common case (Philox noise, no DiagCov window) with everything that does not depend on the quad hoisted: the
    // counter words, the alignment of the sample row (19 720-byte rows: every other slot is only 8-byte aligned)
    const unsigned c1 = (unsigned)X.gstep, c2 = X.key.chain, c3 = (unsigned)((X.gstep >> 32) & 0xFFFFFFu);
    float* const dst = rec.dst;
    const unsigned dmis = (unsigned)(reinterpret_cast<unsigned long long>(dst) & 15ull);
    for_each_quad_n<NT, SP::fixed>(SP::S(X.R.M), [&](int k) {
      const unsigned m = X.s.mask[k >> 2];
      if (!m) return;
      const uint4 r = philox4x32_10_rk(make_uint4((unsigned)(k >> 2), c1, c2, c3), X.R.M.philox_rk);
      const float2 za = box_muller(r.x, r.y), zb = box_muller(r.z, r.w);
      const float4 t = ld4(th + k), gg = ld4(g + k);
      float4 o = make_float4(t.x + ha * gg.x + sa * za.x, t.y + ha * gg.y + sa * za.y, t.z + ha * gg.z + sa * zb.x,
                             t.w + ha * gg.w + sa * zb.y);
      if (m != 15u) o = mask4(o, m);
      st4(th + k, o);
      if (dst) {
        const int J0 = X.s.qflat[k >> 2];
        float* p = dst + J0;
        const unsigned mis = (dmis + 4u * (unsigned)J0) & 15u;
        if (m == 15u && mis == 0u) {
          st4(p, o);
        } else if (m == 15u && (mis & 7u) == 0u) {
          *reinterpret_cast<float2*>(p) = make_float2(o.x, o.y);
          *reinterpret_cast<float2*>(p + 2) = make_float2(o.z, o.w);
        } else {
          if (m & 1u) p[0] = o.x;
          if (m & 2u) p[1] = o.y;
          if (m & 4u) p[2] = o.z;
          if (m & 8u) p[3] = o.w;
        }
      }
    });
  } else {
    for_each_quad_n<NT, SP::fixed>(SP::S(X.R.M), [&](int k) {
      const unsigned m = X.s.mask[k >> 2];
      if (!m) return;
      const int J0 = X.s.qflat[k >> 2];
      const float4 z = noise_quad(X.R.M, X.key, X.gstep, 0, inj, k);
      const float4 t = ld4(th + k), gg = ld4(g + k);
      float4 o;
      BFX_MAP4(o, comp(t, c) + ha * comp(gg, c) + sa * comp(z, c));
      if (m != 15u) o = mask4(o, m);
      st4(th + k, o);
      rec_store(rec, k, m, o, J0);
    });
  }
  BFX_TICK(15);
  __syncthreads();
  X.sc.t += 1;
}

// --------------------------------------------------------------------------------
// SGNHT  (src/inference/mcmc/sgnht.jl:64-90)
// --------------------------------------------------------------------------------
template <int NT, int BT, bool TC, class SP>
BFX_DEV void step_sgnht(Ctx<NT, BT, TC, SP>& X) {
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
  const float* inj = X.inj(0);
  float acc[2] = {0.f, 0.f};  // p'p, nan flag
  // the sample is written by the pass that produces it (a chain that turns NaN here is reported as an error, its
  // slot is undefined either way)
  const Recorder rec = make_recorder(X, X.sc.t);
  for_each_quad<NT>(R.M, [&](int k) {
    const unsigned m = X.s.mask[k >> 2];
    if (!m) return;
    const float4 z = noise_quad(R.M, X.key, X.gstep, 0, inj, k);
    const float4 p0 = gld4(X.mom + k), gg = ld4(g + k), t0 = ld4(th + k);
    float4 p, t;
    BFX_MAP4(p, comp(p0, c) - xi * l * comp(p0, c) + l * comp(gg, c) + cn * comp(z, c));  // g here is +∇, the reference's g is -∇
    p = mask4(p, m);
    BFX_MAP4(t, comp(t0, c) + l * comp(p, c));
    st4(X.mom + k, p);
    st4(th + k, t);
    rec_store(rec, k, m, t, X.s.qflat[k >> 2]);
    acc[0] += p.x * p.x + p.y * p.y + p.z * p.z + p.w * p.w;
    acc[1] += any_nan(t) ? 1.f : 0.f;
  });
  block_sum<NT, 2>(acc, X.s.red);
  const float kin = acc[0] / (float)R.M.P;
  if (R.kinetic_out) {
    const long long hi_ = hist_index(R, X.c, X.sc.t);
    if (threadIdx.x == 0 && hi_ >= 0) R.kinetic_out[hi_] = kin;
  }
  X.sc.xi = xi + (kin - 1.f) * l;
  if (acc[1] > 0.f) {
    X.sc.status = 5;
    return;
  }
  X.sc.nsampled += 1;
  X.sc.t += 1;
}

// --------------------------------------------------------------------------------
// SGNHTS  (src/inference/mcmc/sgnht-s.jl:81-116), BAOAB-like splitting, diagonal Minv
// --------------------------------------------------------------------------------
template <int NT, int BT, bool TC, class SP>
BFX_DEV void step_sgnhts(Ctx<NT, BT, TC, SP>& X) {
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
  const float* inj = X.inj(0);
  float acc[2] = {0.f, 0.f};
  // B half kick + A half drift, accumulate p' Minv p   (:95-97)
  for_each_quad<NT>(R.M, [&](int k) {
    const unsigned m = X.s.mask[k >> 2];
    if (!m) return;
    const float4 mi = um ? gld4(X.minv + k) : ones4();
    const float4 p0 = gld4(X.mom + k), gg = ld4(g + k), t0 = ld4(th + k);
    float4 p, t;
    BFX_MAP4(p, comp(p0, c) + hl * comp(gg, c));
    p = mask4(p, m);
    BFX_MAP4(t, comp(t0, c) + hl * comp(mi, c) * comp(p, c));
    st4(X.mom + k, p);
    st4(th + k, t);
    acc[0] += p.x * p.x * mi.x + p.y * p.y * mi.y + p.z * p.z * mi.z + p.w * p.w * mi.w;
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
  const Recorder rec = make_recorder(X, X.sc.t);  // the sample is written by the pass that finishes theta
  // O step with noise ~ N(0, M), second A half drift  (:98-104)
  for_each_quad<NT>(R.M, [&](int k) {
    const unsigned m = X.s.mask[k >> 2];
    if (!m) return;
    const float4 z = noise_quad(R.M, X.key, X.gstep, 0, inj, k);
    const float4 mi = um ? gld4(X.minv + k) : ones4();
    const float4 p0 = gld4(X.mom + k), t0 = ld4(th + k);
    float4 p, t;
    BFX_MAP4(p, e1 * comp(p0, c) + sc * rsqrtf(comp(mi, c)) * comp(z, c));
    p = mask4(p, m);
    BFX_MAP4(t, comp(t0, c) + hl * comp(mi, c) * comp(p, c));
    st4(X.mom + k, p);
    st4(th + k, t);
    rec_store(rec, k, m, t, X.s.qflat[k >> 2]);
    acc[0] += p.x * p.x * mi.x + p.y * p.y * mi.y + p.z * p.z * mi.z + p.w * p.w * mi.w;
    acc[1] += any_nan(t) ? 1.f : 0.f;
  });
  block_sum<NT, 2>(acc, X.s.red);
  X.sc.xi = xi + l / (2.f * S.mu) * (acc[0] - n);
  if (acc[1] > 0.f) {
    X.sc.status = 5;
    return;
  }
  if (R.kinetic_out && threadIdx.x == 0) {
    const long long hi_ = hist_index(R, X.c, X.sc.t);
    if (hi_ >= 0) R.kinetic_out[hi_] = acc[0] / n;
  }
  X.sc.nsampled += 1;
  X.sc.t += 1;
}

// --------------------------------------------------------------------------------
// GGMC  (src/inference/mcmc/ggmc.jl:140-198), diagonal mass
// --------------------------------------------------------------------------------
template <int NT, int BT, bool TC, class SP>
BFX_DEV void step_ggmc(Ctx<NT, BT, TC, SP>& X) {
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
    // :146-148  lMH = -lπ_full(θ): the log posterior of the full data set is O(n) large, so the window's start value
    // stays in double (start_lp) and only the kinetic-energy terms accumulate in the float lMH
    if (X.sc.lp_epoch != R.data_epoch + 1) X.sc.start_lp = X.full_lp();  // (else: left behind by the last MH test)
    X.sc.lp_epoch = 0;
    X.sc.lMH = 0.f;
    // θ at the start of the window = samples[:, nsampled - steps] on a later reject
    for_each_quad<NT>(R.M, [&](int k) { st4(X.thold + k, ld4(th + k)); });
  }
  float gn2;
  X.grad(gn2);
  float cs = clip_scale(gn2, S.maxnorm);
  h = sqrtf(h);  // :154
  const float sa = sqrtf(a), s1a = sqrtf(1.f - a), hh = 0.5f * h;
  const bool um = X.uses_minv();
  float acc[2] = {0.f, 0.f};  // K(m14)*2, nan
  {
    const float* inj = X.inj(0);
    for_each_quad<NT>(R.M, [&](int k) {
      const unsigned m = X.s.mask[k >> 2];
      if (!m) return;
      const float4 z = noise_quad(R.M, X.key, X.gstep, 0, inj, k);
      const float4 mi = um ? gld4(X.minv + k) : ones4();
      const float4 m0 = gld4(X.mom + k), gg = ld4(g + k), t0 = ld4(th + k);
      float4 m14, m12, t;
      BFX_MAP4(m14, sa * comp(m0, c) + s1a * rsqrtf(comp(mi, c)) * comp(z, c));  // Mhalf = chol(inv(Minv)) = 1/sqrt(Minv)
      m14 = mask4(m14, m);
      BFX_MAP4(m12, comp(m14, c) + hh * cs * comp(gg, c));  // reference g = -∇
      m12 = mask4(m12, m);
      BFX_MAP4(t, comp(t0, c) + h * comp(mi, c) * comp(m12, c));
      st4(X.mom + k, m12);
      st4(th + k, t);
      acc[0] += m14.x * m14.x * mi.x + m14.y * m14.y * mi.y + m14.z * m14.z * mi.z + m14.w * m14.w * mi.w;
      acc[1] += any_nan(t) ? 1.f : 0.f;
    });
  }
  block_sum<NT, 2>(acc, X.s.red);
  if (acc[1] > 0.f) {
    X.sc.status = 5;
    return;
  }
  const float K14 = 0.5f * acc[0];
  X.grad(gn2);  // same minibatch, new θ  (:164-166)
  cs = clip_scale(gn2, S.maxnorm);
  acc[0] = 0.f;
  {
    const float* inj = X.inj(1);
    for_each_quad<NT>(R.M, [&](int k) {
      const unsigned m = X.s.mask[k >> 2];
      if (!m) return;
      const float4 z = noise_quad(R.M, X.key, X.gstep, 1, inj, k);
      const float4 mi = um ? gld4(X.minv + k) : ones4();
      const float4 m0 = gld4(X.mom + k), gg = ld4(g + k);
      float4 m34, mn;
      BFX_MAP4(m34, comp(m0, c) + hh * cs * comp(gg, c));
      m34 = mask4(m34, m);
      BFX_MAP4(mn, sa * comp(m34, c) + s1a * rsqrtf(comp(mi, c)) * comp(z, c));
      st4(X.mom + k, mask4(mn, m));
      acc[0] += m34.x * m34.x * mi.x + m34.y * m34.y * mi.y + m34.z * m34.z * mi.z + m34.w * m34.w * mi.w;
    });
  }
  block_sum<NT, 2>(acc, X.s.red);
  X.sc.lMH += 0.5f * acc[0] - K14;  // :171
  X.sc.nsampled += 1;
  const long long ns = X.sc.nsampled;
  record_sample(X, ns);  // samples[:, t] with t == nsampled for GGMC
  if (X.sc.current_step == S.steps && X.sc.t > S.steps) {
    __syncthreads();
    // (the mass adapter reads the minibatch gradient in s.g, which a value-only evaluation may use as workspace: it runs
    // before the full-data pass; nothing it changes enters the acceptance test)
    mass_adapt(X, gn2);
    const double end_lp = X.full_lp();
    X.sc.lMH = (float)((end_lp - X.sc.start_lp) + (double)X.sc.lMH);  // :176, the difference formed in double
    const float pacc = mh_prob_of(X.sc.lMH);
    X.sc.l = stepsize_adapt(X.sc, S, pacc);
    const float r = X.uniform();
    const bool accept = r < pacc;
    if (R.accepted_out && threadIdx.x == 0) {
      for (int j = 0; j < S.steps; ++j) {
        const long long hi_ = hist_index(R, X.c, ns - j);
        if (hi_ >= 0) R.accepted_out[hi_] = accept ? 1 : 0;
      }
    }
    // the next window starts from theta (accept) or from theta_old (reject): its full-data log posterior is known
    if (accept) X.sc.start_lp = end_lp;
    X.sc.lp_epoch = R.data_epoch + 1;
    if (accept) {
      X.sc.accepted_total += S.steps;
    } else {
      __syncthreads();
      for_each_quad<NT>(R.M, [&](int k) { st4(th + k, gld4(X.thold + k)); });
      __syncthreads();
      for (int j = 0; j < S.steps; ++j) record_sample(X, ns - j);
    }
  }
  X.sc.t += 1;
  X.sc.current_step = (X.sc.current_step == S.steps) ? 1 : X.sc.current_step + 1;
}

// --------------------------------------------------------------------------------
// HMC  (src/inference/mcmc/hmc.jl:105-159): one call = one leapfrog stage
// --------------------------------------------------------------------------------
template <int NT, int BT, bool TC, class SP>
BFX_DEV void step_hmc(Ctx<NT, BT, TC, SP>& X) {
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
    if (X.sc.lp_epoch != R.data_epoch + 1) X.sc.start_lp = X.full_lp();  // (else: left behind by the last MH test)
    X.sc.lp_epoch = 0;
    float acc[1] = {0.f};
    const float* inj = X.inj(0);
    for_each_quad<NT>(R.M, [&](int k) {
      const unsigned m = X.s.mask[k >> 2];
      st4(X.thold + k, ld4(th + k));
      if (!m) return;
      const float4 z = noise_quad(R.M, X.key, X.gstep, 0, inj, k);
      const float4 mi = um ? gld4(X.minv + k) : ones4();
      float4 p;
      BFX_MAP4(p, sqrtf(comp(mi, c)) * comp(z, c));
      p = mask4(p, m);
      st4(X.mom + k, p);
      acc[0] += p.x * p.x / mi.x + p.y * p.y / mi.y + p.z * p.z / mi.z + p.w * p.w / mi.w;
    });
    block_sum<NT, 1>(acc, X.s.red);
    X.sc.lMH = 0.5f * acc[0];  // -logpdf(moment_dist, pold) up to the constant that cancels
  } else {
    for_each_quad<NT>(R.M, [&](int k) {  // :124,128
      const float4 t0 = ld4(th + k), p = gld4(X.mom + k);
      st4(th + k, make_float4(t0.x + l * p.x, t0.y + l * p.y, t0.z + l * p.z, t0.w + l * p.w));
    });
  }
  __syncthreads();
  float gn2;
  X.grad(gn2);
  const float hk = 0.5f * l * clip_scale(gn2, S.maxnorm);
  const bool last = stage == S.path_len;
  float acc[1] = {0.f};
  for_each_quad<NT>(R.M, [&](int k) {
    const unsigned m = X.s.mask[k >> 2];
    if (!m) return;
    const float4 p0 = gld4(X.mom + k), gg = ld4(g + k);
    float4 p;
    BFX_MAP4(p, comp(p0, c) + hk * comp(gg, c));  // momentum .-= l*(-∇)/2
    p = mask4(p, m);
    if (last) {
      p = make_float4(-p.x, -p.y, -p.z, -p.w);  // :133
      const float4 mi = um ? gld4(X.minv + k) : ones4();
      acc[0] += p.x * p.x / mi.x + p.y * p.y / mi.y + p.z * p.z / mi.z + p.w * p.w / mi.w;
    }
    st4(X.mom + k, p);
  });
  if (last) {
    block_sum<NT, 1>(acc, X.s.red);
    const double start_mh = -X.sc.start_lp + (double)X.sc.lMH;
    mass_adapt(X, gn2);  // (before the full-data pass, which may use s.g as workspace: see step_ggmc)
    const double end_lp = X.full_lp();
    const double end_mh = -end_lp + 0.5 * (double)acc[0];
    const float lMH = (float)(start_mh - end_mh);
    const float lr = logf(X.uniform());
    X.sc.l = stepsize_adapt(X.sc, S, mh_prob_of(lMH));
    const bool accept = lr < lMH;
    const long long ns = X.sc.nsampled + 1;
    if (R.accepted_out && threadIdx.x == 0) {
      const long long hi_ = hist_index(R, X.c, ns);
      if (hi_ >= 0) R.accepted_out[hi_] = accept ? 1 : 0;
    }
    if (accept) X.sc.start_lp = end_lp;  // the next trajectory starts from theta (accept) or theta_old (reject)
    X.sc.lp_epoch = R.data_epoch + 1;
    if (accept) {
      X.sc.accepted_total += 1;
    } else {
      __syncthreads();
      for_each_quad<NT>(R.M, [&](int k) { st4(th + k, gld4(X.thold + k)); });
      __syncthreads();
    }
    X.sc.nsampled = ns;
    record_sample(X, ns);
  }
  X.sc.current_step = last ? 1 : stage + 1;
  X.sc.t += 1;
}

// --------------------------------------------------------------------------------
// find_mode: FluxModeFinder.step!  (src/inference/mode/flux.jl:33-54) with the Flux.Optimise rule
// applied to -g (Flux minimises).  Optimiser state: m in the momentum vector, v in the RMSProp vector,
// prevθ in the θold vector.  A converged chain is frozen (find_mode returns at that point,
// src/inference/mode/abstract.jl:74-78).
// --------------------------------------------------------------------------------
template <int NT>
BFX_DEV float block_max(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if (NT > 32) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    v = red[0];
#pragma unroll
    for (int ww = 1; ww < NT / 32; ++ww) v = fmaxf(v, red[ww]);
  }
  return v;
}

template <int NT, int BT, bool TC, class SP>
BFX_DEV void step_mode(Ctx<NT, BT, TC, SP>& X) {
  const RunDev& R = X.R;
  const SamplerDev& S = R.S;
  if (X.sc.converged) return;
  const int w = S.mode_windowlength;
  if (X.sc.mode_i == 1) {  // flux.jl:36-41
    float mx = 0.f;
    for (int i = threadIdx.x; i < w; i += NT) mx = fmaxf(mx, fabsf(__ldcg(X.window + i)));
    mx = block_max<NT>(mx, X.s.red);
    if (mx <= S.mode_abstol) {
      X.sc.converged = 1;
      return;
    }
  }
  float gn2;
  X.grad(gn2);
  float* th = X.s.th;
  const float* g = X.s.g;
  const float eta = S.opt_eta, rho = S.opt_rho, b2 = S.opt_beta2, eps = S.opt_eps;
  const float c1 = 1.f / (1.f - X.sc.bp1), c2 = 1.f / (1.f - X.sc.bp2);
  float mx = 0.f;
  for_each_quad<NT>(R.M, [&](int k) {
    const unsigned m = X.s.mask[k >> 2];
    if (!m) return;
    const float4 t0 = ld4(th + k), gg = ld4(g + k), pv = gld4(X.thold + k);
    const float4 ch = mask4(make_float4(fabsf(pv.x - t0.x), fabsf(pv.y - t0.y), fabsf(pv.z - t0.z), fabsf(pv.w - t0.w)), m);
    mx = fmaxf(mx, fmaxf(fmaxf(ch.x, ch.y), fmaxf(ch.z, ch.w)));  // maximum(abs, prevθ - θ)   :44
    st4(X.thold + k, t0);                                          // prevθ = copy(θ)            :46
    float4 step;
    if (S.opt_kind == OPT_DESCENT) {
      BFX_MAP4(step, eta * -comp(gg, c));
    } else if (S.opt_kind == OPT_MOMENTUM) {
      const float4 v0 = gld4(X.mom + k);
      float4 v;
      BFX_MAP4(v, rho * comp(v0, c) - eta * -comp(gg, c));
      st4(X.mom + k, mask4(v, m));
      BFX_MAP4(step, -comp(v, c));
    } else if (S.opt_kind == OPT_RMSPROP) {
      const float4 a0 = gld4(X.rmsv + k);
      float4 a;
      BFX_MAP4(a, rho * comp(a0, c) + (1.f - rho) * comp(gg, c) * comp(gg, c));
      st4(X.rmsv + k, mask4(a, m));
      BFX_MAP4(step, -comp(gg, c) * (eta / (sqrtf(comp(a, c)) + eps)));
    } else {  // ADAM
      const float4 m0 = gld4(X.mom + k), v0 = gld4(X.rmsv + k);
      float4 mt, vt;
      BFX_MAP4(mt, rho * comp(m0, c) + (1.f - rho) * -comp(gg, c));
      BFX_MAP4(vt, b2 * comp(v0, c) + (1.f - b2) * comp(gg, c) * comp(gg, c));
      st4(X.mom + k, mask4(mt, m));
      st4(X.rmsv + k, mask4(vt, m));
      BFX_MAP4(step, comp(mt, c) * c1 / (sqrtf(comp(vt, c) * c2) + eps) * eta);
    }
    float4 o;
    BFX_MAP4(o, comp(t0, c) - comp(step, c));                      // Flux.update!(opt, θ, -g)   :51
    st4(th + k, mask4(o, m));
  });
  mx = block_max<NT>(mx, X.s.red);
  if (threadIdx.x == 0) X.window[X.sc.mode_i - 1] = mx;            // :45
  X.sc.mode_i = X.sc.mode_i == w ? 1 : X.sc.mode_i + 1;            // :47
  if (S.opt_kind == OPT_ADAM) {
    X.sc.bp1 *= rho;
    X.sc.bp2 *= b2;
  }
  __syncthreads();
  X.sc.nsampled += 1;
  X.sc.t += 1;
}

}  // namespace bfx
