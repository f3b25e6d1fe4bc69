/*
 * bfx_oracle_c.c -- plain-C restatement of the reference hot path for Dense chains
 * (TEST INFRASTRUCTURE / CPU BASELINE ONLY; never linked into the product).
 *
 * Covers what the headline configurations need on the host cores:
 *   - flat-theta layout of Dense chains        src/layers/dense.jl:13-26, src/model/deconstruct.jl:45-58
 *   - grad of num_batches*loglike + logprior    src/model/BNN.jl:50-69
 *       FeedforwardNormal / FeedforwardTDist    src/likelihoods/feedforward.jl:30-43,86-97
 *       GaussianPrior / MixtureScalePrior       src/netpriors/gaussian.jl:25-28, mixturescale.jl:30-33
 *   - clip_gradient!                             src/utils/gradient_utils.jl:6-10
 *   - SGLD update! and the mcmc batch loop       src/inference/mcmc/sgld.jl:96-112, abstract.jl:102-124
 * It is validated against oracle/bfx_oracle.py (tests/test_oracle_c.py), which in turn is
 * pinned to the reference's known-answer tests.  Parity status: see bfx_oracle.py header.
 *
 * Like the reference it evaluates one chain at a time in Float32 with Float64 reductions;
 * independent chains are spread over the host cores with OpenMP (the reference has no
 * multi-chain API: this is the "embarrassingly parallel" extrapolation made real).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXL 8

typedef struct {
  int nlayers;
  int dims[MAXL + 1]; /* dims[0] = d, dims[l+1] = out of layer l */
  int acts[MAXL];     /* 0 identity, 1 relu, 2 tanh, 3 sigmoid */
  int like_kind;      /* 0 normal, 1 tdist */
  double nu;
  int sprior_kind; /* 0 gamma(shape, scale), 1 invgamma(shape, scale) */
  double sprior_a, sprior_b;
  double c0, c2; /* logprior = Pnet*c0 - 0.5*c2*sum(theta^2) */
} bfxo_model;

static int nparams_net(const bfxo_model* m) {
  int p = 0;
  for (int l = 0; l < m->nlayers; ++l) p += m->dims[l] * m->dims[l + 1] + m->dims[l + 1];
  return p;
}

int bfxo_num_params(const bfxo_model* m) { return nparams_net(m) + 1; }

static inline float actf(int a, float z) {
  switch (a) {
    case 1: return z > 0.f ? z : 0.f;
    case 2: return tanhf(z);
    case 3: return 1.f / (1.f + expf(-z));
    default: return z;
  }
}
static inline float dact(int a, float o) {
  switch (a) {
    case 1: return o > 0.f ? 1.f : 0.f;
    case 2: return 1.f - o * o;
    case 3: return o * (1.f - o);
    default: return 1.f;
  }
}

/* scratch: activations of every layer for a batch of B: sum_l dims[l]*B floats, twice */
static size_t scratch_floats(const bfxo_model* m, int B) {
  size_t s = 0;
  for (int l = 0; l <= m->nlayers; ++l) s += (size_t)m->dims[l] * B;
  return 2 * s;
}

/* value and gradient on the minibatch idx[0..B) of (x: d x n column-major, y).  g may be NULL */
static double grad_one(const bfxo_model* m, const float* theta, const float* x, const float* y, const int64_t* idx,
                       int B, float nb, float* g, float* scratch) {
  const int L = m->nlayers;
  const int Pnet = nparams_net(m);
  float* act[MAXL + 1];
  float* del[MAXL + 1];
  float* p = scratch;
  for (int l = 0; l <= L; ++l) { act[l] = p; p += (size_t)m->dims[l] * B; }
  for (int l = 0; l <= L; ++l) { del[l] = p; p += (size_t)m->dims[l] * B; }
  const int d = m->dims[0];
  for (int b = 0; b < B; ++b) memcpy(act[0] + (size_t)b * d, x + (size_t)idx[b] * d, sizeof(float) * d);
  /* forward: a_{l+1}[o,b] = act(sum_i W[o,i] a_l[i,b] + bias[o]),  W[o,i] = theta[off + o + out*i] */
  int off = 0;
  for (int l = 0; l < L; ++l) {
    const int in = m->dims[l], out = m->dims[l + 1];
    const float* W = theta + off;
    const float* bias = W + in * out;
    for (int b = 0; b < B; ++b) {
      float* o_ = act[l + 1] + (size_t)b * out;
      const float* a_ = act[l] + (size_t)b * in;
      for (int o = 0; o < out; ++o) o_[o] = bias[o];
      for (int i = 0; i < in; ++i) {
        const float ai = a_[i];
        const float* w = W + (size_t)i * out;
        for (int o = 0; o < out; ++o) o_[o] += w[o] * ai;
      }
      for (int o = 0; o < out; ++o) o_[o] = actf(m->acts[l], o_[o]);
    }
    off += in * out + out;
  }
  const float s = theta[Pnet];
  const float sigma = expf(s);
  double s0 = 0.0, s1 = 0.0;
  for (int b = 0; b < B; ++b) {
    const float yh = act[L][b];
    const float z = (y[idx[b]] - yh) / sigma;
    float dl;
    if (m->like_kind == 0) {
      s0 += (double)z * z;
      dl = nb * z / sigma;
    } else {
      const float nu = (float)m->nu;
      const float w = (nu + 1.f) * z / (nu + z * z);
      s0 += log1p((double)z * z / m->nu);
      s1 += (double)w * z;
      dl = nb * w / sigma;
    }
    del[L][b] = dl * dact(m->acts[L - 1], yh);
  }
  double lps, dlps;
  if (m->sprior_kind == 0) {
    lps = -lgamma(m->sprior_a) - m->sprior_a * log(m->sprior_b) + (m->sprior_a - 1) * s - sigma / m->sprior_b + s;
    dlps = m->sprior_a - sigma / m->sprior_b;
  } else {
    lps = m->sprior_a * log(m->sprior_b) - lgamma(m->sprior_a) - (m->sprior_a + 1) * s - m->sprior_b / sigma + s;
    dlps = -m->sprior_a + m->sprior_b / sigma;
  }
  double like, dls;
  if (m->like_kind == 0) {
    like = -0.5 * B * 1.8378770664093453 - 0.5 * s0;
    dls = s0;
  } else {
    like = B * (lgamma(0.5 * (m->nu + 1)) - lgamma(0.5 * m->nu) - 0.5 * log(m->nu * M_PI)) - 0.5 * (m->nu + 1) * s0;
    dls = s1;
  }
  like += -(double)B * s + lps;
  dls += -(double)B + dlps;
  double sq = 0.0;
  for (int j = 0; j < Pnet; ++j) sq += (double)theta[j] * theta[j];
  const double v = (double)nb * like + Pnet * m->c0 - 0.5 * m->c2 * sq;
  if (!g) return v;
  /* backward */
  int offs[MAXL];
  off = 0;
  for (int l = 0; l < L; ++l) { offs[l] = off; off += m->dims[l] * m->dims[l + 1] + m->dims[l + 1]; }
  memset(g, 0, sizeof(float) * (Pnet + 1));
  for (int l = L - 1; l >= 0; --l) {
    const int in = m->dims[l], out = m->dims[l + 1];
    const float* W = theta + offs[l];
    float* gW = g + offs[l];
    float* gb = gW + in * out;
    for (int b = 0; b < B; ++b) {
      const float* dz = del[l + 1] + (size_t)b * out;
      const float* a_ = act[l] + (size_t)b * in;
      float* da = del[l] + (size_t)b * in;
      for (int o = 0; o < out; ++o) gb[o] += dz[o];
      for (int i = 0; i < in; ++i) {
        const float ai = a_[i];
        float* gw = gW + (size_t)i * out;
        const float* w = W + (size_t)i * out;
        float acc = 0.f;
        for (int o = 0; o < out; ++o) {
          gw[o] += dz[o] * ai;
          acc += w[o] * dz[o];
        }
        if (l > 0) da[i] = acc * dact(m->acts[l - 1], ai);
      }
    }
  }
  for (int j = 0; j < Pnet; ++j) g[j] -= (float)m->c2 * theta[j];
  g[Pnet] = (float)((double)nb * dls);
  return v;
}

/* single evaluation (validation entry) */
double bfxo_grad(const bfxo_model* m, const float* theta, const float* x, const float* y, const int64_t* idx, int B,
                 float nb, float* g) {
  float* scratch = (float*)malloc(sizeof(float) * scratch_floats(m, B));
  double v = grad_one(m, theta, x, y, idx, B, nb, g, scratch);
  free(scratch);
  return v;
}

/* xoshiro128+ -> N(0,1) by Box-Muller: host noise for timing runs (the reference uses Julia's randn) */
typedef struct { uint32_t s[4]; } rng_t;
static inline uint32_t rotl(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }
static inline uint32_t rng_next(rng_t* r) {
  uint32_t res = r->s[0] + r->s[3], t = r->s[1] << 9;
  r->s[2] ^= r->s[0]; r->s[3] ^= r->s[1]; r->s[1] ^= r->s[2]; r->s[0] ^= r->s[3];
  r->s[2] ^= t; r->s[3] = rotl(r->s[3], 11);
  return res;
}
static inline void rng_seed(rng_t* r, uint64_t seed) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull;
  for (int i = 0; i < 4; ++i) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 27; z *= 0x94D049BB133111EBull; z ^= z >> 31;
    r->s[i] = (uint32_t)(z >> 16) | 1u;
    z += 0x9E3779B97F4A7C15ull;
  }
}
static inline float rng_unif(rng_t* r) { return ((rng_next(r) >> 8) + 0.5f) * (1.0f / 16777216.0f); }

/*
 * mcmc(bnn, B, nsteps, SGLD(a,b,gamma,maxnorm)) for C independent chains, OpenMP over chains.
 * theta: [C][P] in/out.  normals: NULL (internal RNG) or [nsteps][C][P] injected draws.
 * perm: NULL (internal Fisher-Yates per epoch) or [nsteps][C][B] explicit batch indices.
 * samples_out: NULL or [C][nsteps][P].  Returns the number of grad evaluations done.
 */
int64_t bfxo_sgld_run(const bfxo_model* m, const float* x, const float* y, int64_t n, int C, int B, int nsteps,
                      float a, float b, float gamma, float maxnorm, uint64_t seed, float* theta,
                      const float* normals, const int64_t* batches, float* samples_out, int nthreads) {
  const int P = bfxo_num_params(m);
  const int64_t nbatch = (n + B - 1) / B;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
    float* scratch = (float*)malloc(sizeof(float) * scratch_floats(m, B));
    float* g = (float*)malloc(sizeof(float) * P);
    int64_t* perm = (int64_t*)malloc(sizeof(int64_t) * n);
#pragma omp for schedule(dynamic, 1)
    for (int c = 0; c < C; ++c) {
      rng_t r;
      rng_seed(&r, seed * 1315423911ull + (uint64_t)c);
      float* th = theta + (size_t)c * P;
      for (int t = 0; t < nsteps; ++t) {
        const int64_t k = t % nbatch;
        const int64_t* idx;
        int Bk = B;
        if (batches) {
          idx = batches + ((size_t)t * C + c) * B;
        } else {
          if (k == 0) { /* new epoch: randperm(n) */
            for (int64_t i = 0; i < n; ++i) perm[i] = i;
            for (int64_t i = n - 1; i > 0; --i) {
              int64_t j = (int64_t)(rng_unif(&r) * (float)(i + 1));
              if (j > i) j = i;
              int64_t tmp = perm[i]; perm[i] = perm[j]; perm[j] = tmp;
            }
          }
          idx = perm + k * B;
          if ((k + 1) * B > n) Bk = (int)(n - k * B);
        }
        const float alpha = a * powf(b + (float)(t + 1), -gamma);
        grad_one(m, th, x, y, idx, Bk, (float)nbatch, g, scratch);
        double ng2 = 0.0;
        for (int j = 0; j < P; ++j) ng2 += (double)g[j] * g[j];
        const float ng = (float)sqrt(ng2);
        const float sc = ng > maxnorm ? maxnorm / ng : 1.f;
        const float ha = 0.5f * alpha * sc, sa = sqrtf(alpha);
        if (normals) {
          const float* z = normals + ((size_t)t * C + c) * P;
          for (int j = 0; j < P; ++j) th[j] += ha * g[j] + sa * z[j];
        } else {
          for (int j = 0; j < P; j += 2) {
            float u1 = rng_unif(&r), u2 = rng_unif(&r);
            float rr = sqrtf(-2.f * logf(u1));
            th[j] += ha * g[j] + sa * rr * cosf(6.2831853f * u2);
            if (j + 1 < P) th[j + 1] += ha * g[j + 1] + sa * rr * sinf(6.2831853f * u2);
          }
        }
        if (samples_out) memcpy(samples_out + ((size_t)c * nsteps + t) * P, th, sizeof(float) * P);
      }
    }
    free(scratch);
    free(g);
    free(perm);
  }
  return (int64_t)C * nsteps;
}

int bfxo_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
