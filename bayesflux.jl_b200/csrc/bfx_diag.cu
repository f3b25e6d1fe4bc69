// Chain diagnostics over recorded samples: pooled moments, split-Rhat and effective sample
// size per parameter, computed where the samples already live (HBM).  The reference itself ships no diagnostics: its
// README hands `samples'` to MCMCChains (README.md:191-193).  The arithmetic is the published one -- split-Rhat and
// the Geyer initial-sequence ESS of the Stan reference manual / Vehtari et al. (2021), without rank normalisation --
// restated on the CPU in oracle/bfx_oracle.py (chain_diagnostics).
//
// Layout of `samples`: [C][n][P] floats (= the P x n x C column-major array bfx_mcmc_run returns and the ring of
// bfx_mcmc_advance with ring_slots = n).  Every chain is split in halves (the middle draw of an odd n is dropped),
// giving m = 2C sequences of length nh.  Threads own consecutive parameters, so every access is a coalesced row.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <string>

#include <cuda_runtime.h>

#include "../../include/bfx.h"

namespace bfx {
int32_t report_error(int32_t code, const std::string& msg);  // bfx_api.cu
inline int32_t fail(int32_t code, const std::string& msg) { return report_error(code, msg); }

namespace {

__device__ __forceinline__ long long seq_first(int s, long long n, long long nh) {
  return (s & 1) ? (n - nh) : 0;  // second half starts after the dropped middle draw
}

// mean and variance (ddof 1) of every (sequence, parameter)
__global__ void k_diag_moments(const float* __restrict__ x, long long P, long long n, long long nh, float* __restrict__ mu,
                               float* __restrict__ var) {
  const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int s = blockIdx.y;
  if (j >= P) return;
  const float* row = x + ((long long)(s >> 1) * n + seq_first(s, n, nh)) * P + j;
  double sum = 0.0;
  for (long long t = 0; t < nh; ++t) sum += (double)row[t * P];
  const double m = sum / (double)nh;
  double ss = 0.0;
  for (long long t = 0; t < nh; ++t) {
    const double d = (double)row[t * P] - m;
    ss += d * d;
  }
  mu[(long long)s * P + j] = (float)m;
  var[(long long)s * P + j] = nh > 1 ? (float)(ss / (double)(nh - 1)) : 0.f;
}

// mean over the sequences of the biased autocovariance at lags [k0, k0 + nk): acov[k][j] += sum_t d_t d_{t+k} / (nh m)
__global__ void k_diag_acov(const float* __restrict__ x, long long P, long long n, long long nh, int m,
                            const float* __restrict__ mu, int k0, int nk, double* __restrict__ acov) {
  const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int s = blockIdx.y;
  if (j >= P) return;
  const float* row = x + ((long long)(s >> 1) * n + seq_first(s, n, nh)) * P + j;
  const float mj = mu[(long long)s * P + j];
  const double scale = 1.0 / ((double)nh * (double)m);
  for (int k = k0; k < k0 + nk && k < nh; ++k) {
    float acc = 0.f;
    double tot = 0.0;
    int cnt = 0;
    for (long long t = 0; t + k < nh; ++t) {
      acc += (row[t * P] - mj) * (row[(t + k) * P] - mj);
      if (++cnt == 256) {  // float partial sums folded into a double every 256 terms
        tot += (double)acc;
        acc = 0.f;
        cnt = 0;
      }
    }
    tot += (double)acc;
    atomicAdd(&acov[(long long)k * P + j], tot * scale);
  }
}

// per parameter: pooled moments, split-Rhat, Geyer ESS from the mean autocovariances of lags [0, L)
__global__ void k_diag_finish(long long P, long long nh, int m, int L, const float* __restrict__ mu,
                              const float* __restrict__ var, const double* __restrict__ acov, float* __restrict__ mean_out,
                              float* __restrict__ var_out, float* __restrict__ rhat_out, float* __restrict__ ess_out) {
  const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (j >= P) return;
  double gm = 0.0, W = 0.0;
  for (int s = 0; s < m; ++s) {
    gm += (double)mu[(long long)s * P + j];
    W += (double)var[(long long)s * P + j];
  }
  gm /= (double)m;
  W /= (double)m;
  double bsum = 0.0;
  for (int s = 0; s < m; ++s) {
    const double d = (double)mu[(long long)s * P + j] - gm;
    bsum += d * d;
  }
  const double var_means = m > 1 ? bsum / (double)(m - 1) : 0.0;  // B / nh
  const double var_plus = W * (double)(nh - 1) / (double)nh + var_means;
  if (mean_out) mean_out[j] = (float)gm;
  if (var_out) {  // variance (ddof 1) of all m * nh draws
    const double tot = (double)m * (double)nh;
    var_out[j] = (float)((W * (double)(nh - 1) * (double)m + bsum * (double)nh) / (tot - 1.0));
  }
  if (rhat_out) rhat_out[j] = W > 0.0 ? (float)sqrt(var_plus / W) : NAN;
  if (!ess_out) return;
  const double total = (double)m * (double)nh;
  if (!(var_plus > 0.0) || nh < 4 || L < 2) {
    ess_out[j] = NAN;
    return;
  }
  // Geyer's initial positive sequence on pairs, then the initial monotone sequence (Stan reference manual)
  auto rho = [&](int t) { return 1.0 - (W - acov[(long long)t * P + j]) / var_plus; };
  double tau = 0.0;  // sum of the accepted rho_t (t = 0 counted once: tau_hat = -1 + 2 * sum)
  double prev_pair = 1e300;
  int t = 0;
  double last_even = 1.0;
  while (t + 1 < L) {
    double pe = t == 0 ? 1.0 : rho(t), po = rho(t + 1);
    double pair = pe + po;
    if (!(pair > 0.0)) {
      last_even = pe;
      break;
    }
    if (pair > prev_pair) {  // monotone: a pair never exceeds the one before it
      pe = po = 0.5 * prev_pair;
      pair = prev_pair;
    }
    tau += pe + po;
    prev_pair = pair;
    t += 2;
    last_even = 0.0;
  }
  if (last_even > 0.0) tau += 0.5 * last_even;  // the improved estimate adds the positive even term once
  double tau_hat = -1.0 + 2.0 * tau;
  const double floor_tau = 1.0 / log10(total);
  if (tau_hat < floor_tau) tau_hat = floor_tau;
  ess_out[j] = (float)(total / tau_hat);
}

}  // namespace
}  // namespace bfx

extern "C" int32_t bfx_chain_diagnostics(int32_t device, const float* samples, int32_t samples_on_device, int64_t P,
                                         int64_t n, int64_t C, int64_t max_lag, float* mean, float* var, float* rhat,
                                         float* ess) {
  using namespace bfx;
  if (!samples) return fail(BFX_ERR_BAD_ARG, "samples is NULL");
  if (P <= 0 || C <= 0 || n < 4) return fail(BFX_ERR_BAD_ARG, "diagnostics need P > 0, C > 0 and at least 4 draws per chain");
  if (2 * C > 65535) return fail(BFX_ERR_UNSUPPORTED, "diagnostics: at most 32767 chains");
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(BFX_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
  const long long nh = n / 2;
  const int m = (int)(2 * C);
  long long L = nh;  // lags 0 .. nh - 1
  if (max_lag > 0 && max_lag + 1 < L) L = max_lag + 1;
  float *d_x = nullptr, *d_mu = nullptr, *d_var = nullptr, *d_out = nullptr;
  double* d_acov = nullptr;
  auto cleanup = [&]() {
    if (!samples_on_device) cudaFree(d_x);
    cudaFree(d_mu);
    cudaFree(d_var);
    cudaFree(d_out);
    cudaFree(d_acov);
  };
#define DCK(call)                                                                         \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      cleanup();                                                                          \
      return fail(BFX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));     \
    }                                                                                     \
  } while (0)
  const size_t nx = (size_t)C * n * P;
  if (samples_on_device) {
    d_x = const_cast<float*>(samples);
  } else {
    DCK(cudaMalloc(&d_x, nx * 4));
    DCK(cudaMemcpy(d_x, samples, nx * 4, cudaMemcpyHostToDevice));
  }
  DCK(cudaMalloc(&d_mu, (size_t)m * P * 4));
  DCK(cudaMalloc(&d_var, (size_t)m * P * 4));
  DCK(cudaMalloc(&d_out, (size_t)4 * P * 4));
  const dim3 grid((unsigned)((P + 255) / 256), (unsigned)m);
  k_diag_moments<<<grid, 256>>>(d_x, P, n, nh, d_mu, d_var);
  DCK(cudaGetLastError());
  if (ess) {
    DCK(cudaMalloc(&d_acov, (size_t)L * P * sizeof(double)));
    DCK(cudaMemset(d_acov, 0, (size_t)L * P * sizeof(double)));
    for (long long k0 = 0; k0 < L; k0 += 64) {
      k_diag_acov<<<grid, 256>>>(d_x, P, n, nh, m, d_mu, (int)k0, (int)std::min<long long>(64, L - k0), d_acov);
      DCK(cudaGetLastError());
    }
  }
  k_diag_finish<<<(unsigned)((P + 255) / 256), 256>>>(P, nh, m, (int)L, d_mu, d_var, d_acov, mean ? d_out : nullptr,
                                                       var ? d_out + P : nullptr, rhat ? d_out + 2 * P : nullptr,
                                                       ess ? d_out + 3 * P : nullptr);
  DCK(cudaGetLastError());
  DCK(cudaDeviceSynchronize());
  if (mean) DCK(cudaMemcpy(mean, d_out, (size_t)P * 4, cudaMemcpyDeviceToHost));
  if (var) DCK(cudaMemcpy(var, d_out + P, (size_t)P * 4, cudaMemcpyDeviceToHost));
  if (rhat) DCK(cudaMemcpy(rhat, d_out + 2 * P, (size_t)P * 4, cudaMemcpyDeviceToHost));
  if (ess) DCK(cudaMemcpy(ess, d_out + 3 * P, (size_t)P * 4, cudaMemcpyDeviceToHost));
#undef DCK
  cleanup();
  return BFX_OK;
}
