// gemm_probe: bring-up and timing of the stream-regime tcgen05 GEMM (bayesflux.jl_b200/csrc/bfx_tc_gemm.cuh) outside
// the library: every operand majorness (forward / input gradient / weight gradient), both CTA-group modes, ragged
// shapes, split-K, the three epilogues; results against a double-precision host product of the FP32 inputs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DBFX_WATCHDOG -Ibayesflux.jl_b200/csrc \
//        -o tools/gemm_probe.exe tools/gemm_probe.cu
// Run:   tools/gemm_probe.exe [cg=1|2|0(both)] [time=0|1]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "bfx_tc_gemm.cuh"

using namespace bfx;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); fflush(stdout); exit(2);} } while (0)

static void split_host(const std::vector<float>& v, std::vector<__nv_bfloat16>& hi, std::vector<__nv_bfloat16>& lo) {
  hi.resize(v.size()); lo.resize(v.size());
  for (size_t i = 0; i < v.size(); ++i) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v[i]);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(v[i] - __bfloat162float(h));
  }
}
template <class T>
static T* to_dev(const std::vector<T>& v) {
  T* p = nullptr;
  CK(cudaMalloc(&p, v.size() * sizeof(T) + 256));
  CK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return p;
}

struct Case { const char* name; int mode; int M, N, K, ksplit; int epi; int act; };  // mode 0 fwd (A MN, B K), 1 dx (K, K), 2 dw (MN, MN)

template <int CG>
static cudaError_t run_mode(int mode, int epi, const tcg::Args& g, const tcg::Operand& A, const tcg::Operand& B,
                            const tcg::Output& O, int ksplit, cudaStream_t st) {
  if (mode == 0) return epi == tcg::EPI_BIAS_ACT ? tcg::launch<CG, true, false, tcg::EPI_BIAS_ACT>(g, A, B, O, ksplit, st)
                                                  : tcg::launch<CG, true, false, tcg::EPI_PLAIN>(g, A, B, O, ksplit, st);
  if (mode == 1) return epi == tcg::EPI_MUL_DACT ? tcg::launch<CG, false, false, tcg::EPI_MUL_DACT>(g, A, B, O, ksplit, st)
                                                  : tcg::launch<CG, false, false, tcg::EPI_PLAIN>(g, A, B, O, ksplit, st);
  return tcg::launch<CG, true, true, tcg::EPI_PLAIN>(g, A, B, O, ksplit, st);
}

template <int CG>
static int run_case(const Case& c, bool timeit) {
  const int M = c.M, N = c.N, K = c.K;
  std::mt19937 rng(1234 + M + 7 * N + 13 * K + c.mode);
  std::normal_distribution<float> nd(0.f, 1.f);
  const bool a_mn = c.mode != 1, b_mn = c.mode == 2;
  const long long lda = a_mn ? M : K, ldb = b_mn ? N : K;
  std::vector<float> A((size_t)M * K), B((size_t)N * K), bias(M), aux((size_t)M * N);
  for (auto& v : A) v = nd(rng);
  for (auto& v : B) v = nd(rng);
  for (auto& v : bias) v = nd(rng);
  for (auto& v : aux) v = nd(rng);
  auto a_at = [&](int m, int k) { return a_mn ? A[(size_t)k * lda + m] : A[(size_t)m * lda + k]; };
  auto b_at = [&](int n, int k) { return b_mn ? B[(size_t)k * ldb + n] : B[(size_t)n * ldb + k]; };
  std::vector<__nv_bfloat16> Ah, Al, Bh, Bl, Xh, Xl;
  split_host(A, Ah, Al); split_host(B, Bh, Bl); split_host(aux, Xh, Xl);
  __nv_bfloat16 *dAh = to_dev(Ah), *dAl = to_dev(Al), *dBh = to_dev(Bh), *dBl = to_dev(Bl), *dXh = to_dev(Xh), *dXl = to_dev(Xl);
  float* dbias = to_dev(bias);
  const int ksplit = c.ksplit;
  int kchunk = (K + ksplit - 1) / ksplit;
  kchunk = (kchunk + tcg::BK - 1) / tcg::BK * tcg::BK;
  float* dC = nullptr;
  CK(cudaMalloc(&dC, (size_t)M * N * 4 * ksplit));
  CK(cudaMemset(dC, 0xFF, (size_t)M * N * 4 * ksplit));
  __nv_bfloat16 *dChi = nullptr, *dClo = nullptr;
  CK(cudaMalloc(&dChi, (size_t)M * N * 2)); CK(cudaMalloc(&dClo, (size_t)M * N * 2));
  const int nslots = tcg::ROWSLOTS * ((N + 223) / 224);  // tiles are 224 or 256 columns wide (tcg::tile_cols_kmajor_b); unused slots stay 0
  float* drow = nullptr;
  CK(cudaMalloc(&drow, (size_t)nslots * M * 4));
  CK(cudaMemset(drow, 0, (size_t)nslots * M * 4));
  tcg::Args g{};
  g.M = M; g.N = N; g.K = K;
  g.C = dC; g.ldc = M; g.kchunk = kchunk; g.part_stride = (long long)M * N;
  g.bias = dbias; g.act = c.act;
  tcg::Output oc{nullptr, nullptr};
  if (c.epi != tcg::EPI_PLAIN) { oc.hi = dChi; oc.lo = dClo; }
  const int nwords = (N + 31) / 32;
  std::vector<uint32_t> maskh((size_t)nwords * M, 0u);
  for (int n = 0; n < N; ++n)
    for (int m = 0; m < M; ++m)
      if (__bfloat162float(Xh[(size_t)n * M + m]) > 0.f) maskh[(size_t)(n >> 5) * M + m] |= 1u << (n & 31);
  uint32_t* dmask_in = to_dev(maskh);
  uint32_t* dmask_out = nullptr;
  CK(cudaMalloc(&dmask_out, maskh.size() * 4));
  CK(cudaMemset(dmask_out, 0, maskh.size() * 4));
  if (c.epi == tcg::EPI_MUL_DACT) {
    g.aux_hi = dXh; g.aux_lo = dXl; g.ldaux = M; g.rowpart = drow;
    if (c.act == A_RELU) g.mask_in = dmask_in;
  }
  if (c.epi == tcg::EPI_BIAS_ACT && c.act == A_RELU) g.mask_out = dmask_out;
  long long* dtim = nullptr;
  const int nctas = ((M + 128 * CG - 1) / (128 * CG)) * CG * ((N + tcg::BN - 1) / tcg::BN) * ksplit;
  CK(cudaMalloc(&dtim, (size_t)(nctas + 2) * 8 * sizeof(long long)));
  CK(cudaMemset(dtim, 0, (size_t)(nctas + 2) * 8 * sizeof(long long)));
  g.timing = dtim;
  tcg::Operand oa{dAh, dAl, lda}, ob{dBh, dBl, ldb};
  cudaError_t e = run_mode<CG>(c.mode, c.epi, g, oa, ob, oc, ksplit, 0);
  if (e != cudaSuccess) { printf("%-28s cg%d launch error %s\n", c.name, CG, cudaGetErrorString(e)); fflush(stdout); return 1; }
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-28s cg%d run error %s\n", c.name, CG, cudaGetErrorString(e)); fflush(stdout); exit(3); }
  std::vector<float> C((size_t)M * N * ksplit);
  CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
  std::vector<__nv_bfloat16> Chi((size_t)M * N), Clo((size_t)M * N);
  CK(cudaMemcpy(Chi.data(), dChi, Chi.size() * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(Clo.data(), dClo, Clo.size() * 2, cudaMemcpyDeviceToHost));
  std::vector<float> rowp((size_t)nslots * M);
  CK(cudaMemcpy(rowp.data(), drow, rowp.size() * 4, cudaMemcpyDeviceToHost));
  // reference on a sample of entries (all entries for small shapes)
  const long long total = (long long)M * N;
  const long long nsamp = total <= 200000 ? total : 20000;
  double num = 0, den = 0, maxabs = 0, num2 = 0;
  std::uniform_int_distribution<long long> pick(0, total - 1);
  for (long long s = 0; s < nsamp; ++s) {
    const long long e2 = total <= 200000 ? s : pick(rng);
    const int m = (int)(e2 % M), n = (int)(e2 / M);
    double ref = 0;
    for (int k = 0; k < K; ++k) ref += (double)a_at(m, k) * (double)b_at(n, k);
    double got = 0;
    for (int z = 0; z < ksplit; ++z) got += (double)C[(size_t)z * M * N + (size_t)n * M + m];
    if (c.epi == tcg::EPI_BIAS_ACT) ref = c.act == A_RELU ? fmax(ref + bias[m], 0.0) : tanh(ref + bias[m]);
    if (c.epi == tcg::EPI_MUL_DACT) {
      const double a = (double)__bfloat162float(Xh[e2]) + (double)__bfloat162float(Xl[e2]);
      ref *= c.act == A_RELU ? (__bfloat162float(Xh[e2]) > 0.f ? 1.0 : 0.0) : 1.0 - a * a;
    }
    num += (got - ref) * (got - ref); den += ref * ref;
    maxabs = fmax(maxabs, fabs(got - ref));
    if (c.epi != tcg::EPI_PLAIN) {
      const double hl = (double)__bfloat162float(Chi[e2]) + (double)__bfloat162float(Clo[e2]);
      num2 += (hl - ref) * (hl - ref);
    }
  }
  double rowerr = 0;
  if (c.epi == tcg::EPI_MUL_DACT && total <= 200000) {
    for (int m = 0; m < M; ++m) {
      double ref = 0, got = 0;
      for (int n = 0; n < N; ++n) { double z = 0; for (int z2 = 0; z2 < ksplit; ++z2) z += C[(size_t)z2 * M * N + (size_t)n * M + m]; ref += z; }
      for (int sl = 0; sl < nslots; ++sl) got += rowp[(size_t)sl * M + m];
      rowerr = fmax(rowerr, fabs(ref - got) / (fabs(ref) + 1.0));
    }
  }
  long long maskbad = 0;
  if (g.mask_out) {
    std::vector<uint32_t> mo(maskh.size());
    CK(cudaMemcpy(mo.data(), dmask_out, mo.size() * 4, cudaMemcpyDeviceToHost));
    for (int n = 0; n < N; ++n)
      for (int m = 0; m < M; ++m) {
        const bool bit = (mo[(size_t)(n >> 5) * M + m] >> (n & 31)) & 1u;
        if (bit != (C[(size_t)n * M + m] > 0.f)) ++maskbad;
      }
  }
  const double rel = sqrt(num / (den + 1e-300)), rel2 = sqrt(num2 / (den + 1e-300));
  const bool ok = rel < 3e-5 && rel2 < 3e-5 && rowerr < 1e-4 && rel == rel && maskbad == 0;
  printf("%-28s cg%d M=%d N=%d K=%d ks=%d epi=%d  rel=%.3e maxabs=%.3e hilo_rel=%.3e rowsum=%.2e maskbad=%lld  %s\n", c.name,
         CG, M, N, K, ksplit, c.epi, rel, maxabs, rel2, rowerr, maskbad, ok ? "OK" : "MISMATCH");
  fflush(stdout);
  if (timeit) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    if (c.epi != tcg::EPI_PLAIN) g.C = nullptr;  // production: only the BF16 halves leave the kernel
    for (int i = 0; i < 3; ++i) run_mode<CG>(c.mode, c.epi, g, oa, ob, oc, ksplit, 0);
    // the timed launches replay from a CUDA graph (as the library does): no host launch cost in the number
    const int iters = 20;
    cudaStream_t cs;
    CK(cudaStreamCreate(&cs));
    cudaGraph_t graph;
    cudaGraphExec_t gexec;
    CK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed));
    for (int i = 0; i < iters; ++i) CK(run_mode<CG>(c.mode, c.epi, g, oa, ob, oc, ksplit, cs));
    CK(cudaStreamEndCapture(cs, &graph));
    CK(cudaGraphInstantiate(&gexec, graph, 0));
    CK(cudaGraphLaunch(gexec, cs));
    CK(cudaStreamSynchronize(cs));
    CK(cudaEventRecord(e0, cs));
    CK(cudaGraphLaunch(gexec, cs));
    CK(cudaEventRecord(e1, cs));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double us = 1e3 * ms / iters, fl = 2.0 * M * N * (double)K;
    printf("    time %.2f us  %.1f TFLOP/s FP32-equivalent, %.1f TFLOP/s BF16 issued (%.1f%% of 2250)\n", us, fl / us * 1e-6,
           3 * fl / us * 1e-6, 100.0 * 3 * fl / us * 1e-6 / 2250.0);
#ifdef BFX_GEMM_TIMING
    {
      std::vector<long long> tm((size_t)nctas * 8);
      CK(cudaMemcpy(tm.data(), dtim, tm.size() * sizeof(long long), cudaMemcpyDeviceToHost));
      // marks: 0 start, 7 setup done, 1 first stage landed, 2 last MMA issued, 3 accumulator ready, 4 epilogue math done,
      // 5 stores done, 6 exit (1, 2 exist on leader CTAs only)
      double d[7] = {0, 0, 0, 0, 0, 0, 0};
      int nl = 0, na = 0;
      for (int b = 0; b < nctas; ++b) {
        const long long* t = &tm[(size_t)b * 8];
        if (!t[0]) continue;
        ++na;
        d[0] += (double)(t[7] - t[0]); d[3] += (double)(t[3] - t[7]); d[4] += (double)(t[4] - t[3]);
        d[5] += (double)(t[5] - t[4]); d[6] += (double)(t[6] - t[5]);
        if (t[1]) { ++nl; d[1] += (double)(t[1] - t[7]); d[2] += (double)(t[2] - t[1]); }
      }
      printf("    cycles/CTA: setup %.0f | first stage %.0f, issue loop %.0f | setup->acc ready %.0f | epilogue math %.0f | stores %.0f | teardown %.0f  (%d CTAs)\n",
             d[0] / na, d[1] / (nl ? nl : 1), d[2] / (nl ? nl : 1), d[3] / na, d[4] / na, d[5] / na, d[6] / na, na);
      fflush(stdout);
    }
#endif
  }
  cudaFree(dmask_in); cudaFree(dmask_out); cudaFree(dtim);
  cudaFree(dAh); cudaFree(dAl); cudaFree(dBh); cudaFree(dBl); cudaFree(dXh); cudaFree(dXl); cudaFree(dbias);
  cudaFree(dC); cudaFree(dChi); cudaFree(dClo); cudaFree(drow);
  return ok ? 0 : 1;
}

int main(int argc, char** argv) {
  const int cg = argc > 1 ? atoi(argv[1]) : 0;
  const bool timeit = argc > 2 ? atoi(argv[2]) != 0 : true;
  const Case cases[] = {
      {"fwd small exact tile", 0, 256, 256, 64, 1, tcg::EPI_PLAIN, A_IDENTITY},
      {"dx small exact tile", 1, 256, 256, 64, 1, tcg::EPI_PLAIN, A_IDENTITY},
      {"dw small exact tile", 2, 256, 256, 64, 1, tcg::EPI_PLAIN, A_IDENTITY},
      {"fwd K=512", 0, 256, 256, 512, 1, tcg::EPI_PLAIN, A_IDENTITY},
      {"dw K=512", 2, 256, 256, 512, 1, tcg::EPI_PLAIN, A_IDENTITY},
      {"fwd ragged", 0, 320, 200, 72, 1, tcg::EPI_PLAIN, A_IDENTITY},
      {"dx ragged", 1, 200, 328, 136, 1, tcg::EPI_PLAIN, A_IDENTITY},
      {"dw ragged split", 2, 320, 72, 1000, 4, tcg::EPI_PLAIN, A_IDENTITY},
      {"fwd bias+relu ragged", 0, 320, 200, 72, 1, tcg::EPI_BIAS_ACT, A_RELU},
      {"fwd bias+tanh ragged", 0, 328, 520, 72, 1, tcg::EPI_BIAS_ACT, A_TANH},
      {"dx tanh'+rowsum ragged", 1, 200, 328, 136, 1, tcg::EPI_MUL_DACT, A_TANH},
      {"dx relu mask ragged", 1, 200, 328, 136, 1, tcg::EPI_MUL_DACT, A_RELU},
      {"cfg3 fwd L1", 0, 512, 8192, 512, 1, tcg::EPI_BIAS_ACT, A_RELU},
      {"cfg3 fwd L0", 0, 512, 8192, 64, 1, tcg::EPI_BIAS_ACT, A_RELU},
      {"cfg3 dx", 1, 512, 8192, 512, 1, tcg::EPI_MUL_DACT, A_RELU},
      {"cfg3 dx tanh", 1, 512, 8192, 512, 1, tcg::EPI_MUL_DACT, A_TANH},
      {"cfg3 dw ks16", 2, 512, 512, 8192, 16, tcg::EPI_PLAIN, A_IDENTITY},
      {"cfg3 dw L0 ks32", 2, 512, 64, 8192, 32, tcg::EPI_PLAIN, A_IDENTITY},
  };
  int bad = 0;
  for (const Case& c : cases) {
    const bool big = c.N >= 8192 || c.K >= 8192;
    if (cg == 0 || cg == 1) bad += run_case<1>(c, timeit && big);
    if (cg == 0 || cg == 2) bad += run_case<2>(c, timeit && big);
  }
  printf("gemm_probe: %d mismatching cases\n", bad);
  return bad ? 1 : 0;
}
