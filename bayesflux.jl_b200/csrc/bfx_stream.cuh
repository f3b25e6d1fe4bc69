// Large-P ("stream") regime: theta, gradient and sampler state live in HBM in the reference's flat
// order, the network is evaluated layer by layer with tiled FP32 GEMMs over the whole minibatch, and
// every sampler update is one or two fused element-wise passes over theta with grid-wide reductions
// through double atomics.  Used when theta + gradient + one batch tile do not fit the 227 KB of shared
// memory of a CTA (configs[2]: P = 559 106), one chain at a time, optionally with the minibatch sharded
// over ranks (the caller all-reduces the P + 3 floats of StreamState::gbuf between grad and update).
//
// Reference arithmetic: the same rows of SURVEY.md section 8a as the shared-memory regime
// (src/model/BNN.jl:50-69, src/likelihoods/feedforward.jl:30-43,86-97, src/netpriors/*.jl,
// src/inference/mcmc/sgld.jl:96-112, sgnht.jl:64-90, sgnht-s.jl:81-116).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "bfx_common.cuh"

namespace bfx {

// per-layer flat offsets (reference layout: W column-major out x in at w_off, then bias)
struct StreamLayer {
  int nin, nout, act;
  long long w_off, b_off;
  int tc;  // 1: this layer's three GEMMs run on the tensor cores (bfx_stream_tc.cu)
};

struct StreamModel {
  int nlayers;
  long long P, Pnet;
  int d_in;
  int like_kind, sprior_kind;
  float nu, sprior_a, sprior_b, prior_c2;
  double prior_c0n, tdist_const, sprior_const;
  StreamLayer L[BFX_MAX_LAYERS];
  int tc;  // any layer on the tensor cores
};

// device-side scalars of one chain in the stream regime
struct StreamScalars {
  double ds0, ds1, bcount;   // likelihood sums of the current evaluation (local to this rank before the all-reduce)
  double th2;                // sum theta_net^2
  double gn2;                // ||g||^2 over all P
  double acc0, acc1;         // sampler reductions (p' Minv p, NaN flag)
  double value;              // num_batches*loglike + logprior of the last evaluation
  float glike;               // d/d theta_like
  float xi, l;
  float e1, sc;              // SGNHTS O-step coefficients
  float kin;
  int status;
  long long t, nsampled;
  int ma_t;
};

struct StreamWork {
  long long Bc;               // batch-chunk capacity (columns)
  float* act[BFX_MAX_LAYERS + 1];  // act[0] = gathered x (d x Bc); act[l+1] = output of layer l (nout x Bc)
  float* delta[2];            // ping-pong deltas (max width x Bc)
  float* yb;                  // [Bc] targets
  float* dl;                  // [Bc] likelihood deltas (scaled by num_batches)
  float* part;                // split-K partials of the weight-gradient GEMMs
  long long part_floats;
  float* rowpart;             // [64][max width] partial row sums (bias gradients of the tensor-core layers)
  // tensor-core mode: BF16 hi / lo halves of the activations, the deltas and theta
  __nv_bfloat16* act_hi[BFX_MAX_LAYERS + 1];
  __nv_bfloat16* act_lo[BFX_MAX_LAYERS + 1];
  __nv_bfloat16* delta_hi[2];
  __nv_bfloat16* delta_lo[2];
  __nv_bfloat16* w_hi;
  __nv_bfloat16* w_lo;
};

// ---- tcgen05 GEMM (bfx_stream_tc.cu) ---------------------------------------------------------------------
enum { EPI_BIAS_ACT_TC = 0, EPI_MUL_DACT_TC = 1, EPI_PLAIN_TC = 2 };
// one operand: element (mn, k) lives at hi[mn * ld + k] (K-major) or hi[k * ld + mn] (MN-major)
struct TcOperand {
  const __nv_bfloat16* hi;
  const __nv_bfloat16* lo;
  long long ld;
};
struct TcGemmArgs {
  int M, N, K;
  TcOperand A, B;
  float* C;              // FP32 output; EPI_PLAIN: the split-K partial / accumulate target
  __nv_bfloat16* Chi;    // BF16 halves of the output (nullptr: not written)
  __nv_bfloat16* Clo;
  long long ldc;
  const float* bias;
  int act;
  const float* aux;      // EPI_MUL_DACT: multiply by act'(aux[m + ldaux * n])
  long long ldaux;
  int accumulate;        // EPI_PLAIN, single split: C += acc
  long long kchunk;      // split-K: blockIdx.z handles k in [z * kchunk, ...), writes C + z * part_stride
  long long part_stride;
  int debug;             // timing experiments only (env BFX_GEMM_DEBUG): 1 no epilogue stores, 2 no MMA, 4 no loads
};
cudaError_t tc_gemm_forward(const TcGemmArgs& g, cudaStream_t st);          // A MN-major (W), B K-major (activations)
cudaError_t tc_gemm_dx(const TcGemmArgs& g, cudaStream_t st);               // A K-major (W^T), B K-major (delta)
cudaError_t tc_gemm_dw(const TcGemmArgs& g, int ksplit, cudaStream_t st);   // A MN-major (delta), B MN-major (activations)

cudaError_t stream_work_alloc(const StreamModel& M, long long Bc, StreamWork& W);
void stream_work_free(StreamWork& W);

// minibatch description for the stream regime
struct StreamBatch {
  const long long* idx;  // explicit indices or nullptr
  long long n, first, B; // data-set size, first position inside the epoch permutation, batch size
  unsigned k0, k1;
  int shuffle, full;
};

// likelihood part of the evaluation on one minibatch: g[0..Pnet) += num_batches * d loglike / d theta_net
// (g must be zeroed by the caller when accumulate == 0), sc->ds0/ds1/bcount += sums.
cudaError_t stream_eval_like(const StreamModel& M, StreamWork& W, const float* theta, const float* x, const float* y,
                             const StreamBatch& b, float nb, bool want_grad, float* g, StreamScalars* sc,
                             cudaStream_t st, float* yhat_out = nullptr);
// prior + sigma prior + value + ||g||^2 (all on the device; nothing is synchronised)
cudaError_t stream_finish(const StreamModel& M, const float* theta, float* g, StreamScalars* sc, float nb, bool want_grad,
                          cudaStream_t st);
// gbuf tail <-> scalars (minibatch-sharded chains all-reduce [g ; ds0 ; ds1 ; bcount] as P + 3 floats)
cudaError_t stream_pack_sums(float* g, long long P, StreamScalars* sc, bool to_buffer, cudaStream_t st);
cudaError_t stream_zero(float* g, long long P, StreamScalars* sc, cudaStream_t st);

struct StreamUpdate {
  const SamplerDev* S;   // host copy
  long long P;
  float* theta;
  float* mom;
  float* minv;  // nullptr: identity
  float* rmsv;
  float* g;
  StreamScalars* sc;
  unsigned long long seed, gstep;
  unsigned chain;
  const float* inj;      // injected normals (P floats) or nullptr
  float* sample_out;     // where to record theta after the update (device, P floats) or nullptr
  float* kinetic_out;    // device scalar slot or nullptr
};
cudaError_t stream_update(const StreamUpdate& U, cudaStream_t st);

}  // namespace bfx
