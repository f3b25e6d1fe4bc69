// Large-P ("stream") regime: theta, gradient and sampler state live in HBM in the reference's flat
// order, the network is evaluated layer by layer with GEMMs over the whole minibatch (tcgen05 for the wide
// layers, bfx_tc_gemm.cuh; a tiled FP32 kernel for the rest), and every sampler update is one or two fused
// element-wise passes over theta with deterministic grid-wide reductions.  Used when theta + gradient + one
// batch tile do not fit the 227 KB of shared memory of a CTA (configs[2]: P = 559 106), one chain at a time,
// optionally with the minibatch sharded over ranks (the P + 3 floats of the gradient buffer are all-reduced
// between the likelihood part of the evaluation and the update).
//
// Everything an update needs beyond its static description is derived ON THE DEVICE from the chain's scalars
// (global step -> epoch, batch position, permutation keys, Philox counter, sample slot), so that one update is a
// fixed launch sequence: it is captured once in a CUDA graph and replayed (bfx_api.cu: stream_run).
//
// Reference arithmetic: the same rows of SURVEY.md section 8a as the shared-memory regime
// (src/model/BNN.jl:50-69, src/likelihoods/feedforward.jl:30-43,86-97, src/netpriors/*.jl,
// src/inference/mcmc/sgld.jl:96-112, sgnht.jl:64-90, sgnht-s.jl:81-116, abstract.jl:102-124).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "bfx_common.cuh"
#include "bfx_tc_gemm_args.cuh"

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

constexpr int STREAM_RED_BLOCKS = 592;  // grid of every element-wise pass (4 x 148); also the partial-sum slots

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
  long long gstep;           // update! calls done (device mirror of bfx_sampler::gstep; graph replays count here)
  unsigned ticket;           // arrival counter of the grid-wide reductions (wraps to 0 by itself)
  unsigned ticket2;          // same, head kernel
  double part[STREAM_RED_BLOCKS][2];  // per-block partial sums, folded in block order by the last block to arrive
};

struct StreamWork {
  long long Bc;               // batch-chunk capacity (columns)
  float* act[BFX_MAX_LAYERS + 1];  // act[0] = gathered x (d x Bc); act[l+1] = output of layer l (nout x Bc); FP32 copies
                                   // exist only where an FP32 kernel consumes them (need_f32)
  int need_f32[BFX_MAX_LAYERS + 1];
  float* delta[2];            // ping-pong deltas (max width x Bc), FP32 (consumers on the FP32 path only)
  float* yb;                  // [Bc] targets
  float* dl;                  // [Bc] likelihood deltas (scaled by num_batches)
  float* part;                // split-K partials of the FP32 weight-gradient GEMMs
  long long part_floats;
  float* rowpart;             // [64][max width] partial row sums (1-wide output layer on the FP32 path)
  // tensor-core mode: BF16 hi / lo halves of the activations, the deltas and theta
  __nv_bfloat16* act_hi[BFX_MAX_LAYERS + 1];
  __nv_bfloat16* act_lo[BFX_MAX_LAYERS + 1];
  __nv_bfloat16* delta_hi[2];
  __nv_bfloat16* delta_lo[2];
  __nv_bfloat16* w_hi;
  __nv_bfloat16* w_lo;
  unsigned* mask[BFX_MAX_LAYERS + 1];   // ReLU sign bits of act[l] (word (b / 32) * width + i), tensor-core layers
  float* tc_part[BFX_MAX_LAYERS];       // per tensor-core layer: split-K partials of [dW] (ksplit x nout x nin)
  int tc_ksplit[BFX_MAX_LAYERS];
  float* tc_rowpart[BFX_MAX_LAYERS];    // per layer: bias-gradient partials (slots x nout) left by whoever produced its delta
  int tc_rowslots[BFX_MAX_LAYERS];
  float* head_part;                     // fused head kernel: [blocks][2 * nin + 1] (dW_out, row sums of its delta, db_out)
  int head_blocks;
};

// tcgen05 GEMMs of a Dense layer (bfx_stream_tc.cu)
cudaError_t tc2_forward(const tcg::Args& g, const tcg::Operand& A, const tcg::Operand& B, const tcg::Output& O, cudaStream_t st);
cudaError_t tc2_dx(const tcg::Args& g, const tcg::Operand& A, const tcg::Operand& B, const tcg::Output& O, cudaStream_t st);
cudaError_t tc2_dw(const tcg::Args& g, const tcg::Operand& A, const tcg::Operand& B, int ksplit, cudaStream_t st);

cudaError_t stream_work_alloc(const StreamModel& M, long long Bc, StreamWork& W);
void stream_work_free(StreamWork& W);

// The minibatch of one evaluation.  Either explicit (idx / full) or the scheduled one of DataLoader semantics
// (abstract.jl:102-105): global step g -> epoch g / nbat, batch k = g % nbat = positions [k B, min(n, (k+1) B)) of
// the epoch's keyed permutation.  gstep >= 0 fixes g on the host, gstep < 0 reads the chain's device counter.
struct StreamBatch {
  const long long* idx;  // explicit indices or nullptr
  long long n;           // data-set size
  long long B;           // columns of this evaluation (what the host sizes the launches for)
  long long Bsched;      // scheduled: the DataLoader batch size (B < Bsched only for the short last batch of an epoch)
  int full;              // 1: batch = [0, n) in order
  int scheduled;         // 1: derive first / keys from the step counter
  long long nbat;
  long long gstep;
  unsigned long long seed;
  unsigned chainkey;
  int shuffle;
};

// Bucketed exchange of a minibatch-sharded chain: as soon as the backward pass has finished the gradient of a layer,
// that slice of the exchange buffer is all-reduced on a side stream while the main stream continues with the GEMMs of
// the layers below; the slice of the output layer carries the three likelihood sums (g[P .. P+2]).  Fork / join by
// events, so the whole pattern is capturable in the update graph.
struct StreamBucketHook {
  cudaError_t (*reduce)(void* ctx, float* ptr, long long count, cudaStream_t st);
  void* ctx;
  cudaStream_t side;
  cudaEvent_t ev_fork[BFX_MAX_LAYERS];
  cudaEvent_t ev_join;
  int used;  // set by stream_eval_like: 1 = the buckets were issued (the caller joins on ev_join), 0 = not applicable
};

// ---- peer-memory exchange of the minibatch-sharded chain (NVLink / NVSwitch) -----------------------
// Every rank exports its exchange buffer [g ; ds0 ; ds1 ; B] (P + 3 floats) and a block of flags with CUDA IPC; the
// finish kernel of every rank then READS all ranks' buffers over NVLink and sums them in rank order while it applies
// the prior: the all-reduce and the finish pass are one kernel, the sums are bit-identical on every rank, and nothing
// but two flag words per peer crosses the fabric besides the payload.  Flags of rank q (in q's memory, written by the
// peers with system-scope stores): ready[j] = last epoch whose local gradient rank j has completed, done[j] = last
// epoch whose buffers rank j has finished reading, epoch = updates this rank has finished.
constexpr int PEER_MAX = 8;
//
// Two protocols.  PULL (2 ranks): every rank reads the whole buffer of every peer ((n - 1) x 2.24 MB in).  SCATTER
// (more ranks): rank r pulls only ITS 1/n slice of every buffer, finishes it (prior, norms) and PUSHES the finished
// slice into every rank's result buffer (2 x (n - 1) / n x 2.24 MB per rank over the fabric, independent of n); the
// partial norms travel with a sliced[r] flag, a one-block join kernel waits for all slices and folds the norms in rank
// order before the update kernels run.  sliced[] doubles as "done reading" of the next epoch's first kernel.
constexpr int PEER_READY = 0, PEER_DONE = PEER_MAX, PEER_EPOCH = 2 * PEER_MAX, PEER_SLICED = 3 * PEER_MAX,
              PEER_NORMS = 4 * PEER_MAX /* 2 doubles per rank */, PEER_WORDS = 6 * PEER_MAX;
enum { PEER_MODE_PULL = 1, PEER_MODE_SCATTER = 2 };
struct PeerDev {
  int n, self;                          // n <= 1: no exchange
  int mode, pad_;
  const float* buf[PEER_MAX];           // exchange buffers of all ranks (own one included)
  float* sum[PEER_MAX];                 // result buffers of all ranks (SCATTER pushes finished slices into them)
  unsigned long long* flg[PEER_MAX];    // flag blocks of all ranks
};
// SCATTER: the slice kernel and the join that makes the norms / the gathered gradient available to the update kernels
cudaError_t stream_finish_scatter(const StreamModel& M, const float* theta, StreamScalars* sc, float nb, const PeerDev& pd,
                                  cudaStream_t st);
// behind the local gradient: publish epoch + 1 in every rank's ready[self]
cudaError_t stream_peer_signal(const PeerDev& pd, cudaStream_t st);
// stream_pack_sums and the signal in one launch
cudaError_t stream_pack_sums_signal(float* g, long long P, StreamScalars* sc, const PeerDev& pd, cudaStream_t st);
// stream_finish with the exchange folded in: g_out[i] = sum_r buf[r][i] - c2 theta (...), likelihood sums from the tails
cudaError_t stream_finish_peer(const StreamModel& M, const float* theta, float* g_out, StreamScalars* sc, float nb,
                               const PeerDev& pd, cudaStream_t st);
// stream_zero that first waits until every peer has finished reading this rank's buffer of the previous epoch
cudaError_t stream_zero_peer(float* g, long long P, StreamScalars* sc, const PeerDev& pd, cudaStream_t st);

// likelihood part of the evaluation on one minibatch: g[0..Pnet) += num_batches * d loglike / d theta_net
// (g must be zeroed by the caller), sc->ds0/ds1/bcount += sums.  theta_split_valid: W.w_hi / w_lo already hold
// the BF16 halves of theta (the sampler updates keep them current).
cudaError_t stream_eval_like(const StreamModel& M, StreamWork& W, const float* theta, const float* x, const float* y,
                             const StreamBatch& b, float nb, bool want_grad, float* g, StreamScalars* sc,
                             cudaStream_t st, float* yhat_out = nullptr, bool theta_split_valid = false,
                             StreamBucketHook* hook = nullptr);
// prior + sigma prior + value + ||g||^2 (one kernel; nothing is synchronised).  from_buffer: the likelihood sums are
// taken from g[P .. P+2] (the all-reduced exchange buffer) instead of the scalars.
cudaError_t stream_finish(const StreamModel& M, const float* theta, float* g, StreamScalars* sc, float nb, bool want_grad,
                          bool from_buffer, cudaStream_t st);
// scalars -> gbuf tail (minibatch-sharded chains all-reduce [g ; ds0 ; ds1 ; bcount] as P + 3 floats)
cudaError_t stream_pack_sums(float* g, long long P, StreamScalars* sc, cudaStream_t st);
cudaError_t stream_zero(float* g, long long P, StreamScalars* sc, cudaStream_t st);
cudaError_t stream_set_gstep(StreamScalars* sc, long long gstep, cudaStream_t st);
// BF16 halves of theta for the tensor-core layers (the sampler updates keep them current afterwards)
cudaError_t stream_split_theta(const StreamModel& M, StreamWork& W, const float* theta, cudaStream_t st);

struct StreamUpdate {
  const SamplerDev* S;   // host copy
  long long P;
  float* theta;
  float* mom;
  float* minv;  // nullptr: identity
  float* rmsv;
  float* g;
  StreamScalars* sc;
  unsigned long long seed;
  long long gstep;       // >= 0: Philox step on the host's count; < 0: the chain's device counter
  unsigned chain;
  const float* inj;      // injected normals (P floats) or nullptr
  // recording (all optional): explicit slot, or the device-derived slot of a ring / run buffer
  float* sample_out;     // where to record theta after the update (device, P floats) or nullptr
  float* samples_base;   // [slots][P]: sample ns (1-based) with ns % keep_every == 0 goes to slot ns / keep_every - 1 - kept_origin
  long long keep_every, kept_origin, ring_slots;  // ring_slots > 0: slot modulo ring_slots
  float* kinetic_base;   // [hist_stride]: entry t - 1 - hist_origin
  long long hist_origin, hist_stride;
  float* kinetic_out;    // explicit device scalar slot or nullptr
  __nv_bfloat16* w_hi;   // BF16 halves of the updated theta for the tensor-core layers (nullptr: not kept)
  __nv_bfloat16* w_lo;
};
cudaError_t stream_update(const StreamUpdate& U, cudaStream_t st);

}  // namespace bfx
