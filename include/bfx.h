/*
 * bfx.h  --  C ABI of libbfx, the B200-native engine for the BayesFlux hot path.
 *
 * The reference (enweg/BayesFlux.jl v0.2.4) has no FFI of its own: its seams
 * are the Julia generic functions cited beside every entry point below.  A
 * Julia shim (julia/BayesFluxB200.jl) translates the
 * reference's objects into these plain-C descriptors and `ccall`s the entry
 * points; the Python host mirror (bayesflux.jl_b200/) binds the very same
 * symbols through ctypes.
 *
 * Conventions
 *   - every function returns an int32 status (BFX_OK == 0); the message of the
 *     last failure on the calling thread is available via bfx_last_error().
 *     Nothing aborts, nothing throws across the boundary.
 *   - all host buffers are owned by the caller; the library copies in/out and
 *     keeps no host pointer after returning.
 *   - Float32 everywhere, `double` only for returned log-density values.
 *   - Julia column-major: theta is P x C (one chain per column), samples are
 *     P x nkept x C, feed-forward x is d x n, sequence x is T x d x n.
 *   - indices are 0-based on this side of the boundary.
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     returns BFX_ERR_CUDA.
 *
 * All file:line citations are relative to the reference tree.
 */
#ifndef BFX_H
#define BFX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BFX_VERSION 100 /* 0.1.0 */

/* ---- status codes ------------------------------------------------------ */
enum {
  BFX_OK = 0,
  BFX_ERR_BAD_ARG = 1,
  BFX_ERR_UNSUPPORTED = 2, /* e.g. FullCovMassAdapter, nets beyond the smem regime */
  BFX_ERR_CUDA = 3,
  BFX_ERR_NAN_G = 4,     /* error("NaN in g")  src/inference/mcmc/sgnht.jl:70-72   */
  BFX_ERR_NAN_THETA = 5, /* error("NaN in θ")  src/inference/mcmc/sgnht.jl:81-83, ggmc.jl:160-162 */
  BFX_ERR_NO_DATA = 6,
  BFX_ERR_STATE = 7,
  BFX_ERR_PEER = 8 /* minibatch-sharded chain: a peer did not publish its part of an update within about a minute */
};

/* ---- model description ------------------------------------------------- */
/* layer kinds: src/layers/dense.jl:13-26, src/layers/recurrent.jl:13-34,41-68 */
enum { BFX_DENSE = 0, BFX_RNN = 1, BFX_LSTM = 2 };
/* activations kept from the template layer (dense.jl:14,23) */
enum { BFX_IDENTITY = 0, BFX_RELU = 1, BFX_TANH = 2, BFX_SIGMOID = 3 };

typedef struct bfx_layer_desc {
  int32_t kind; /* BFX_DENSE | BFX_RNN | BFX_LSTM */
  int32_t in;
  int32_t out;
  int32_t act; /* Dense: activation; RNN: cell activation; LSTM: ignored */
} bfx_layer_desc;

/* likelihoods: FeedforwardNormal / SeqToOneNormal (feedforward.jl:21-43,
 * seq_to_one.jl:22-44) and FeedforwardTDist / SeqToOneTDist (feedforward.jl:76-97,
 * seq_to_one.jl:79-100).  Feed-forward vs seq-to-one follows the rank of x. */
enum { BFX_LIKE_NORMAL = 0, BFX_LIKE_TDIST = 1 };
/* prior on sigma, mapped through the log link (sigma = exp(theta_like)) */
enum { BFX_SIGMA_GAMMA = 0 /* Gamma(shape a, scale b) */, BFX_SIGMA_INVGAMMA = 1 /* InverseGamma(shape a, scale b) */ };

typedef struct bfx_like_desc {
  int32_t kind;
  int32_t sigma_prior_kind;
  double nu; /* TDist degrees of freedom (fixed) */
  double sigma_prior_a;
  double sigma_prior_b;
} bfx_like_desc;

/* network priors: GaussianPrior (netpriors/gaussian.jl:6-28),
 * MixtureScalePrior (netpriors/mixturescale.jl:18-33, ctor order sigma1, sigma2, pi1) */
enum { BFX_PRIOR_GAUSSIAN = 0, BFX_PRIOR_MIXTURE = 1 };

typedef struct bfx_prior_desc {
  int32_t kind;
  int32_t _pad;
  double sigma0;
  double sigma1;
  double sigma2;
  double pi1;
} bfx_prior_desc;

typedef struct bfx_model bfx_model;     /* opaque: BNN (src/model/BNN.jl:5-38) on one device */
typedef struct bfx_sampler bfx_sampler; /* opaque: MCMCState for nchains chains            */

/* ---- samplers ---------------------------------------------------------- */
enum { BFX_SGLD = 0, BFX_SGNHT = 1, BFX_SGNHTS = 2, BFX_GGMC = 3, BFX_HMC = 4,
       BFX_MODE = 5 /* find_mode: FluxModeFinder + a Flux optimiser, src/inference/mode/flux.jl:8-54 */ };
/* Flux.Optimise rules usable by BFX_MODE (Flux 0.13 semantics, applied to -g) */
enum { BFX_OPT_DESCENT = 0, BFX_OPT_MOMENTUM = 1, BFX_OPT_RMSPROP = 2, BFX_OPT_ADAM = 3 };
enum { BFX_STEP_CONSTANT = 0, BFX_STEP_DUAL_AVERAGING = 1 };
enum { BFX_MASS_FIXED = 0, BFX_MASS_DIAGCOV = 1, BFX_MASS_RMSPROP = 2, BFX_MASS_FULLCOV = 3 /* rejected */ };

/* One struct carries the keyword sets of all five constructors:
 *   SGLD   sgld.jl:47-54      stepsize_a, stepsize_b, stepsize_gamma, min_stepsize, maxnorm
 *   SGNHT  sgnht.jl:32        l, A, xi
 *   SGNHTS sgnht-s.jl:41-47   l, sigmaA, xi, mu, madapter
 *   GGMC   ggmc.jl:59-67      beta, l, steps, maxnorm, sadapter, madapter
 *   HMC    hmc.jl:59-65       l, path_len, maxnorm, sadapter, madapter
 *   MODE   mode/flux.jl:18-31 opt_*, mode_windowlength, mode_abstol  (find_mode = bfx_mcmc_run with
 *                             numsamples = epochs * num_batches update steps and samples_out = NULL; the
 *                             result is BFX_STATE_THETA, convergence per chain BFX_STATE_CONVERGED)
 *   DualAveragingStepSize dualaveragestepsize.jl:25-32; DiagCovMassAdapter
 *   diagcovariancemassadapter.jl:23-33; RMSPropMassAdapter rmspropmassadapter.jl:17-24 */
typedef struct bfx_sampler_desc {
  int32_t kind;
  int32_t steps;    /* GGMC delayed-acceptance window */
  int32_t path_len; /* HMC leapfrog stages */
  int32_t sadapter_kind;
  int32_t madapter_kind;
  int32_t da_t0;
  int32_t da_adapt_steps;
  int32_t ma_adapt_steps;
  int32_t ma_windowlength;
  int32_t _pad;
  float stepsize_a, stepsize_b, stepsize_gamma, min_stepsize;
  float maxnorm;
  float l;
  float A;      /* SGNHT diffusion */
  float sigmaA; /* SGNHTS diffusion */
  float xi;     /* initial thermostat */
  float mu;
  float beta;
  float da_target_accept, da_gamma, da_kappa;
  float ma_kappa, ma_epsilon; /* DiagCov */
  float ma_lambda, ma_alpha;  /* RMSProp */
  /* BFX_MODE: FluxModeFinder(bnn, opt; windowlength, ϵ) flux.jl:18-31 and the optimiser's hyper-parameters
   * (Descent η; Momentum η, ρ; RMSProp η, ρ, ϵ; ADAM η, β1 = opt_rho, β2, ϵ) */
  int32_t opt_kind;
  int32_t mode_windowlength;
  float opt_eta, opt_rho, opt_beta2, opt_eps;
  float mode_abstol;
  float _pad2;
} bfx_sampler_desc;

/* state fields for bfx_sampler_get_state / set_state (checkpoint / resume,
 * mirrors the mutable fields of the reference structs).
 * Scope: these fields restore a chain exactly on a LIVE handle (continue_sampling) and, across processes, for the
 * samplers whose whole state they are (SGLD; SGNHT / SGNHTS with a fixed mass).  The adapters' internal state
 * (dual-averaging sums, RMSProp's V, the DiagCov window, GGMC's position in its window) and the engine's own step
 * counter (batch schedule, Philox counters) are not exported: a GGMC / HMC / adapting chain restored in a new process
 * restarts its adaptation and its batch / noise streams.  Setting BFX_STATE_THETA drops cached log posteriors. */
enum {
  BFX_STATE_THETA = 0,    /* float  P x C */
  BFX_STATE_MOMENTUM = 1, /* float  P x C   (p / momentum) */
  BFX_STATE_MINV = 2,     /* float  P x C   diagonal of the inverse mass matrix */
  BFX_STATE_XI = 3,       /* float  C       thermostat */
  BFX_STATE_STEPSIZE = 4, /* float  C       l */
  BFX_STATE_T = 5,        /* int64  C       step counter t (1-based like the reference) */
  BFX_STATE_NSAMPLED = 6, /* int64  C */
  BFX_STATE_STATUS = 7,   /* int32  C       BFX_OK | BFX_ERR_NAN_G | BFX_ERR_NAN_THETA | BFX_ERR_PEER */
  BFX_STATE_CONVERGED = 8 /* int32  C       BFX_MODE: 1 once the convergence window fell below ϵ */
};

enum { BFX_BATCH_PER_CHAIN = 0, BFX_BATCH_SHARED = 1 };

/* mcmc(bnn, batchsize, numsamples, sampler; shuffle, partial, continue_sampling, θstart)
 * src/inference/mcmc/abstract.jl:74-127, plus the engine's extensions (nchains is
 * fixed by the sampler handle; keep_every, seed, batch_mode). */
typedef struct bfx_run_desc {
  int64_t batchsize;
  int64_t numsamples;        /* total target, as in the reference */
  int32_t shuffle;           /* default 1 */
  int32_t partial;           /* default 1 */
  int32_t continue_sampling; /* 1: draw numsamples - nsampled more, starting from the state */
  int32_t keep_every;        /* record sample t (1-based) iff t % keep_every == 0; default 1 */
  int32_t batch_mode;        /* BFX_BATCH_PER_CHAIN (reference semantics) | BFX_BATCH_SHARED */
  int32_t _pad;
  uint64_t seed;        /* Philox key */
  int64_t chain_offset; /* global id of this handle's chain 0: chains sharded over several GPUs / ranks
                           keep distinct noise and batch streams (chain c uses stream chain_offset + c) */
  /* multi-GPU partitioning of the run (no reference counterpart: the reference is one process).
   *   BFX_SHARD_CHAINS     every rank / device runs its own chains (chain_offset = first global chain id), the data
   *                        set is replicated, NO collective
   *   BFX_SHARD_MINIBATCH  ONE chain, every rank holds an equal shard of the data and a replica of the sampler
   *                        (bfx_sampler_comm_init); batchsize is the LOCAL batch size, chain_offset the rank; each
   *                        update! all-reduces the P + 3 floats [gradient ; likelihood sums] over NCCL / NVLink inside
   *                        bfx_mcmc_run / bfx_mcmc_advance, theta stays bit-identical on all ranks
   * ngpus = number of ranks / devices taking part (informational for CHAINS, checked for MINIBATCH). */
  int32_t ngpus;      /* default 1 */
  int32_t shard_mode; /* BFX_SHARD_NONE | BFX_SHARD_CHAINS | BFX_SHARD_MINIBATCH */
} bfx_run_desc;

enum { BFX_SHARD_NONE = 0, BFX_SHARD_CHAINS = 1, BFX_SHARD_MINIBATCH = 2 };

/* ---- library ----------------------------------------------------------- */
int32_t bfx_version(void);
/* copies the calling thread's last error message (NUL-terminated) into buf */
int32_t bfx_last_error(char* buf, size_t n);
int32_t bfx_device_count(int32_t* count);

/* ---- model: destruct(::Chain) + likelihood + prior + BNN ---------------- *
 * replaces  destruct  src/model/deconstruct.jl:45-58,  BNN(...)  src/model/BNN.jl:32-38 */
int32_t bfx_model_create(const bfx_layer_desc* layers, int32_t nlayers, const bfx_like_desc* like,
                         const bfx_prior_desc* prior, int32_t device, bfx_model** out);
int32_t bfx_model_destroy(bfx_model* m);
/* num_total_params (BNN.jl:36) and like.nc.num_params_network */
int32_t bfx_model_num_params(const bfx_model* m, int64_t* p_total, int64_t* p_net);
/* 0-based [start,end) of layer i in theta_net  (NetConstructor.starts/ends, deconstruct.jl:24-25) */
int32_t bfx_model_layer_bounds(const bfx_model* m, int32_t layer, int64_t* start, int64_t* end);
/* bnn.x / bnn.y: one H2D copy, Julia column-major as is.  ndims 2: dims = {d, n};
 * ndims 3: dims = {T, d, n}  (make_rnn_tensor layout, src/utils/rnn_utils.jl:11-21) */
int32_t bfx_model_set_data(bfx_model* m, const float* x, const int64_t* dims, int32_t ndims, const float* y,
                           int64_t n);
/* same, but x and y are DEVICE pointers on the model's device (copied device-to-device) */
int32_t bfx_model_set_data_device(bfx_model* m, const float* x_dev, const int64_t* dims, int32_t ndims,
                                  const float* y_dev, int64_t n);

/* Arithmetic of the contractions (no reference counterpart: the reference has one Float32 path).
 *   BFX_COMPUTE_FP32      FP32 FMA on the CUDA cores -- the correctness gate, every model
 *   BFX_COMPUTE_TC_BF16X3 tcgen05 tensor cores, 3-term BF16 split with FP32 accumulation in TMEM
 *                         (same 1e-4 gradient bar).  Shared-memory regime: MLPs whose hidden widths are
 *                         multiples of 16 and <= 64, input dimension <= 64.  Large-P regime: every layer
 *                         with fan-out >= 64 and fan-in >= 16, both multiples of 8 (128 x 128 x 64 tcgen05
 *                         GEMM tiles); narrower layers of the same net stay FP32.  Otherwise
 *                         BFX_ERR_UNSUPPORTED and the model stays on the FP32 path. */
enum { BFX_COMPUTE_FP32 = 0, BFX_COMPUTE_TC_BF16X3 = 1 };
int32_t bfx_model_set_compute(bfx_model* m, int32_t mode);
int32_t bfx_model_get_compute(const bfx_model* m, int32_t* mode);

/* ---- log posterior and its gradient (parity entries) -------------------- *
 * loglikeprior(bnn, θ, x, y; num_batches)   src/model/BNN.jl:50-56
 * ∇loglikeprior(bnn, θ, x, y; num_batches)  src/model/BNN.jl:61-69
 * theta: P x C host floats.  idx == NULL: the full data set; otherwise B
 * observation indices (0-based) per chain, chain c reading idx + c*idx_chain_stride
 * (stride 0 = one batch shared by all chains).  v_out: C doubles. g_out: P x C. */
int32_t bfx_loglikeprior(bfx_model* m, const float* theta, int32_t nchains, const int64_t* idx, int64_t B,
                         int64_t idx_chain_stride, float num_batches, double* v_out);
int32_t bfx_grad_loglikeprior(bfx_model* m, const float* theta, int32_t nchains, const int64_t* idx, int64_t B,
                              int64_t idx_chain_stride, float num_batches, double* v_out, float* g_out);

/* ---- predictive helpers (forward only, many theta x many inputs) ------------------------------------- *
 * yhat = vec(net(x)) of posterior_predict (src/likelihoods/feedforward.jl:45-55,99-110, seq_to_one.jl:46-57,
 * 102-114) for every column of theta (P x C): yhat_out is n x C column-major, sigma_out[c] = invlink(theta_like)
 * = exp(theta_like).  x_new == NULL: the model's own data.  The random draw around yhat (Normal or sigma * TDist)
 * stays on the host side like every other RNG of the reference's post-processing
 * (sample_posterior_predict, src/model/BNN.jl:129-135). */
int32_t bfx_predict(bfx_model* m, const float* theta, int32_t nchains, const float* x_new, const int64_t* dims,
                    int32_t ndims, float* yhat_out, float* sigma_out);

/* ---- chain diagnostics over recorded samples --------------------------------- *
 * The reference ships none: its README hands samples' to MCMCChains (README.md:191-193).  samples: [C][n][P]
 * floats = the P x n x C column-major array bfx_mcmc_run returns (host pointer) or the device ring of
 * bfx_mcmc_advance with ring_slots = n (samples_on_device = 1).  Every chain is split in halves (the middle draw
 * of an odd n is dropped): m = 2C sequences of nh = n / 2 draws.  Outputs are P host floats each, any may be NULL:
 * mean / var (ddof 1) of the m * nh draws, split-Rhat = sqrt(((nh-1)/nh W + B/nh) / W), and the effective sample
 * size m nh / tau with tau from Geyer's initial positive + monotone sequence of the pair sums of
 * rho_t = 1 - (W - mean_s acov_s(t)) / var+ (Stan reference manual, no rank normalisation), lags t <= max_lag
 * (max_lag <= 0: all nh - 1 lags). */
int32_t bfx_chain_diagnostics(int32_t device, const float* samples, int32_t samples_on_device, int64_t P, int64_t n,
                              int64_t C, int64_t max_lag, float* mean, float* var, float* rhat, float* ess);

/* ---- samplers ----------------------------------------------------------- *
 * SGLD/SGNHT/SGNHTS/GGMC/HMC constructors + initialise!  (sgld.jl:47-88 etc.) for
 * nchains independent chains living on the model's device. */
int32_t bfx_sampler_create(const bfx_sampler_desc* desc, bfx_model* m, int32_t nchains, bfx_sampler** out);
/* The model must outlive every other call on the sampler; destroy itself does not touch the model, so a garbage
 * collector may finalise the two handles in either order. */
int32_t bfx_sampler_destroy(bfx_sampler* s);
/* initialise!(sampler, θ, numsamples; continue_sampling=false): zero momentum, reset
 * counters/adapters, theta_start P x C host floats. */
int32_t bfx_sampler_init(bfx_sampler* s, const float* theta_start);
int32_t bfx_sampler_get_state(bfx_sampler* s, int32_t field, void* out, int64_t nbytes);
int32_t bfx_sampler_set_state(bfx_sampler* s, int32_t field, const void* in, int64_t nbytes);

/* update!(sampler, θ, bnn, ∇θ) with every random draw supplied by the caller
 * (sgld.jl:96-112, sgnht.jl:64-90, sgnht-s.jl:81-116, ggmc.jl:140-198, hmc.jl:113-159).
 * One call = one update! of every chain on the minibatch idx (layout as above).
 * normals: ndraws x P x C floats (SGLD/SGNHT/SGNHTS/HMC: 1 draw, GGMC: 2);
 * uniforms: C floats (GGMC/HMC accept test) or NULL.  theta_out (P x C) may be NULL. */
int32_t bfx_sampler_step_injected(bfx_sampler* s, const int64_t* idx, int64_t B, int64_t idx_chain_stride,
                                  float num_batches, const float* normals, const float* uniforms,
                                  float* theta_out);

/* the whole mcmc loop on the device (abstract.jl:102-126).
 * theta_start: P x C or NULL with continue_sampling.  samples_out: P x nkept x C where
 * nkept = number of NEW recorded samples (bfx_mcmc_num_kept); may be NULL (throughput
 * runs that only want the final state).  kinetic_out: nnew x C (SGNHT/SGNHTS) or NULL;
 * accepted_out: nnew x C int32 (GGMC/HMC) or NULL.  Returns BFX_ERR_NAN_* if any chain
 * hit a NaN guard (per-chain codes via BFX_STATE_STATUS); other chains keep going. */
int32_t bfx_mcmc_num_kept(const bfx_sampler* s, const bfx_run_desc* run, int64_t* nnew, int64_t* nkept);
int32_t bfx_mcmc_run(bfx_sampler* s, const bfx_run_desc* run, const float* theta_start, float* samples_out,
                     float* kinetic_out, int32_t* accepted_out);
/* progress poll from another host thread (ProgressMeter replacement, abstract.jl:122) */
int32_t bfx_progress(const bfx_sampler* s, int64_t* nsampled);

/* device-resident variant (throughput runs, pipelines that consume the samples on the
 * device): runs nsteps update! calls for all chains without touching host memory.  The state
 * stays on the device; recorded samples (run->keep_every) go to the caller's DEVICE ring
 * samples_dev laid out [C][ring_slots][P] (kept sample k of a chain lands in slot
 * k % ring_slots), or nowhere when samples_dev is NULL.
 * Asynchronous: returns after enqueueing.  stream_handle: a cudaStream_t or 0 (own stream). */
int32_t bfx_mcmc_advance(bfx_sampler* s, const bfx_run_desc* run, int64_t nsteps, void* stream_handle,
                         float* samples_dev, int64_t ring_slots);
/* Optional: page-lock a caller-owned host buffer that is passed to bfx_mcmc_run again and again (theta_start,
 * samples_out).  Transfers to / from registered memory are DMA copies without the driver's bounce buffer; with
 * several processes per box this removes the contention for host memory bandwidth.  The caller keeps the buffer
 * alive (GC.@preserve / a held reference) until bfx_host_unregister.  Thin wrappers of cudaHostRegister. */
int32_t bfx_host_register(void* ptr, size_t bytes);
int32_t bfx_host_unregister(void* ptr);

/* number of kernel launches issued by this sampler handle so far */
int32_t bfx_sampler_launch_count(const bfx_sampler* s, int64_t* n);

/* ---- large-P regime and the minibatch-sharded single chain (SURVEY.md section 8e, configs[2]) ------------ *
 * Models whose theta + gradient + one batch tile exceed a CTA's shared memory run in the "stream" regime:
 * theta in HBM in the flat order, layer-wise GEMMs, SGLD / SGNHT / SGNHTS updates as fused passes with
 * grid-wide reductions.  Every entry above works unchanged in that regime.  One chain can additionally be
 * MINIBATCH-SHARDED over ranks (no reference counterpart: the reference is a single process): every rank
 * holds its shard of the data and an identical copy of the sampler, and each update! is
 *     bfx_stream_grad_local   -> this rank's likelihood part of the gradient into the CALLER's device buffer of
 *                                count = P + 3 floats  [num_batches * dloglike/dtheta ; sum1 ; sum2 ; B_local]
 *     (caller)                   all-reduce(sum) of that buffer over the ranks (NCCL over NVLink)
 *     bfx_stream_apply_update -> prior and sigma-prior terms added once, update! with identical Philox
 *                                streams on every rank, so theta / p / xi stay replicated without a broadcast
 * Both calls only enqueue work on stream_handle (a cudaStream_t, 0 = the engine's stream).
 * run->batchsize is the LOCAL batch size, num_batches the GLOBAL number of batches. */
enum { BFX_REGIME_SMEM = 0, BFX_REGIME_STREAM = 1 };
int32_t bfx_model_regime(const bfx_model* m, int32_t* regime);
int32_t bfx_stream_grad_local(bfx_sampler* s, const bfx_run_desc* run, float num_batches, void* stream_handle,
                              float* buf_dev, int64_t count);
int32_t bfx_stream_apply_update(bfx_sampler* s, const bfx_run_desc* run, float num_batches, void* stream_handle,
                                float* buf_dev, float* sample_dev /* P floats on the device, or NULL */);
/* the same exchange done by the library: ncclAllReduce(sum) of buf_dev over the sampler's communicator
 * (bfx_sampler_comm_init below) on stream_handle; a no-op without a communicator.  bfx_mcmc_run / bfx_mcmc_advance
 * with shard_mode MINIBATCH issue exactly this between the two halves; the three-call form exists for callers that
 * want to time or interleave the phases. */
int32_t bfx_stream_allreduce(bfx_sampler* s, void* stream_handle, float* buf_dev);

/* In-library collective of the minibatch-sharded chain: one NCCL communicator per sampler replica.  NCCL is bound at
 * run time (dlopen of libnccl.so.2, or the path in $BFX_NCCL_LIB), so a process that already carries NCCL (torch,
 * NCCL.jl) shares that copy.  Rank 0 creates the 128-byte id, the caller distributes it (MPI, torch.distributed, a
 * file ...), every rank calls bfx_sampler_comm_init (collective).  With a communicator attached and
 * run->shard_mode == BFX_SHARD_MINIBATCH, bfx_mcmc_run / bfx_mcmc_advance perform
 *     local gradient  ->  sum of the P + 3 floats over the ranks  ->  replicated update
 * for every update!, captured together in one CUDA graph per update.  The sum: bfx_sampler_comm_init also exports
 * every rank's exchange buffer with CUDA IPC; when all ranks can open all peers (one box, NVLink / NVSwitch), the
 * finish kernel of each rank reads the peers' buffers directly and sums them in rank order (exchange and finish pass
 * are one kernel, flags in peer memory order it: no collective call); otherwise, and with BFX_NO_P2P=1,
 * ncclAllReduce on the engine's stream.  bfx_sampler_peer_ranks reports which one is in use. */
#define BFX_COMM_ID_BYTES 128
int32_t bfx_comm_unique_id(void* id_out /* BFX_COMM_ID_BYTES */);
int32_t bfx_sampler_comm_init(bfx_sampler* s, const void* id, int32_t nranks, int32_t rank);
int32_t bfx_sampler_comm_destroy(bfx_sampler* s);
/* ranks whose exchange buffers this sampler reads over peer memory (0 or 1: the NCCL all-reduce is used instead) */
int32_t bfx_sampler_peer_ranks(const bfx_sampler* s, int32_t* n);
/* number of update! graphs replayed by this sampler so far (stream regime: one cudaGraphLaunch per update!) */
int32_t bfx_sampler_graph_replays(const bfx_sampler* s, int64_t* n);

/* ---- test hooks: expose the engine's own random streams so the oracle can
 * replay a device run draw for draw -------------------------------------- */
/* the minibatch (indices into the data set) chain `chain` uses at global step `step` */
int32_t bfx_debug_batch_indices(const bfx_sampler* s, const bfx_run_desc* run, int64_t chain, int64_t step,
                                int64_t* idx_out, int64_t* B_out);
/* the P standard normals of draw slot `slot` and the uniform used at global step `step` */
int32_t bfx_debug_noise(const bfx_sampler* s, const bfx_run_desc* run, int64_t chain, int64_t step, int32_t slot,
                        float* normals_out, float* uniform_out);

#ifdef __cplusplus
}
#endif
#endif /* BFX_H */
