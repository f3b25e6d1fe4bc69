// Shared host/device descriptors of the bfx engine (sm_100a).
//
// Device-side picture of one BNN (reference: src/model/BNN.jl:5-38):
//   * flat theta  = [theta_net ; theta_like]  exactly as the reference lays it out
//     (src/model/deconstruct.jl:45-58, src/layers/dense.jl:13-26, recurrent.jl:13-68);
//     this is the HBM layout (one row of P floats per chain) and the ABI layout.
//   * inside a chain's CTA the parameters live in shared memory in a PADDED layout:
//     every weight matrix keeps the reference's column-major order (W[o,i] at
//     i*ldw + o) but its leading dimension ldw is rounded so that ldw/4 is odd
//     (conflict-free 128-bit shared loads for all three GEMM shapes) and the row
//     count is rounded to 4 (zero filled).  SegDesc maps flat <-> padded.
#pragma once
#include <cstdint>

#define BFX_MAX_LAYERS 8
#define BFX_MAX_SEGS (3 * BFX_MAX_LAYERS + 1)

namespace bfx {

enum { L_DENSE = 0, L_RNN = 1, L_LSTM = 2 };
enum { A_IDENTITY = 0, A_RELU = 1, A_TANH = 2, A_SIGMOID = 3 };
enum { LK_NORMAL = 0, LK_TDIST = 1 };
enum { SP_GAMMA = 0, SP_INVGAMMA = 1 };
enum { S_SGLD = 0, S_SGNHT = 1, S_SGNHTS = 2, S_GGMC = 3, S_HMC = 4, S_MODE = 5 };
enum { OPT_DESCENT = 0, OPT_MOMENTUM = 1, OPT_RMSPROP = 2, OPT_ADAM = 3 };
enum { SA_CONST = 0, SA_DUAL = 1 };
enum { MA_FIXED = 0, MA_DIAGCOV = 1, MA_RMSPROP = 2 };

// one contiguous run of the flat theta (a weight matrix, a bias vector, theta_like)
struct SegDesc {
  int flat_off;  // first flat index
  int count;     // number of flat entries
  int s_off;     // offset in the padded shared-memory parameter block
  int rows;      // fast (contiguous) dimension of the matrix (= out); count for vectors
  int ld;        // padded leading dimension in shared memory (== rows for vectors)
  float rcp_rows;
};

struct LayerDev {
  int kind, nin, nout, act;
  int nout_pad;  // round4(rows of the gate/output block): Dense out, RNN H, LSTM 4H
  int nin_pad;   // round4(nin)
  int H;         // hidden size for recurrent layers (== nout)
  int ldw;       // leading dim of W / Wi in smem
  int ldh;       // leading dim of Wh in smem (recurrent only)
  int w_s, wh_s, b_s;  // smem offsets: W (or Wi), Wh, bias
};

// ---- tensor-core (tcgen05) path: MLPs whose hidden widths are multiples of 16 and <= 64 ----
// Hidden layer l = Dense(Hin -> Hout): its input activations (X for l = 0) live in a bf16 "core
// block" tile of 64 batch rows x ccols columns (ccols = round16(Hin) + 8; column round16(Hin) is a
// constant 1 so that the weight-gradient GEMM also yields the bias gradient), its weights in a tile
// of kpad rows (fan-in) x hout columns.  Every tile exists twice (hi and lo halves of the 3-term
// BF16 split).  Byte offsets are relative to the 128-byte aligned tensor-core region of the CTA.
struct TcLayer {
  int kpad, ccols, hout;
  int a_off, a_bytes;  // input-activation tile (hi); lo at a_off + a_bytes
  int w_off, w_bytes;  // weight tile (hi); lo at w_off + w_bytes
  int dw_col;          // TMEM column of the dW accumulator (64 rows = fan-out, ccols columns)
};

struct TcDesc {
  int enabled;
  int nh;              // hidden layers on the tensor cores (= nlayers - 1; the 1-wide output layer runs in the epilogue)
  int d_off, d_bytes;  // delta tile 64 x 64 (hi); lo at d_off + d_bytes
  int misc_off;        // float yh[128], float glast[72], uint64 bar[4], uint32 tmem slot, float gpart[4][65]
  int total_bytes;
  int tmem_cols;       // allocation (power of two >= 32); the Z accumulator sits at column 0
  int smem_pad;        // extra dynamic smem requested at launch so that co-resident CTAs never over-subscribe TMEM
  TcLayer L[BFX_MAX_LAYERS];
};

struct ModelDev {
  // Philox round keys of the run's seed (filled per launch by the launcher, 0 = not set): the noise pass reads them
  // as constant-bank operands instead of re-deriving the key schedule for every quad
  unsigned philox_rk[20];
  int philox_rk_set;
  int nlayers;
  int P, Pnet;   // total / network parameters
  int S;         // floats in the padded smem parameter block (multiple of 4)
  int nsegs;
  int like_s;    // smem offset of theta_like
  int like_kind, sprior_kind;
  float nu;
  float sprior_a, sprior_b;
  float prior_c2;      // logprior = Pnet*c0 - 0.5*c2*sum(theta_net^2)
  double prior_c0n;    // Pnet*c0
  double tdist_const;  // lgamma((nu+1)/2) - lgamma(nu/2) - 0.5*log(nu*pi)
  double sprior_const; // constant of logpdf(prior_sigma)
  double sprior_rb;    // 1 / sprior_b (the gradient path multiplies instead of dividing in double)
  int d_in;            // features per observation (per time step)
  int seq;             // 1: x is T x d x n
  int T;
  int rec_layer;       // index of the recurrent layer or -1
  int act_off[BFX_MAX_LAYERS + 1];   // float offsets of the activation buffers in smem
  int act_rows[BFX_MAX_LAYERS + 1];  // padded row counts
  int act_total;                     // floats of activation smem for one batch tile
  int rec_hp, rec_c, rec_dc, rec_cp, rec_g;  // recurrent path: float offsets of h_{t-1}, c_t, dc, c_{t-1}, gates
  int rec_rows;                      // rows of one per-step scratch record (LSTM 6H, RNN H)
  LayerDev L[BFX_MAX_LAYERS];
  SegDesc seg[BFX_MAX_SEGS];
  TcDesc tc;
  // device tables (uploaded once per model): flat index J -> padded slot, padded slot -> J (-1 = padding),
  // 4 validity bits per padded quad
  const int* flat_pad;          // [P]
  const int* pad_flat;          // [S]
  const unsigned char* pad_mask;  // [S/4]
};

// ---- shape policies of the tensor-core kernels -------------------------------------------------------------------
// The fused tensor-core kernel reads every tile offset, leading dimension and layer width through one of these.
// TcDyn forwards to the run-time descriptors (any eligible model).  TcFixed<...> restates the host's layout rules
// (bfx_api.cu: build_model_dev / build_tc) as constant expressions for one MLP shape with two hidden layers, so that
// the layer loops unroll and all address arithmetic folds; the launcher uses such an instance only after `matches`
// has compared every field with the run-time descriptors, everything else runs the TcDyn instance.
#if defined(__CUDACC__)
#define BFX_HD __host__ __device__
#else
#define BFX_HD
#endif
struct TcDyn {
  static constexpr bool fixed = false;
  BFX_HD static int nh(const ModelDev& M) { return M.tc.nh; }
  BFX_HD static const TcLayer& tl(const ModelDev& M, int l) { return M.tc.L[l]; }
  BFX_HD static const LayerDev& ld(const ModelDev& M, int l) { return M.L[l]; }
  BFX_HD static int d_off(const ModelDev& M) { return M.tc.d_off; }
  BFX_HD static int d_bytes(const ModelDev& M) { return M.tc.d_bytes; }
  BFX_HD static int S(const ModelDev& M) { return M.S; }
  BFX_HD static int like_s(const ModelDev& M) { return M.like_s; }
  BFX_HD static int d_in(const ModelDev& M) { return M.d_in; }
  BFX_HD static int misc_off(const ModelDev& M) { return M.tc.misc_off; }
};
template <int DIN, int H1, int H2, int ACT1, int ACT2>
struct TcFixed {
  static constexpr bool fixed = true;
  BFX_HD static constexpr int r4(int v) { return (v + 3) & ~3; }
  BFX_HD static constexpr int oddld(int v) { return ((r4(v) >> 2) & 1) ? r4(v) : r4(v) + 4; }
  BFX_HD static constexpr int nin(int l) { return l == 0 ? DIN : l == 1 ? H1 : H2; }
  BFX_HD static constexpr int nout(int l) { return l == 0 ? H1 : l == 1 ? H2 : 1; }
  BFX_HD static constexpr int nh(const ModelDev&) { return 2; }
  BFX_HD static constexpr TcLayer tl(const ModelDev&, int l) {
    TcLayer L{};
    int off = 0, col = 64;
    for (int i = 0; i <= l; ++i) {
      L.kpad = (nin(i) + 15) & ~15;
      L.ccols = L.kpad + 8;
      L.hout = nout(i);
      L.a_bytes = 64 * L.ccols * 2;
      L.w_bytes = L.kpad * L.hout * 2;
      L.a_off = off;
      off += 2 * L.a_bytes;
      L.w_off = off;
      off += 2 * L.w_bytes;
      off = (off + 127) & ~127;
      L.dw_col = col;
      col += L.ccols;
    }
    return L;
  }
  BFX_HD static constexpr int d_off(const ModelDev& M) {
    const TcLayer L = tl(M, 1);
    return (L.w_off + 2 * L.w_bytes + 127) & ~127;
  }
  BFX_HD static constexpr int d_bytes(const ModelDev&) { return 64 * 64 * 2; }
  BFX_HD static constexpr LayerDev ld(const ModelDev&, int l) {
    LayerDev L{};
    int s_off = 0;
    for (int i = 0; i <= l; ++i) {
      L.kind = L_DENSE;
      L.nin = nin(i);
      L.nout = nout(i);
      L.act = i == 0 ? ACT1 : i == 1 ? ACT2 : A_IDENTITY;
      L.H = L.nout;
      L.nin_pad = r4(L.nin);
      L.nout_pad = r4(L.nout);
      L.ldw = oddld(L.nout);
      L.ldh = 0;
      L.w_s = s_off;
      s_off += L.nin_pad * L.ldw;
      L.wh_s = 0;
      L.b_s = s_off;
      s_off += L.nout_pad;
    }
    return L;
  }
  BFX_HD static constexpr int like_s(const ModelDev& M) { return ld(M, 2).b_s + 4; }
  BFX_HD static constexpr int S(const ModelDev& M) { return like_s(M) + 4; }
  BFX_HD static constexpr int d_in(const ModelDev&) { return DIN; }
  BFX_HD static constexpr int misc_off(const ModelDev& M) { return d_off(M) + 2 * d_bytes(M); }
  // host: does the run-time model have exactly this layout?
  static bool matches(const ModelDev& M) {
    if (!M.tc.enabled || M.nlayers != 3 || M.tc.nh != 2 || M.rec_layer >= 0 || M.seq) return false;
    if (M.S != S(M) || M.like_s != like_s(M) || M.d_in != DIN) return false;
    if (M.tc.d_off != d_off(M) || M.tc.d_bytes != d_bytes(M) || M.tc.misc_off != misc_off(M)) return false;
    for (int l = 0; l < 3; ++l) {
      const LayerDev a = ld(M, l);
      const LayerDev& b = M.L[l];
      if (a.kind != b.kind || a.nin != b.nin || a.nout != b.nout || a.act != b.act || a.nin_pad != b.nin_pad ||
          a.nout_pad != b.nout_pad || a.ldw != b.ldw || a.w_s != b.w_s || a.b_s != b.b_s)
        return false;
    }
    for (int l = 0; l < 2; ++l) {
      const TcLayer a = tl(M, l);
      const TcLayer& b = M.tc.L[l];
      if (a.kpad != b.kpad || a.ccols != b.ccols || a.hout != b.hout || a.a_off != b.a_off || a.a_bytes != b.a_bytes ||
          a.w_off != b.w_off || a.w_bytes != b.w_bytes || a.dw_col != b.dw_col)
        return false;
    }
    return true;
  }
};

// sampler hyper-parameters (device copy of bfx_sampler_desc, see include/bfx.h)
struct SamplerDev {
  int kind, steps, path_len, sadapter_kind, madapter_kind;
  int da_t0, da_adapt_steps, ma_adapt_steps, ma_windowlength;
  float stepsize_a, stepsize_b, stepsize_gamma, min_stepsize, maxnorm;
  float l, A, sigmaA, xi, mu, beta;
  float da_target_accept, da_gamma, da_kappa, da_mu;
  float ma_kappa, ma_epsilon, ma_lambda, ma_alpha;
  // find_mode (src/inference/mode/flux.jl): Flux optimiser rule + convergence window
  int opt_kind, mode_windowlength;
  float opt_eta, opt_rho, opt_beta2, opt_eps, mode_abstol;
};

// per-chain scalar state kept in HBM between launches (mutable fields of the
// reference's sampler and adapter structs)
struct ChainScalars {
  long long t;          // sampler step counter, 1-based like the reference
  long long nsampled;
  float xi;             // thermostat (SGNHT / SGNHTS)
  float l;              // current step size
  float lMH;            // GGMC running log MH ratio
  int current_step;     // GGMC / HMC position inside the window / trajectory
  // DualAveragingStepSize (adapters/stepsize/dualaveragestepsize.jl:11-23)
  int da_t;
  float da_error_sum;
  float da_log_avg;
  // mass adapters: own counter t (diagcovariancemassadapter.jl:13-20, rmspropmassadapter.jl:9-16)
  int ma_t;
  int status;           // 0 | BFX_ERR_NAN_G | BFX_ERR_NAN_THETA
  double start_lp;      // GGMC / HMC: loglikeprior of the full data at the start of the window / trajectory
  long long accepted_total;
  // find_mode: FluxModeFinder.i, convergence flag, ADAM's running beta powers
  int mode_i, converged;
  float bp1, bp2;
  // GGMC / HMC: start_lp is the full-data log posterior of the CURRENT theta iff lp_epoch == RunDev::data_epoch + 1.
  // The MH test leaves it behind (the end-of-window value on accept, the old start value on reject, which is exactly
  // what the reference recomputes at ggmc.jl:147 / hmc.jl:135 for the next window), so every window after the first
  // costs one full-data pass instead of two.  Reset by anything that changes theta or the data from outside.
  int lp_epoch, pad_;
};

// everything a launch needs (passed by value as a __grid_constant__ kernel parameter)
// division by a per-launch constant: q = hi64(n * m) for n < 2^32 (m = floor(2^64 / d) + 1, exact while n * d < 2^64),
// so the per-step schedule arithmetic (batch of the epoch, kept-sample slot) costs a multiply, not a 64-bit divide
struct FastDiv {
  unsigned long long m;  // 0: use the plain division
  unsigned long long d;
};
inline FastDiv make_fastdiv(long long dv) {
  FastDiv f{0ull, dv > 0 ? (unsigned long long)dv : 1ull};
  if (f.d > 1 && f.d < (1ull << 32)) f.m = ~0ull / f.d + 1ull;
  return f;
}

struct RunDev {
  ModelDev M;
  SamplerDev S;
  const float* x;
  const float* y;
  const uint4* xs;  // tensor-core path: x pre-split into bf16 hi/lo rows (bfx_api.cu ensure_xsplit), or nullptr
  long long n;
  int C;
  unsigned chain0;  // global id of local chain 0 (noise / permutation keys)
  // per-chain state in the PADDED layout, row stride Ps = M.S floats
  int Ps;
  float* theta;
  float* mom;
  float* minv;
  float* theta_old;
  float* rms_v;
  float* ring;  // [C][w][Ps] DiagCov window
  float* mode_window;  // [C][mode_windowlength] find_mode: max |theta change| of the last steps
  ChainScalars* sc;
  // batch schedule (src/inference/mcmc/abstract.jl:102-105,117-118)
  long long B, nb;
  float num_batches;
  int shuffle, batch_shared;
  const long long* idx;  // explicit minibatch (injected step) or nullptr
  long long idx_chain_stride;
  unsigned long long seed;
  long long gstep0;  // global update! count before this launch
  int nsteps;
  // injected noise (parity entry) or nullptr
  const float* inj_normals;  // [ndraw][C][P]
  const float* inj_uniforms; // [C]
  // outputs
  float* samples;            // [C][slots][P] or nullptr
  long long samples_chain_stride;
  long long kept_origin;     // number of kept samples that precede slot 0
  long long ring_slots;      // > 0: the sample buffer is a ring of this many slots per chain
  int keep_every;
  float* kinetic_out;        // [C][hist_stride]
  int* accepted_out;         // [C][hist_stride]
  long long hist_stride, hist_origin;
  unsigned long long* progress;
  float* scratch;              // recurrent path: per-CTA scratch, scratch_stride floats each
  long long scratch_stride;
  // dynamic (chain, chunk) scheduling of a launch: work[0] = item counter, work[1 + c] = finished chunks of chain c
  int* work;
  int chunk_len;               // update! calls per work item
  long long grid_cap;          // upper bound of the CTA pool (recurrent scratch is sized by it)
  // a launch may cover only the chains [c_begin, c_begin + c_count) (c_count = 0: all C): bfx_mcmc_run pipelines
  // chain groups so that the host copies of one group overlap the kernel of the next
  int c_begin, c_count;
  int data_epoch;  // bumped by bfx_model_set_data* / set_compute: invalidates cached full-data log posteriors
  // filled by the launcher from nb, keep_every, ring_slots, S.ma_windowlength
  FastDiv fd_nb, fd_keep, fd_slots, fd_window;
};

// evaluation-only launch (bfx_loglikeprior / bfx_grad_loglikeprior)
struct EvalDev {
  ModelDev M;
  const float* x;
  const float* y;
  const uint4* xs;  // as RunDev::xs
  long long n;
  int C;
  const float* theta;  // [C][P] flat (ABI layout)
  const long long* idx;
  long long idx_chain_stride;
  long long B;
  float num_batches;
  int want_grad;
  double* v_out;  // [C]
  float* g_out;   // [C][P]
  float* scratch;
  long long scratch_stride;
  float* yhat_out;  // [C][B] network outputs (bfx_predict) or nullptr
};

}  // namespace bfx
