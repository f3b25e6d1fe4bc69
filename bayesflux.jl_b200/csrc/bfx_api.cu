// C ABI of libbfx (include/bfx.h): handles, descriptor translation, memory, launches.
// No exceptions cross the boundary; every entry returns a status and records a message.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include <dlfcn.h>

#include "../../include/bfx.h"
#include "bfx_common.cuh"
#include "bfx_kernels.h"
#include "bfx_stream.cuh"

using namespace bfx;

namespace {

thread_local std::string g_err;

int32_t fail(int32_t code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return fail(BFX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));              \
  } while (0)

inline int round4(int v) { return (v + 3) & ~3; }
inline int odd_ld(int v) {
  int ld = round4(v);
  if (((ld >> 2) & 1) == 0) ld += 4;
  return ld;
}

constexpr size_t kMaxSmem = 227 * 1024;

}  // namespace

// ---------------------------------------------------------------------------------
// NCCL, bound at run time (dlopen): the library has no link-time dependency on it, a process that already carries
// NCCL (torch, a Julia NCCL.jl session) shares that copy.  Only the minibatch-sharded chain uses it.
// ---------------------------------------------------------------------------------
namespace {
struct NcclId { char internal[128]; };
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*CommInitAll)(void**, int, const int*) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string err;
};
NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.handle ? &api : nullptr;
  tried = true;
  const char* names[] = {std::getenv("BFX_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* nme : names) {
    if (!nme || !*nme) continue;
    api.handle = dlopen(nme, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
    api.err = dlerror();
  }
  if (!api.handle) return nullptr;
  api.GetUniqueId = (int (*)(NcclId*))dlsym(api.handle, "ncclGetUniqueId");
  api.CommInitRank = (int (*)(void**, int, NcclId, int))dlsym(api.handle, "ncclCommInitRank");
  api.CommInitAll = (int (*)(void**, int, const int*))dlsym(api.handle, "ncclCommInitAll");
  api.CommDestroy = (int (*)(void*))dlsym(api.handle, "ncclCommDestroy");
  api.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(api.handle, "ncclAllReduce");
  api.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(api.handle, "ncclAllGather");
  api.GetErrorString = (const char* (*)(int))dlsym(api.handle, "ncclGetErrorString");
  if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce) {
    api.err = "NCCL symbols missing";
    api.handle = nullptr;
    return nullptr;
  }
  return &api;
}
constexpr int kNcclFloat32 = 7, kNcclInt8 = 0, kNcclSum = 0;
}  // namespace


struct bfx_model {
  int device = 0;
  ModelDev M{};
  LaunchCfg cfg{256, 64};
  std::vector<bfx_layer_desc> layers;
  std::vector<int64_t> starts, ends;
  float* x = nullptr;
  float* y = nullptr;
  int64_t n = 0;
  void* xsplit = nullptr;  // tensor-core path: x as bf16 hi/lo rows (4 * kpad bytes per observation), built lazily
  int xsplit_kpad = 0;
  int data_epoch = 0;      // bumped whenever the data or the compute mode change (RunDev::data_epoch)
  cudaStream_t stream = nullptr;
  // flat (ABI) <-> padded (engine) parameter layout
  std::vector<int> flat_pad, pad_flat;
  std::vector<unsigned char> pad_mask;
  // large-P ("stream") regime, bfx_stream.cuh
  bool stream_regime = false;
  StreamModel SM{};
  StreamWork W{};
  // whose theta the BF16 halves W.w_hi / w_lo currently hold (tensor-core layers): a single-chain sampler keeps them
  // current through its updates, every other evaluation overwrites them
  const void* wsplit_owner = nullptr;
  long long wsplit_gstep = -1;
  float* scratch = nullptr;  // recurrent path: per-CTA per-step records
  size_t scratch_ctas = 0;
  int* d_flat_pad = nullptr;
  int* d_pad_flat = nullptr;
  unsigned char* d_pad_mask = nullptr;
};

struct bfx_sampler {
  bfx_model* m = nullptr;
  int device = 0;  // copy of m->device: destroying a sampler must not touch the model (it may already be gone)
  bfx_sampler_desc d{};
  SamplerDev S{};
  int C = 0;
  int Ps = 0;
  float *theta = nullptr, *mom = nullptr, *minv = nullptr, *theta_old = nullptr, *rms_v = nullptr, *ring = nullptr;
  float* mode_window = nullptr;
  ChainScalars* sc = nullptr;
  int* work = nullptr;           // [1 + C] work-item counter and per-chain chunk counters of a launch
  float* stage = nullptr;        // [C][P] flat staging of per-parameter state crossing the ABI
  float* samp_cache[2] = {nullptr, nullptr};  // device sample chunks of bfx_mcmc_run, kept between calls
  size_t samp_cap = 0;
  StreamScalars* ssc = nullptr;  // stream regime: per-chain device scalars
  float* gbuf = nullptr;         // stream regime: [P + 3] gradient + likelihood sums (the all-reduce payload)
  unsigned long long* progress_host = nullptr;  // mapped pinned
  unsigned long long* progress_dev = nullptr;
  long long gstep = 0;     // update! calls done since init
  long long nsampled = 0;  // lock-step host mirror
  long long launches = 0;
  bool initialised = false;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t adv_event = nullptr;  // recorded behind every bfx_mcmc_advance (asynchronous: the host mirror lags behind it)
  // stream regime: one update! as a CUDA graph (captured once per static description, see stream_run)
  cudaGraphExec_t upd_graph = nullptr;
  std::vector<unsigned char> upd_graph_key;
  long long graph_replays = 0;
  // minibatch-sharded chain: NCCL communicator of the ranks that hold the shards (bfx_sampler_comm_init)
  void* comm = nullptr;
  int comm_rank = 0, comm_nranks = 1;
  cudaStream_t comm_stream = nullptr;  // side stream of the bucketed exchange (fork / join by events)
  StreamBucketHook hook{};
  // peer-memory exchange (bfx_stream.cuh: PeerDev): the ranks' exchange buffers and flag blocks opened through CUDA
  // IPC; n > 1 only when EVERY rank managed to open every peer (else the NCCL all-reduce is used)
  PeerDev peer{};
  float* gsum = nullptr;                 // [P + 4] the summed gradient (the peers keep reading gbuf meanwhile)
  unsigned long long* peer_flags = nullptr;
  void* peer_opened[3 * PEER_MAX] = {};  // IPC mappings to close
};

// ---------------------------------------------------------------------------------
// model construction
// ---------------------------------------------------------------------------------
static int32_t build_model_dev(bfx_model* m, const bfx_like_desc* like, const bfx_prior_desc* prior) {
  ModelDev& M = m->M;
  const int nl = (int)m->layers.size();
  M.nlayers = nl;
  M.rec_layer = -1;
  int flat = 0, s_off = 0, nseg = 0;
  m->starts.clear();
  m->ends.clear();
  for (int l = 0; l < nl; ++l) {
    const bfx_layer_desc& d = m->layers[l];
    if (d.in <= 0 || d.out <= 0) return fail(BFX_ERR_BAD_ARG, "layer sizes must be positive");
    if (l > 0 && m->layers[l - 1].out != d.in) return fail(BFX_ERR_BAD_ARG, "layer fan-in does not match previous fan-out");
    if (d.act < BFX_IDENTITY || d.act > BFX_SIGMOID) return fail(BFX_ERR_BAD_ARG, "unknown activation");
    LayerDev& L = M.L[l];
    L.kind = d.kind;
    L.nin = d.in;
    L.nout = d.out;
    L.act = d.act;
    L.H = d.out;
    L.nin_pad = round4(d.in);
    m->starts.push_back(flat);
    if (d.kind == BFX_DENSE) {
      L.nout_pad = round4(d.out);
      L.ldw = odd_ld(d.out);
      L.ldh = 0;
      L.w_s = s_off;
      M.seg[nseg++] = SegDesc{flat, d.in * d.out, s_off, d.out, L.ldw, 1.0f / (float)d.out};
      flat += d.in * d.out;
      s_off += L.nin_pad * L.ldw;
      L.wh_s = 0;
      L.b_s = s_off;
      M.seg[nseg++] = SegDesc{flat, d.out, s_off, d.out, d.out, 1.0f / (float)d.out};
      flat += d.out;
      s_off += L.nout_pad;
    } else if (d.kind == BFX_RNN || d.kind == BFX_LSTM) {
      if (M.rec_layer >= 0) return fail(BFX_ERR_UNSUPPORTED, "at most one recurrent layer is supported");
      if (l != 0) return fail(BFX_ERR_UNSUPPORTED, "the recurrent layer must be the first layer of the chain");
      M.rec_layer = l;
      const int G = d.out * (d.kind == BFX_LSTM ? 4 : 1);
      L.nout_pad = round4(G);
      L.ldw = odd_ld(G);
      L.ldh = odd_ld(G);
      L.w_s = s_off;
      M.seg[nseg++] = SegDesc{flat, d.in * G, s_off, G, L.ldw, 1.0f / (float)G};
      flat += d.in * G;
      s_off += L.nin_pad * L.ldw;
      L.wh_s = s_off;
      M.seg[nseg++] = SegDesc{flat, d.out * G, s_off, G, L.ldh, 1.0f / (float)G};
      flat += d.out * G;
      s_off += round4(d.out) * L.ldh;
      L.b_s = s_off;
      M.seg[nseg++] = SegDesc{flat, G, s_off, G, G, 1.0f / (float)G};
      flat += G;
      s_off += L.nout_pad;
    } else {
      return fail(BFX_ERR_BAD_ARG, "unknown layer kind");
    }
    m->ends.push_back(flat);
  }
  if (m->layers[nl - 1].out != 1)
    return fail(BFX_ERR_BAD_ARG, "the likelihoods assume a single network output (feedforward.jl:13)");
  M.Pnet = flat;
  M.like_s = s_off;
  M.seg[nseg++] = SegDesc{flat, 1, s_off, 1, 1, 1.0f};
  s_off += 4;
  M.P = flat + 1;
  M.S = s_off;
  M.nsegs = nseg;
  m->flat_pad.assign(M.P, 0);
  m->pad_flat.assign(M.S, -1);
  m->pad_mask.assign(M.S / 4, 0);
  for (int sgi = 0; sgi < nseg; ++sgi) {
    const SegDesc& sg = M.seg[sgi];
    for (int j = 0; j < sg.count; ++j) {
      const int col = j / sg.rows, row = j - col * sg.rows;
      const int k = sg.s_off + col * sg.ld + row;
      m->flat_pad[sg.flat_off + j] = k;
      m->pad_flat[k] = sg.flat_off + j;
      m->pad_mask[k >> 2] |= (unsigned char)(1u << (k & 3));
    }
  }
  M.d_in = m->layers[0].in;
  // likelihood
  if (like->kind != BFX_LIKE_NORMAL && like->kind != BFX_LIKE_TDIST) return fail(BFX_ERR_BAD_ARG, "unknown likelihood");
  M.like_kind = like->kind;
  M.nu = (float)like->nu;
  if (like->kind == BFX_LIKE_TDIST && !(like->nu > 0)) return fail(BFX_ERR_BAD_ARG, "TDist needs nu > 0");
  M.tdist_const = like->kind == BFX_LIKE_TDIST
                      ? std::lgamma(0.5 * (like->nu + 1)) - std::lgamma(0.5 * like->nu) - 0.5 * std::log(like->nu * M_PI)
                      : 0.0;
  M.sprior_kind = like->sigma_prior_kind;
  M.sprior_a = (float)like->sigma_prior_a;
  M.sprior_b = (float)like->sigma_prior_b;
  M.sprior_rb = 1.0 / (double)M.sprior_b;
  if (!(like->sigma_prior_a > 0) || !(like->sigma_prior_b > 0)) return fail(BFX_ERR_BAD_ARG, "sigma prior parameters must be > 0");
  if (like->sigma_prior_kind == BFX_SIGMA_GAMMA)
    M.sprior_const = -std::lgamma(like->sigma_prior_a) - like->sigma_prior_a * std::log(like->sigma_prior_b);
  else if (like->sigma_prior_kind == BFX_SIGMA_INVGAMMA)
    M.sprior_const = like->sigma_prior_a * std::log(like->sigma_prior_b) - std::lgamma(like->sigma_prior_a);
  else
    return fail(BFX_ERR_BAD_ARG, "unknown sigma prior");
  // network prior: logprior = Pnet*c0 - 0.5*c2*sum(theta^2)
  const double l2pi = std::log(2.0 * M_PI);
  double c0, c2;
  if (prior->kind == BFX_PRIOR_GAUSSIAN) {
    if (!(prior->sigma0 > 0)) return fail(BFX_ERR_BAD_ARG, "sigma0 must be > 0");
    c0 = -0.5 * l2pi - std::log(prior->sigma0);
    c2 = 1.0 / (prior->sigma0 * prior->sigma0);
  } else if (prior->kind == BFX_PRIOR_MIXTURE) {
    if (!(prior->sigma1 > 0) || !(prior->sigma2 > 0)) return fail(BFX_ERR_BAD_ARG, "sigma1/sigma2 must be > 0");
    const double pi2 = 1.0 - prior->pi1;
    c0 = prior->pi1 * (-0.5 * l2pi - std::log(prior->sigma1)) + pi2 * (-0.5 * l2pi - std::log(prior->sigma2));
    c2 = prior->pi1 / (prior->sigma1 * prior->sigma1) + pi2 / (prior->sigma2 * prior->sigma2);
  } else {
    return fail(BFX_ERR_BAD_ARG, "unknown network prior");
  }
  M.prior_c0n = c0 * M.Pnet;
  M.prior_c2 = (float)c2;
  // launch shape: tiny nets get one warp per chain, wide ones 256 threads
  int max_rows = 0;
  for (int l = 0; l < nl; ++l) max_rows = std::max(max_rows, M.L[l].nout_pad);
  if (M.rec_layer >= 0) {
    if (nl < 2 || M.L[1].kind != L_DENSE)
      return fail(BFX_ERR_UNSUPPORTED, "a recurrent layer must be followed by a Dense head");
    m->cfg = LaunchCfg{256, 16};  // 16 sequences per tile: the per-step records of a tile stay small (L2 resident)
  } else if (max_rows >= 48)
    m->cfg = LaunchCfg{256, 64};
  else if (max_rows >= 16)
    m->cfg = LaunchCfg{128, 64};
  else
    m->cfg = LaunchCfg{32, 32};
  const int LDB = m->cfg.bt + 4;
  int off = 0;
  M.act_rows[0] = round4(M.d_in);
  M.act_off[0] = 0;
  off = M.act_rows[0] * LDB;
  for (int l = 0; l < nl; ++l) {
    // the output of a recurrent layer is h (H rows), not its gate block
    const int rows = M.L[l].kind == L_DENSE ? M.L[l].nout_pad : round4(M.L[l].H);
    M.act_rows[l + 1] = rows;
    M.act_off[l + 1] = off;
    off += rows * LDB;
  }
  if (M.rec_layer >= 0) {
    const int hp = round4(M.L[0].H) * LDB;
    M.rec_hp = off; off += hp;
    M.rec_c = off; off += hp;
    M.rec_dc = off; off += hp;
    M.rec_cp = off; off += hp;
    M.rec_g = off; off += M.L[0].nout_pad * LDB;
    M.rec_rows = M.L[0].kind == L_LSTM ? 6 * M.L[0].H : M.L[0].H;
  }
  M.act_total = off;
  if (smallp_smem_bytes(M, m->cfg.bt) > kMaxSmem) {
    if (M.rec_layer >= 0)
      return fail(BFX_ERR_UNSUPPORTED,
                  "recurrent network does not fit the shared-memory regime (theta + gradient + one batch tile > 227 KB)");
    // large-P regime: theta in HBM in the flat order, layer-wise GEMMs (bfx_stream.cuh)
    m->stream_regime = true;
    StreamModel& SMd = m->SM;
    SMd = StreamModel{};
    SMd.nlayers = nl;
    SMd.P = M.P;
    SMd.Pnet = M.Pnet;
    SMd.d_in = M.d_in;
    SMd.like_kind = M.like_kind;
    SMd.sprior_kind = M.sprior_kind;
    SMd.nu = M.nu;
    SMd.sprior_a = M.sprior_a;
    SMd.sprior_b = M.sprior_b;
    SMd.prior_c2 = M.prior_c2;
    SMd.prior_c0n = M.prior_c0n;
    SMd.tdist_const = M.tdist_const;
    SMd.sprior_const = M.sprior_const;
    for (int l = 0; l < nl; ++l) {
      SMd.L[l].nin = M.L[l].nin;
      SMd.L[l].nout = M.L[l].nout;
      SMd.L[l].act = M.L[l].act;
      SMd.L[l].w_off = m->starts[l];
      SMd.L[l].b_off = m->starts[l] + (long long)M.L[l].nin * M.L[l].nout;
    }
    // the engine-internal layout of this regime IS the flat order (rows padded to a multiple of 4)
    M.S = round4(M.P);
    m->flat_pad.resize(M.P);
    m->pad_flat.assign(M.S, -1);
    m->pad_mask.assign(M.S / 4, 0);
    for (int J = 0; J < M.P; ++J) {
      m->flat_pad[J] = J;
      m->pad_flat[J] = J;
      m->pad_mask[J >> 2] |= (unsigned char)(1u << (J & 3));
    }
  }
  return BFX_OK;
}

// tensor-core path: eligibility + tile / TMEM layout (bfx_tc.cuh)
static int32_t build_tc(bfx_model* m) {
  ModelDev& M = m->M;
  TcDesc T{};
  const int nl = M.nlayers;
  if (nl < 2) return fail(BFX_ERR_UNSUPPORTED, "tensor-core path: needs at least one hidden layer");
  for (int l = 0; l < nl; ++l)
    if (M.L[l].kind != L_DENSE) return fail(BFX_ERR_UNSUPPORTED, "tensor-core path: Dense layers only");
  if (M.L[0].nin > 64) return fail(BFX_ERR_UNSUPPORTED, "tensor-core path: input dimension must be <= 64");
  for (int l = 0; l + 1 < nl; ++l)
    if (M.L[l].nout % 16 != 0 || M.L[l].nout > 64)
      return fail(BFX_ERR_UNSUPPORTED, "tensor-core path: hidden widths must be multiples of 16 and <= 64");
  T.nh = nl - 1;
  int off = 0, col = 64;  // TMEM columns [0,64): the Z / dA accumulator
  for (int l = 0; l < T.nh; ++l) {
    TcLayer& L = T.L[l];
    L.kpad = (M.L[l].nin + 15) & ~15;
    L.ccols = L.kpad + 8;
    L.hout = M.L[l].nout;
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
  T.d_bytes = 64 * 64 * 2;
  T.d_off = off;
  off += 2 * T.d_bytes;
  T.misc_off = off;
  off += 2304 + 128 * T.nh;  // yh[128], glast[72], 4 barriers, tmem slot, gpart[4][65]; then one IssueDesc per layer
  T.total_bytes = off;
  int cols = 32;
  while (cols < col) cols <<= 1;
  if (cols > 512) return fail(BFX_ERR_UNSUPPORTED, "tensor-core path: accumulators exceed the 512 TMEM columns");
  T.tmem_cols = cols;
  T.enabled = 1;
  TcDesc saved = M.tc;
  M.tc = T;
  M.tc.smem_pad = 0;
  size_t need = smallp_smem_bytes(M, 64);
  if (need > kMaxSmem) {
    M.tc = saved;
    return fail(BFX_ERR_UNSUPPORTED, "tensor-core path: working set exceeds 227 KB of shared memory");
  }
  // CTAs co-resident on an SM share 512 TMEM columns: never let more CTAs fit by shared memory than by TMEM
  const int by_tmem = 512 / cols;
  const size_t min_smem = (kMaxSmem + 1024) / (size_t)(by_tmem + 1) + 1;
  if (need < min_smem && min_smem <= kMaxSmem) M.tc.smem_pad = (int)(min_smem - need);
  return BFX_OK;
}

// tensor-core path: the bf16 hi/lo copy of x the tiles are staged from (one pass over x per data set / first layer width)
static int32_t ensure_xsplit(bfx_model* m) {
  if (!m->M.tc.enabled || !m->x || m->M.seq) return BFX_OK;
  const int kpad = m->M.tc.L[0].kpad;
  if (m->xsplit && m->xsplit_kpad == kpad) return BFX_OK;
  if (m->xsplit) cudaFree(m->xsplit);
  m->xsplit = nullptr;
  CK(cudaMalloc(&m->xsplit, (size_t)m->n * 4 * (size_t)kpad));
  CK(launch_split_x(m->x, m->xsplit, m->n, m->M.d_in, kpad, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  m->xsplit_kpad = kpad;
  return BFX_OK;
}

// ---------------------------------------------------------------------------------
extern "C" {

int32_t bfx_model_set_compute(bfx_model* m, int32_t mode) {
  if (!m) return fail(BFX_ERR_BAD_ARG, "model is NULL");
  if (mode != BFX_COMPUTE_FP32 && mode != BFX_COMPUTE_TC_BF16X3) return fail(BFX_ERR_BAD_ARG, "unknown compute mode");
  m->data_epoch += 1;  // values of the two engines differ in the last digits: cached log posteriors are dropped
  if (m->stream_regime) {
    // large-P regime: per layer, the three GEMMs go to the tcgen05 kernel when the layer is a real GEMM whose
    // operands can be fetched in aligned 16-byte chunks (bfx_stream_tc.cu)
    int any = 0;
    for (int l = 0; l < m->SM.nlayers; ++l) {
      StreamLayer& L = m->SM.L[l];
      L.tc = mode == BFX_COMPUTE_TC_BF16X3 && L.nout >= 64 && L.nin >= 16 && L.nout % 8 == 0 && L.nin % 8 == 0 &&
             L.w_off % 8 == 0;
      any |= L.tc;
    }
    if (mode == BFX_COMPUTE_TC_BF16X3 && !any)
      return fail(BFX_ERR_UNSUPPORTED, "tensor-core path: no layer of this network is a 16-byte aligned GEMM of width >= 64");
    m->SM.tc = any;
    if (m->W.Bc) {  // the workspace depends on the mode
      cudaSetDevice(m->device);
      if (m->stream) cudaStreamSynchronize(m->stream);
      stream_work_free(m->W);
    }
    m->wsplit_owner = nullptr;
    return BFX_OK;
  }
  if (mode == BFX_COMPUTE_FP32) {
    m->M.tc = TcDesc{};
    return BFX_OK;
  }
  return build_tc(m);
}

int32_t bfx_model_get_compute(const bfx_model* m, int32_t* mode) {
  if (!m || !mode) return fail(BFX_ERR_BAD_ARG, "NULL argument");
  *mode = (m->M.tc.enabled || m->SM.tc) ? BFX_COMPUTE_TC_BF16X3 : BFX_COMPUTE_FP32;
  return BFX_OK;
}

int32_t bfx_version(void) { return BFX_VERSION; }

int32_t bfx_last_error(char* buf, size_t n) {
  if (!buf || n == 0) return BFX_ERR_BAD_ARG;
  std::snprintf(buf, n, "%s", g_err.c_str());
  return BFX_OK;
}

int32_t bfx_device_count(int32_t* count) {
  if (!count) return fail(BFX_ERR_BAD_ARG, "count is NULL");
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) {
    *count = 0;
    return fail(BFX_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
  }
  *count = c;
  return BFX_OK;
}

int32_t bfx_model_create(const bfx_layer_desc* layers, int32_t nlayers, const bfx_like_desc* like,
                         const bfx_prior_desc* prior, int32_t device, bfx_model** out) {
  if (!layers || !like || !prior || !out) return fail(BFX_ERR_BAD_ARG, "NULL argument");
  if (nlayers < 1 || nlayers > BFX_MAX_LAYERS) return fail(BFX_ERR_BAD_ARG, "nlayers must be in 1..8");
  bfx_model* m = new (std::nothrow) bfx_model();
  if (!m) return fail(BFX_ERR_BAD_ARG, "out of host memory");
  m->device = device;
  m->layers.assign(layers, layers + nlayers);
  int32_t st = build_model_dev(m, like, prior);
  if (st != BFX_OK) {
    delete m;
    return st;
  }
  *out = m;
  return BFX_OK;
}

int32_t bfx_model_destroy(bfx_model* m) {
  if (!m) return BFX_OK;
  if (m->x || m->y || m->stream) {
    cudaSetDevice(m->device);
    if (m->x) cudaFree(m->x);
    if (m->y) cudaFree(m->y);
    if (m->xsplit) cudaFree(m->xsplit);
    if (m->stream) cudaStreamDestroy(m->stream);
  }
  if (m->scratch) {
    cudaSetDevice(m->device);
    cudaFree(m->scratch);
  }
  if (m->W.Bc) {
    cudaSetDevice(m->device);
    stream_work_free(m->W);
  }
  if (m->d_flat_pad) {
    cudaSetDevice(m->device);
    cudaFree(m->d_flat_pad);
    cudaFree(m->d_pad_flat);
    cudaFree(m->d_pad_mask);
  }
  delete m;
  return BFX_OK;
}

int32_t bfx_model_num_params(const bfx_model* m, int64_t* p_total, int64_t* p_net) {
  if (!m) return fail(BFX_ERR_BAD_ARG, "model is NULL");
  if (p_total) *p_total = m->M.P;
  if (p_net) *p_net = m->M.Pnet;
  return BFX_OK;
}

int32_t bfx_model_layer_bounds(const bfx_model* m, int32_t layer, int64_t* start, int64_t* end) {
  if (!m || layer < 0 || layer >= (int)m->layers.size()) return fail(BFX_ERR_BAD_ARG, "bad layer index");
  if (start) *start = m->starts[layer];
  if (end) *end = m->ends[layer];
  return BFX_OK;
}

static int32_t ensure_stream(bfx_model* m) {
  CK(cudaSetDevice(m->device));
  if (!m->stream) CK(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
  if (!m->d_flat_pad) {
    CK(cudaMalloc(&m->d_flat_pad, m->flat_pad.size() * sizeof(int)));
    CK(cudaMalloc(&m->d_pad_flat, m->pad_flat.size() * sizeof(int)));
    CK(cudaMalloc(&m->d_pad_mask, m->pad_mask.size()));
    CK(cudaMemcpy(m->d_flat_pad, m->flat_pad.data(), m->flat_pad.size() * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(m->d_pad_flat, m->pad_flat.data(), m->pad_flat.size() * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(m->d_pad_mask, m->pad_mask.data(), m->pad_mask.size(), cudaMemcpyHostToDevice));
    m->M.flat_pad = m->d_flat_pad;
    m->M.pad_flat = m->d_pad_flat;
    m->M.pad_mask = m->d_pad_mask;
  }
  return BFX_OK;
}

// recurrent path: scratch for `ctas` concurrently launched CTAs (T per-step records of one tile each)
static int32_t ensure_scratch(bfx_model* m, size_t ctas, float** out, long long* stride) {
  *out = nullptr;
  *stride = 0;
  if (m->M.rec_layer < 0) return BFX_OK;
  const long long st = (long long)m->M.T * m->M.rec_rows * m->cfg.bt;
  if (ctas > m->scratch_ctas) {
    if (m->scratch) {
      CK(cudaStreamSynchronize(m->stream));
      CK(cudaFree(m->scratch));
      m->scratch = nullptr;
      m->scratch_ctas = 0;
    }
    CK(cudaMalloc(&m->scratch, (size_t)st * ctas * sizeof(float)));
    m->scratch_ctas = ctas;
  }
  *out = m->scratch;
  *stride = st;
  return BFX_OK;
}

// host-side layout conversion of per-parameter state (parity entry only; the hot paths convert on the device)
static void padded_to_flat(const bfx_model* m, const std::vector<float>& pad, int C, float* flat) {
  const int P = m->M.P, S = m->M.S;
  for (int c = 0; c < C; ++c)
    for (int J = 0; J < P; ++J) flat[(size_t)c * P + J] = pad[(size_t)c * S + m->flat_pad[J]];
}

static int32_t set_data_impl(bfx_model* m, const float* x, const int64_t* dims, int32_t ndims, const float* y,
                             int64_t n, cudaMemcpyKind kind) {
  if (!m || !x || !y || !dims) return fail(BFX_ERR_BAD_ARG, "NULL argument");
  if (ndims != 2 && ndims != 3) return fail(BFX_ERR_BAD_ARG, "x must be d x n or T x d x n");
  const int64_t d = dims[ndims - 2], nn = dims[ndims - 1];
  if (nn != n || n <= 0) return fail(BFX_ERR_BAD_ARG, "last dimension of x must equal length(y) > 0");
  if (d != m->M.d_in) return fail(BFX_ERR_BAD_ARG, "feature dimension of x does not match the first layer");
  if (n > 2147483647LL) return fail(BFX_ERR_UNSUPPORTED, "n must fit in int32");
  if (ndims == 3 && m->M.rec_layer < 0) return fail(BFX_ERR_BAD_ARG, "3-d x needs a recurrent network");
  if (ndims == 2 && m->M.rec_layer >= 0) return fail(BFX_ERR_BAD_ARG, "recurrent networks need T x d x n input");
  int32_t st = ensure_stream(m);
  if (st) return st;
  if (m->x) cudaFree(m->x);
  if (m->y) cudaFree(m->y);
  if (m->xsplit) cudaFree(m->xsplit);
  m->x = m->y = nullptr;
  m->xsplit = nullptr;
  size_t per = (size_t)d * (ndims == 3 ? (size_t)dims[0] : 1);
  CK(cudaMalloc(&m->x, per * (size_t)n * sizeof(float)));
  CK(cudaMalloc(&m->y, (size_t)n * sizeof(float)));
  CK(cudaMemcpyAsync(m->x, x, per * (size_t)n * sizeof(float), kind, m->stream));
  CK(cudaMemcpyAsync(m->y, y, (size_t)n * sizeof(float), kind, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  m->n = n;
  m->data_epoch += 1;
  m->M.seq = ndims == 3;
  m->M.T = ndims == 3 ? (int)dims[0] : 1;
  m->scratch_ctas = 0;  // the scratch stride depends on T: force a re-allocation
  if (m->scratch) {
    cudaFree(m->scratch);
    m->scratch = nullptr;
  }
  return BFX_OK;
}

int32_t bfx_model_set_data(bfx_model* m, const float* x, const int64_t* dims, int32_t ndims, const float* y,
                           int64_t n) {
  return set_data_impl(m, x, dims, ndims, y, n, cudaMemcpyHostToDevice);
}

int32_t bfx_model_set_data_device(bfx_model* m, const float* x_dev, const int64_t* dims, int32_t ndims,
                                  const float* y_dev, int64_t n) {
  return set_data_impl(m, x_dev, dims, ndims, y_dev, n, cudaMemcpyDeviceToDevice);
}


// ---------------------------------------------------------------------------------
// stream regime helpers
// ---------------------------------------------------------------------------------
// the scheduled minibatch of the update that follows `gstep` updates (abstract.jl:102-105): sized on the host, resolved
// (positions, permutation keys) on the device.  device_step: the kernels read the chain's own step counter instead
// of `gstep` (graph replays).
static StreamBatch stream_batch(const bfx_model* m, const bfx_run_desc* run, unsigned chainkey, long long gstep,
                                bool device_step = false) {
  StreamBatch b{};
  b.idx = nullptr;
  b.n = m->n;
  long long B = run->batchsize < m->n ? run->batchsize : m->n;
  long long nb = run->partial ? (m->n + B - 1) / B : m->n / B;
  if (nb < 1) nb = 1;
  const unsigned long long k = (unsigned long long)gstep % (unsigned long long)nb;
  const long long rem = m->n - (long long)k * B;
  b.B = rem < B ? rem : B;
  b.Bsched = B;
  b.scheduled = 1;
  b.nbat = nb;
  b.gstep = device_step ? -1 : gstep;
  b.seed = run->seed;
  b.chainkey = run->batch_mode == BFX_BATCH_SHARED ? 0xFFFFFFFFu : chainkey;
  b.shuffle = run->shuffle ? 1 : 0;
  b.full = 0;
  return b;
}

static int32_t ensure_work(bfx_model* m, long long B) {
  long long Bc = B < 8192 ? B : 8192;
  if (Bc < 1) Bc = 1;
  if (m->W.Bc >= Bc) return BFX_OK;
  if (m->W.Bc) {
    CK(cudaStreamSynchronize(m->stream));
    stream_work_free(m->W);
  }
  m->wsplit_owner = nullptr;
  CK(stream_work_alloc(m->SM, Bc, m->W));
  return BFX_OK;
}

// one evaluation in the stream regime: g (P floats, device) and the scalars of `ssc`; single rank
static int32_t stream_eval_full(bfx_model* m, const float* theta_dev, const StreamBatch& b, float nb, bool want_grad,
                                float* g_dev, StreamScalars* ssc, cudaStream_t st) {
  int32_t rc = ensure_work(m, b.B);
  if (rc) return rc;
  m->wsplit_owner = nullptr;
  CK(stream_zero(g_dev, m->M.P, ssc, st));
  CK(stream_eval_like(m->SM, m->W, theta_dev, m->x, m->y, b, nb, want_grad, g_dev, ssc, st));
  CK(stream_finish(m->SM, theta_dev, g_dev, ssc, nb, want_grad, false, st));
  return BFX_OK;
}

static int32_t stream_eval_api(bfx_model* m, const float* theta, int32_t C, const int64_t* idx, int64_t B, int64_t stride,
                               float num_batches, double* v_out, float* g_out) {
  const long long P = m->M.P;
  float *d_theta = nullptr, *d_g = nullptr;
  StreamScalars* d_sc = nullptr;
  long long* d_idx = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_theta);
    cudaFree(d_g);
    cudaFree(d_sc);
    cudaFree(d_idx);
  };
#define CKS(call)                                                                     \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      cleanup();                                                                      \
      return fail(BFX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    }                                                                                 \
  } while (0)
  CKS(cudaMalloc(&d_theta, (size_t)P * 4));
  CKS(cudaMalloc(&d_g, (size_t)(P + 4) * 4));
  CKS(cudaMalloc(&d_sc, sizeof(StreamScalars)));
  CKS(cudaMemsetAsync(d_sc, 0, sizeof(StreamScalars), m->stream));
  const int64_t nidx = idx ? (stride ? (int64_t)(C - 1) * stride + B : B) : 0;
  if (idx) {
    CKS(cudaMalloc(&d_idx, (size_t)nidx * 8));
    CKS(cudaMemcpyAsync(d_idx, idx, (size_t)nidx * 8, cudaMemcpyHostToDevice, m->stream));
  }
  for (int c = 0; c < C; ++c) {
    CKS(cudaMemcpyAsync(d_theta, theta + (size_t)c * P, (size_t)P * 4, cudaMemcpyHostToDevice, m->stream));
    StreamBatch b{};
    b.n = m->n;
    b.full = idx ? 0 : 1;
    b.idx = idx ? d_idx + (long long)c * stride : nullptr;
    b.B = idx ? B : m->n;
    int32_t rc = stream_eval_full(m, d_theta, b, num_batches, g_out != nullptr, d_g, d_sc, m->stream);
    if (rc) {
      cleanup();
      return rc;
    }
    StreamScalars h;
    CKS(cudaMemcpyAsync(&h, d_sc, sizeof h, cudaMemcpyDeviceToHost, m->stream));
    if (g_out) CKS(cudaMemcpyAsync(g_out + (size_t)c * P, d_g, (size_t)P * 4, cudaMemcpyDeviceToHost, m->stream));
    CKS(cudaStreamSynchronize(m->stream));
    v_out[c] = h.value;
  }
  cleanup();
  return BFX_OK;
}

// ---------------------------------------------------------------------------------
// loglikeprior / ∇loglikeprior
// ---------------------------------------------------------------------------------
static int32_t eval_impl(bfx_model* m, const float* theta, int32_t C, const int64_t* idx, int64_t B, int64_t stride,
                         float num_batches, double* v_out, float* g_out) {
  if (!m || !theta || !v_out) return fail(BFX_ERR_BAD_ARG, "NULL argument");
  if (C <= 0) return fail(BFX_ERR_BAD_ARG, "nchains must be > 0");
  if (!m->x) return fail(BFX_ERR_NO_DATA, "bfx_model_set_data has not been called");
  if (idx && B <= 0) return fail(BFX_ERR_BAD_ARG, "B must be > 0 with explicit indices");
  int32_t st = ensure_stream(m);
  if (st) return st;
  const int P = m->M.P;
  const int64_t nidx = idx ? (stride ? (int64_t)(C - 1) * stride + B : B) : 0;
  if (idx)
    for (int64_t i = 0; i < nidx; ++i)
      if (idx[i] < 0 || idx[i] >= m->n) return fail(BFX_ERR_BAD_ARG, "batch index out of range");
  if (m->stream_regime) return stream_eval_api(m, theta, C, idx, B, stride, num_batches, v_out, g_out);
  float *d_theta = nullptr, *d_g = nullptr;
  double* d_v = nullptr;
  long long* d_idx = nullptr;
  int32_t rc = BFX_OK;
  auto cleanup = [&]() {
    cudaFree(d_theta);
    cudaFree(d_g);
    cudaFree(d_v);
    cudaFree(d_idx);
  };
#define CKC(call)                                                                     \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      cleanup();                                                                      \
      return fail(BFX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    }                                                                                 \
  } while (0)
  CKC(cudaMalloc(&d_theta, (size_t)C * P * sizeof(float)));
  CKC(cudaMalloc(&d_v, (size_t)C * sizeof(double)));
  if (g_out) CKC(cudaMalloc(&d_g, (size_t)C * P * sizeof(float)));
  if (idx) {
    CKC(cudaMalloc(&d_idx, (size_t)nidx * sizeof(long long)));
    CKC(cudaMemcpyAsync(d_idx, idx, (size_t)nidx * sizeof(long long), cudaMemcpyHostToDevice, m->stream));
  }
  CKC(cudaMemcpyAsync(d_theta, theta, (size_t)C * P * sizeof(float), cudaMemcpyHostToDevice, m->stream));
  EvalDev E{};
  E.M = m->M;
  E.x = m->x;
  E.y = m->y;
  if (int32_t xst = ensure_xsplit(m)) {
    cleanup();
    return xst;
  }
  E.xs = m->M.tc.enabled ? static_cast<const uint4*>(m->xsplit) : nullptr;
  E.n = m->n;
  E.C = C;
  E.theta = d_theta;
  E.idx = d_idx;
  E.idx_chain_stride = stride;
  E.B = idx ? B : m->n;
  E.num_batches = num_batches;
  E.want_grad = g_out ? 1 : 0;
  E.v_out = d_v;
  E.g_out = d_g;
  {
    int32_t sst = ensure_scratch(m, (size_t)C, &E.scratch, &E.scratch_stride);
    if (sst) {
      cleanup();
      return sst;
    }
  }
  CKC(launch_eval(E, m->cfg, m->stream));
  CKC(cudaMemcpyAsync(v_out, d_v, (size_t)C * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
  if (g_out) CKC(cudaMemcpyAsync(g_out, d_g, (size_t)C * P * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
  CKC(cudaStreamSynchronize(m->stream));
  cleanup();
  return rc;
}

int32_t bfx_loglikeprior(bfx_model* m, const float* theta, int32_t nchains, const int64_t* idx, int64_t B,
                         int64_t idx_chain_stride, float num_batches, double* v_out) {
  return eval_impl(m, theta, nchains, idx, B, idx_chain_stride, num_batches, v_out, nullptr);
}

int32_t bfx_grad_loglikeprior(bfx_model* m, const float* theta, int32_t nchains, const int64_t* idx, int64_t B,
                              int64_t idx_chain_stride, float num_batches, double* v_out, float* g_out) {
  if (!g_out) return fail(BFX_ERR_BAD_ARG, "g_out is NULL");
  return eval_impl(m, theta, nchains, idx, B, idx_chain_stride, num_batches, v_out, g_out);
}


// ---------------------------------------------------------------------------------
// posterior predictive support: network outputs for many theta x many inputs
// ---------------------------------------------------------------------------------
int32_t bfx_predict(bfx_model* m, const float* theta, int32_t C, const float* x_new, const int64_t* dims, int32_t ndims,
                    float* yhat_out, float* sigma_out) {
  if (!m || !theta || !yhat_out) return fail(BFX_ERR_BAD_ARG, "NULL argument");
  if (C <= 0) return fail(BFX_ERR_BAD_ARG, "nchains must be > 0");
  if (!x_new && !m->x) return fail(BFX_ERR_NO_DATA, "no x given and bfx_model_set_data has not been called");
  int32_t st = ensure_stream(m);
  if (st) return st;
  const int P = m->M.P;
  int64_t n = m->n;
  size_t per = (size_t)m->M.d_in * (m->M.seq ? (size_t)m->M.T : 1);
  if (x_new) {
    if (!dims || (ndims != 2 && ndims != 3)) return fail(BFX_ERR_BAD_ARG, "x must be d x n or T x d x n");
    if ((ndims == 3) != (m->M.rec_layer >= 0)) return fail(BFX_ERR_BAD_ARG, "rank of x does not match the network");
    if (dims[ndims - 2] != m->M.d_in) return fail(BFX_ERR_BAD_ARG, "feature dimension of x does not match the first layer");
    if (ndims == 3 && m->x && dims[0] != m->M.T) return fail(BFX_ERR_BAD_ARG, "sequence length differs from the model's data");
    n = dims[ndims - 1];
    per = (size_t)dims[ndims - 2] * (ndims == 3 ? (size_t)dims[0] : 1);
    if (n <= 0 || n > 2147483647LL) return fail(BFX_ERR_BAD_ARG, "bad number of observations");
  }
  float *d_x = nullptr, *d_y = nullptr, *d_theta = nullptr, *d_out = nullptr;
  double* d_v = nullptr;
  StreamScalars* d_sc = nullptr;
  float* d_g = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_x);
    cudaFree(d_y);
    cudaFree(d_theta);
    cudaFree(d_out);
    cudaFree(d_v);
    cudaFree(d_sc);
    cudaFree(d_g);
  };
  const float* xs = m->x;
  const float* ys = m->y;
  int savedT = m->M.T, savedSeq = m->M.seq;
  if (x_new) {
    CKC(cudaMalloc(&d_x, per * (size_t)n * 4));
    CKC(cudaMalloc(&d_y, (size_t)n * 4));
    CKC(cudaMemcpyAsync(d_x, x_new, per * (size_t)n * 4, cudaMemcpyHostToDevice, m->stream));
    CKC(cudaMemsetAsync(d_y, 0, (size_t)n * 4, m->stream));
    xs = d_x;
    ys = d_y;
    if (ndims == 3) {
      m->M.T = (int)dims[0];
      m->M.seq = 1;
    }
  }
  CKC(cudaMalloc(&d_out, (size_t)C * n * 4));
  int32_t rc = BFX_OK;
  if (m->stream_regime) {
    CKC(cudaMalloc(&d_theta, (size_t)P * 4));
    CKC(cudaMalloc(&d_g, ((size_t)P + 4) * 4));
    CKC(cudaMalloc(&d_sc, sizeof(StreamScalars)));
    CKC(cudaMemsetAsync(d_sc, 0, sizeof(StreamScalars), m->stream));  // reduction tickets start at zero
    rc = ensure_work(m, n);
    for (int c = 0; c < C && !rc; ++c) {
      CKC(cudaMemcpyAsync(d_theta, theta + (size_t)c * P, (size_t)P * 4, cudaMemcpyHostToDevice, m->stream));
      StreamBatch b{};
      b.n = n;
      b.full = 1;
      b.B = n;
      m->wsplit_owner = nullptr;
      CKC(stream_zero(d_g, P, d_sc, m->stream));
      CKC(stream_eval_like(m->SM, m->W, d_theta, xs, ys, b, 1.f, false, d_g, d_sc, m->stream, d_out + (size_t)c * n));
    }
  } else {
    CKC(cudaMalloc(&d_theta, (size_t)C * P * 4));
    CKC(cudaMalloc(&d_v, (size_t)C * sizeof(double)));
    CKC(cudaMemcpyAsync(d_theta, theta, (size_t)C * P * 4, cudaMemcpyHostToDevice, m->stream));
    EvalDev E{};
    E.M = m->M;
    E.M.tc = TcDesc{};  // forward only: the FP32 tiles
    E.x = xs;
    E.y = ys;
    E.n = n;
    E.C = C;
    E.theta = d_theta;
    E.idx = nullptr;
    E.B = n;
    E.num_batches = 1.f;
    E.want_grad = 0;
    E.v_out = d_v;
    E.g_out = nullptr;
    E.yhat_out = d_out;
    rc = ensure_scratch(m, (size_t)C, &E.scratch, &E.scratch_stride);
    if (!rc) CKC(launch_eval(E, m->cfg, m->stream));
  }
  m->M.T = savedT;
  m->M.seq = savedSeq;
  if (rc) {
    cleanup();
    return rc;
  }
  CKC(cudaMemcpyAsync(yhat_out, d_out, (size_t)C * n * 4, cudaMemcpyDeviceToHost, m->stream));
  CKC(cudaStreamSynchronize(m->stream));
  if (sigma_out)
    for (int c = 0; c < C; ++c) sigma_out[c] = std::exp(theta[(size_t)c * P + (P - 1)]);  // invlink = exp (log link)
  cleanup();
  return BFX_OK;
}

// ---------------------------------------------------------------------------------
// samplers
// ---------------------------------------------------------------------------------
static void free_sampler(bfx_sampler* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  cudaFree(s->theta);
  cudaFree(s->mom);
  cudaFree(s->minv);
  cudaFree(s->theta_old);
  cudaFree(s->rms_v);
  cudaFree(s->ring);
  cudaFree(s->mode_window);
  cudaFree(s->sc);
  cudaFree(s->work);
  cudaFree(s->stage);
  cudaFree(s->samp_cache[0]);
  cudaFree(s->samp_cache[1]);
  cudaFree(s->ssc);
  for (void*& p : s->peer_opened) {
    if (p) cudaIpcCloseMemHandle(p);
    p = nullptr;
  }
  cudaFree(s->gsum);
  cudaFree(s->peer_flags);
  cudaFree(s->gbuf);
  if (s->progress_host) cudaFreeHost(s->progress_host);
  if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
  if (s->adv_event) cudaEventDestroy(s->adv_event);
  if (s->upd_graph) {  // before the communicator: the captured collective holds a reference on it
    cudaGraphExecDestroy(s->upd_graph);
    cudaDeviceSynchronize();
  }
  if (s->comm) {
    if (NcclApi* api = nccl_api()) api->CommDestroy(s->comm);
  }
  if (s->comm_stream) {
    cudaStreamDestroy(s->comm_stream);
    for (int i = 0; i < BFX_MAX_LAYERS; ++i) cudaEventDestroy(s->hook.ev_fork[i]);
    cudaEventDestroy(s->hook.ev_join);
  }
  delete s;
}

int32_t bfx_sampler_create(const bfx_sampler_desc* d, bfx_model* m, int32_t nchains, bfx_sampler** out) {
  if (!d || !m || !out) return fail(BFX_ERR_BAD_ARG, "NULL argument");
  if (nchains <= 0) return fail(BFX_ERR_BAD_ARG, "nchains must be > 0");
  if (d->kind < BFX_SGLD || d->kind > BFX_MODE) return fail(BFX_ERR_BAD_ARG, "unknown sampler kind");
  if (d->kind == BFX_MODE) {
    if (d->opt_kind < BFX_OPT_DESCENT || d->opt_kind > BFX_OPT_ADAM) return fail(BFX_ERR_BAD_ARG, "unknown optimiser");
    if (d->mode_windowlength < 1) return fail(BFX_ERR_BAD_ARG, "windowlength must be >= 1");
    if (!(d->opt_eta > 0)) return fail(BFX_ERR_BAD_ARG, "optimiser step size must be > 0");
    if (m->stream_regime) return fail(BFX_ERR_UNSUPPORTED, "large-P (stream) regime: find_mode is not available");
  }
  if (d->madapter_kind == BFX_MASS_FULLCOV)
    return fail(BFX_ERR_UNSUPPORTED,
                "FullCovMassAdapter (dense P x P mass matrix) is outside the many-chain engine; use DiagCovMassAdapter");
  if (d->madapter_kind < BFX_MASS_FIXED || d->madapter_kind > BFX_MASS_RMSPROP)
    return fail(BFX_ERR_BAD_ARG, "unknown mass adapter");
  if (d->sadapter_kind != BFX_STEP_CONSTANT && d->sadapter_kind != BFX_STEP_DUAL_AVERAGING)
    return fail(BFX_ERR_BAD_ARG, "unknown stepsize adapter");
  if (d->kind == BFX_GGMC && d->steps < 1) return fail(BFX_ERR_BAD_ARG, "GGMC steps must be >= 1");
  if (d->kind == BFX_HMC && d->path_len < 1) return fail(BFX_ERR_BAD_ARG, "HMC path_len must be >= 1");
  if (d->kind != BFX_SGLD && d->kind != BFX_MODE && !(d->l > 0)) return fail(BFX_ERR_BAD_ARG, "stepsize l must be > 0");
  if (d->madapter_kind == BFX_MASS_DIAGCOV && d->ma_windowlength < 2)
    return fail(BFX_ERR_BAD_ARG, "DiagCov windowlength must be >= 2");
  if (m->stream_regime) {
    if (d->kind == BFX_GGMC || d->kind == BFX_HMC)
      return fail(BFX_ERR_UNSUPPORTED, "large-P (stream) regime: SGLD, SGNHT and SGNHTS are available; GGMC / HMC are not");
    if (d->kind == BFX_SGNHTS && d->madapter_kind == BFX_MASS_DIAGCOV)
      return fail(BFX_ERR_UNSUPPORTED, "large-P (stream) regime: DiagCovMassAdapter is not available");
  }
  int32_t st = ensure_stream(m);
  if (st) return st;
  bfx_sampler* s = new (std::nothrow) bfx_sampler();
  if (!s) return fail(BFX_ERR_BAD_ARG, "out of host memory");
  s->m = m;
  s->device = m->device;
  s->d = *d;
  s->C = nchains;
  s->Ps = m->M.S;
  SamplerDev& S = s->S;
  S.kind = d->kind;
  S.steps = d->kind == BFX_GGMC ? d->steps : 1;
  S.path_len = d->kind == BFX_HMC ? d->path_len : 1;
  const bool mh = d->kind == BFX_GGMC || d->kind == BFX_HMC;
  S.sadapter_kind = mh ? d->sadapter_kind : SA_CONST;
  S.madapter_kind = (d->kind == BFX_SGLD || d->kind == BFX_SGNHT || d->kind == BFX_MODE) ? MA_FIXED : d->madapter_kind;
  S.opt_kind = d->opt_kind;
  S.mode_windowlength = d->kind == BFX_MODE ? d->mode_windowlength : 1;
  S.opt_eta = d->opt_eta;
  S.opt_rho = d->opt_rho;
  S.opt_beta2 = d->opt_beta2;
  S.opt_eps = d->opt_eps;
  S.mode_abstol = d->mode_abstol;
  S.da_t0 = d->da_t0;
  S.da_adapt_steps = d->da_adapt_steps;
  S.ma_adapt_steps = d->ma_adapt_steps;
  S.ma_windowlength = d->ma_windowlength;
  S.stepsize_a = d->stepsize_a;
  S.stepsize_b = d->stepsize_b;
  S.stepsize_gamma = d->stepsize_gamma;
  S.min_stepsize = d->min_stepsize;
  S.maxnorm = d->maxnorm;
  S.l = d->l;
  S.A = d->A;
  S.sigmaA = d->sigmaA;
  S.xi = d->xi;
  S.mu = d->mu;
  S.beta = d->beta;
  S.da_target_accept = d->da_target_accept;
  S.da_gamma = d->da_gamma;
  S.da_kappa = d->da_kappa;
  S.da_mu = std::log(10.0f * d->l);  // dualaveragestepsize.jl:36
  S.ma_kappa = d->ma_kappa;
  S.ma_epsilon = d->ma_epsilon;
  S.ma_lambda = d->ma_lambda;
  S.ma_alpha = d->ma_alpha;
  const size_t row = (size_t)s->Ps * sizeof(float), all = row * (size_t)nchains;
  cudaError_t e = cudaMalloc(&s->theta, all);
  if (e == cudaSuccess) e = cudaMalloc(&s->sc, sizeof(ChainScalars) * (size_t)nchains);
  if (e == cudaSuccess && d->kind != BFX_SGLD) e = cudaMalloc(&s->mom, all);
  const bool mode = d->kind == BFX_MODE;
  if (e == cudaSuccess && (mh || mode)) e = cudaMalloc(&s->theta_old, all);
  if (e == cudaSuccess && mode) e = cudaMalloc(&s->rms_v, all);
  if (e == cudaSuccess && mode) e = cudaMalloc(&s->mode_window, sizeof(float) * (size_t)nchains * S.mode_windowlength);
  if (e == cudaSuccess && S.madapter_kind != MA_FIXED) e = cudaMalloc(&s->minv, all);
  if (e == cudaSuccess && S.madapter_kind == MA_RMSPROP) e = cudaMalloc(&s->rms_v, all);
  if (e == cudaSuccess && S.madapter_kind == MA_DIAGCOV) e = cudaMalloc(&s->ring, all * (size_t)S.ma_windowlength);
  if (e == cudaSuccess) e = cudaMalloc(&s->work, sizeof(int) * ((size_t)nchains + 1));
  if (e == cudaSuccess) e = cudaMalloc(&s->stage, sizeof(float) * (size_t)nchains * m->M.P);
  if (e == cudaSuccess && m->stream_regime) e = cudaMalloc(&s->ssc, sizeof(StreamScalars) * (size_t)nchains);
  if (e == cudaSuccess && m->stream_regime) e = cudaMalloc(&s->gbuf, ((size_t)m->M.P + 4) * sizeof(float));
  if (e == cudaSuccess) e = cudaHostAlloc(&s->progress_host, sizeof(unsigned long long), cudaHostAllocMapped);
  if (e == cudaSuccess) e = cudaHostGetDevicePointer(&s->progress_dev, s->progress_host, 0);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    free_sampler(s);
    return fail(BFX_ERR_CUDA, std::string("sampler allocation: ") + cudaGetErrorString(e));
  }
  *s->progress_host = 0;
  *out = s;
  return BFX_OK;
}

int32_t bfx_sampler_destroy(bfx_sampler* s) {
  free_sampler(s);
  return BFX_OK;
}

// per-parameter state across the ABI: host flat P x C  <->  device padded rows, converted on the device
static int32_t state_h2d(bfx_sampler* s, float* dst_pad, const float* src_flat, float fill) {
  bfx_model* m = s->m;
  const size_t bytes = (size_t)s->C * m->M.P * 4;
  CK(cudaMemcpyAsync(s->stage, src_flat, bytes, cudaMemcpyHostToDevice, m->stream));
  CK(launch_flat_to_padded(s->stage, dst_pad, m->d_flat_pad, m->M.P, s->Ps, s->C, fill, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  return BFX_OK;
}
static int32_t state_d2h(bfx_sampler* s, const float* src_pad, float* dst_flat) {
  bfx_model* m = s->m;
  const size_t bytes = (size_t)s->C * m->M.P * 4;
  CK(launch_padded_to_flat(src_pad, s->stage, m->d_flat_pad, m->M.P, s->Ps, s->C, m->stream));
  CK(cudaMemcpyAsync(dst_flat, s->stage, bytes, cudaMemcpyDeviceToHost, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  return BFX_OK;
}

// upload_theta = false: the caller brings theta in itself (bfx_mcmc_run, chain group by chain group)
static int32_t sampler_init_impl(bfx_sampler* s, const float* theta_start, bool upload_theta) {
  if (!s || !theta_start) return fail(BFX_ERR_BAD_ARG, "NULL argument");
  bfx_model* m = s->m;
  int32_t st = ensure_stream(m);
  if (st) return st;
  const size_t all = (size_t)s->Ps * sizeof(float) * (size_t)s->C;
  if (upload_theta) {
    st = state_h2d(s, s->theta, theta_start, 0.f);
    if (st) return st;
  }
  if (s->mom) CK(cudaMemsetAsync(s->mom, 0, all, m->stream));
  if (s->theta_old) CK(cudaMemsetAsync(s->theta_old, 0, all, m->stream));
  if (s->minv) {
    std::vector<float> ones((size_t)m->M.P * s->C, 1.0f);
    st = state_h2d(s, s->minv, ones.data(), 1.f);
    if (st) return st;
  }
  if (s->rms_v) CK(cudaMemsetAsync(s->rms_v, 0, all, m->stream));
  if (s->S.kind == S_MODE) {
    // FluxModeFinder: window = fill(-Inf), prevθ = fill(Inf)  (flux.jl:26-27)
    std::vector<float> inf((size_t)s->Ps * s->C, INFINITY);
    CK(cudaMemcpyAsync(s->theta_old, inf.data(), all, cudaMemcpyHostToDevice, m->stream));
    std::vector<float> ninf((size_t)s->C * s->S.mode_windowlength, -INFINITY);
    CK(cudaMemcpyAsync(s->mode_window, ninf.data(), ninf.size() * 4, cudaMemcpyHostToDevice, m->stream));
    CK(cudaStreamSynchronize(m->stream));
  }
  if (s->ring) CK(cudaMemsetAsync(s->ring, 0, all * (size_t)s->S.ma_windowlength, m->stream));
  ChainScalars z{};
  z.t = 1;
  z.nsampled = 0;
  z.xi = s->S.xi;
  z.l = s->S.l;
  z.lMH = 0.f;
  z.current_step = 1;
  z.da_t = s->S.da_t0;  // counter starts at t0 (dualaveragestepsize.jl:39)
  z.da_error_sum = 0.f;
  z.da_log_avg = 0.f;
  z.ma_t = 1;
  z.status = 0;
  z.mode_i = 1;
  z.converged = 0;
  z.bp1 = s->S.opt_rho;   // ADAM: βp = β  (Flux.Optimise.ADAM state)
  z.bp2 = s->S.opt_beta2;
  std::vector<ChainScalars> sc((size_t)s->C, z);
  CK(cudaMemcpyAsync(s->sc, sc.data(), sizeof(ChainScalars) * sc.size(), cudaMemcpyHostToDevice, m->stream));
  if (s->ssc) {
    StreamScalars zs{};
    zs.t = 1;
    zs.xi = s->S.xi;
    zs.l = s->S.l;
    zs.ma_t = 1;
    std::vector<StreamScalars> ssc((size_t)s->C, zs);
    CK(cudaMemcpyAsync(s->ssc, ssc.data(), sizeof(StreamScalars) * ssc.size(), cudaMemcpyHostToDevice, m->stream));
  }
  CK(cudaStreamSynchronize(m->stream));
  s->gstep = 0;
  s->nsampled = 0;
  *s->progress_host = 0;
  s->initialised = true;
  return BFX_OK;
}

int32_t bfx_sampler_init(bfx_sampler* s, const float* theta_start) { return sampler_init_impl(s, theta_start, true); }

static int32_t fetch_scalars(bfx_sampler* s, std::vector<ChainScalars>& sc) {
  sc.resize((size_t)s->C);
  CK(cudaSetDevice(s->m->device));
  CK(cudaMemcpy(sc.data(), s->sc, sizeof(ChainScalars) * sc.size(), cudaMemcpyDeviceToHost));
  if (s->ssc) {  // stream regime keeps its own scalars: mirror them
    std::vector<StreamScalars> ss((size_t)s->C);
    CK(cudaMemcpy(ss.data(), s->ssc, sizeof(StreamScalars) * ss.size(), cudaMemcpyDeviceToHost));
    for (int c = 0; c < s->C; ++c) {
      sc[c].t = ss[c].t;
      sc[c].nsampled = ss[c].nsampled;
      sc[c].xi = ss[c].xi;
      sc[c].l = ss[c].l;
      sc[c].status = ss[c].status;
      sc[c].ma_t = ss[c].ma_t;
    }
  }
  return BFX_OK;
}

static int32_t store_scalars(bfx_sampler* s, const std::vector<ChainScalars>& sc) {
  CK(cudaMemcpy(s->sc, sc.data(), sizeof(ChainScalars) * sc.size(), cudaMemcpyHostToDevice));
  if (s->ssc) {
    std::vector<StreamScalars> ss((size_t)s->C);
    CK(cudaMemcpy(ss.data(), s->ssc, sizeof(StreamScalars) * ss.size(), cudaMemcpyDeviceToHost));
    for (int c = 0; c < s->C; ++c) {
      ss[c].t = sc[c].t;
      ss[c].nsampled = sc[c].nsampled;
      ss[c].xi = sc[c].xi;
      ss[c].l = sc[c].l;
      ss[c].status = sc[c].status;
      ss[c].ma_t = sc[c].ma_t;
    }
    CK(cudaMemcpy(s->ssc, ss.data(), sizeof(StreamScalars) * ss.size(), cudaMemcpyHostToDevice));
  }
  return BFX_OK;
}

int32_t bfx_sampler_get_state(bfx_sampler* s, int32_t field, void* out, int64_t nbytes) {
  if (!s || !out) return fail(BFX_ERR_BAD_ARG, "NULL argument");
  if (!s->initialised) return fail(BFX_ERR_STATE, "sampler not initialised");
  bfx_model* m = s->m;
  int32_t st = ensure_stream(m);
  if (st) return st;
  const int P = m->M.P, C = s->C;
  const float* src = nullptr;
  switch (field) {
    case BFX_STATE_THETA: src = s->theta; break;
    case BFX_STATE_MOMENTUM: src = s->mom; break;
    case BFX_STATE_MINV: src = s->minv; break;
    default: break;
  }
  if (field == BFX_STATE_THETA || field == BFX_STATE_MOMENTUM || field == BFX_STATE_MINV) {
    if (nbytes != (int64_t)P * C * 4) return fail(BFX_ERR_BAD_ARG, "buffer must hold P x C floats");
    float* o = static_cast<float*>(out);
    if (!src) {  // state this sampler does not carry: momentum 0, Minv identity
      for (int64_t i = 0; i < (int64_t)P * C; ++i) o[i] = field == BFX_STATE_MINV ? 1.f : 0.f;
      return BFX_OK;
    }
    return state_d2h(s, src, o);
  }
  std::vector<ChainScalars> sc;
  st = fetch_scalars(s, sc);
  if (st) return st;
  if (field == BFX_STATE_XI || field == BFX_STATE_STEPSIZE) {
    if (nbytes != (int64_t)C * 4) return fail(BFX_ERR_BAD_ARG, "buffer must hold C floats");
    float* o = static_cast<float*>(out);
    for (int c = 0; c < C; ++c) o[c] = field == BFX_STATE_XI ? sc[c].xi : sc[c].l;
    return BFX_OK;
  }
  if (field == BFX_STATE_T || field == BFX_STATE_NSAMPLED) {
    if (nbytes != (int64_t)C * 8) return fail(BFX_ERR_BAD_ARG, "buffer must hold C int64");
    int64_t* o = static_cast<int64_t*>(out);
    for (int c = 0; c < C; ++c) o[c] = field == BFX_STATE_T ? sc[c].t : sc[c].nsampled;
    return BFX_OK;
  }
  if (field == BFX_STATE_STATUS || field == BFX_STATE_CONVERGED) {
    if (nbytes != (int64_t)C * 4) return fail(BFX_ERR_BAD_ARG, "buffer must hold C int32");
    int32_t* o = static_cast<int32_t*>(out);
    for (int c = 0; c < C; ++c) o[c] = field == BFX_STATE_STATUS ? sc[c].status : sc[c].converged;
    return BFX_OK;
  }
  return fail(BFX_ERR_BAD_ARG, "unknown state field");
}

int32_t bfx_sampler_set_state(bfx_sampler* s, int32_t field, const void* in, int64_t nbytes) {
  if (!s || !in) return fail(BFX_ERR_BAD_ARG, "NULL argument");
  if (!s->initialised) return fail(BFX_ERR_STATE, "sampler not initialised");
  bfx_model* m = s->m;
  int32_t st = ensure_stream(m);
  if (st) return st;
  const int P = m->M.P, C = s->C;
  if (field == BFX_STATE_THETA || field == BFX_STATE_MOMENTUM || field == BFX_STATE_MINV) {
    float* dst = field == BFX_STATE_THETA ? s->theta : field == BFX_STATE_MOMENTUM ? s->mom : s->minv;
    if (!dst) return fail(BFX_ERR_STATE, "this sampler does not carry that state");
    if (nbytes != (int64_t)P * C * 4) return fail(BFX_ERR_BAD_ARG, "buffer must hold P x C floats");
    st = state_h2d(s, dst, static_cast<const float*>(in), field == BFX_STATE_MINV ? 1.f : 0.f);
    if (st || field != BFX_STATE_THETA || m->stream_regime) return st;
    std::vector<ChainScalars> sc0;  // theta changed from outside: cached full-data log posteriors are stale
    st = fetch_scalars(s, sc0);
    if (st) return st;
    for (auto& c : sc0) c.lp_epoch = 0;
    return store_scalars(s, sc0);
  }
  std::vector<ChainScalars> sc;
  st = fetch_scalars(s, sc);
  if (st) return st;
  if (field == BFX_STATE_XI || field == BFX_STATE_STEPSIZE) {
    if (nbytes != (int64_t)C * 4) return fail(BFX_ERR_BAD_ARG, "buffer must hold C floats");
    const float* v = static_cast<const float*>(in);
    for (int c = 0; c < C; ++c) (field == BFX_STATE_XI ? sc[c].xi : sc[c].l) = v[c];
  } else if (field == BFX_STATE_T || field == BFX_STATE_NSAMPLED) {
    if (nbytes != (int64_t)C * 8) return fail(BFX_ERR_BAD_ARG, "buffer must hold C int64");
    const int64_t* v = static_cast<const int64_t*>(in);
    for (int c = 0; c < C; ++c) (field == BFX_STATE_T ? sc[c].t : sc[c].nsampled) = v[c];
    if (field == BFX_STATE_NSAMPLED) s->nsampled = v[0];
  } else if (field == BFX_STATE_STATUS) {
    if (nbytes != (int64_t)C * 4) return fail(BFX_ERR_BAD_ARG, "buffer must hold C int32");
    const int32_t* v = static_cast<const int32_t*>(in);
    for (int c = 0; c < C; ++c) sc[c].status = v[c];
  } else {
    return fail(BFX_ERR_BAD_ARG, "unknown state field");
  }
  return store_scalars(s, sc);
}

static int32_t fill_run_common(bfx_sampler* s, RunDev& R) {
  bfx_model* m = s->m;
  {
    float* scr;
    long long stride;
    int32_t st = ensure_scratch(m, (size_t)s->C, &scr, &stride);
    if (st) return st;
  }
  R.M = m->M;
  R.S = s->S;
  R.x = m->x;
  R.y = m->y;
  if (int32_t xst = ensure_xsplit(m)) return xst;
  R.xs = m->M.tc.enabled ? static_cast<const uint4*>(m->xsplit) : nullptr;
  R.n = m->n;
  R.data_epoch = m->data_epoch;
  R.C = s->C;
  R.chain0 = 0;
  R.Ps = s->Ps;
  R.theta = s->theta;
  R.mom = s->mom;
  R.minv = s->minv;
  R.theta_old = s->theta_old;
  R.rms_v = s->rms_v;
  R.ring = s->ring;
  R.mode_window = s->mode_window;
  R.sc = s->sc;
  R.keep_every = 1;
  R.progress = nullptr;
  R.scratch = m->scratch;
  R.scratch_stride = (long long)m->M.T * m->M.rec_rows * m->cfg.bt;
  return BFX_OK;
}

// one launch of the persistent (chain, chunk) kernel: reset the work counters, pick the chunk length
static cudaError_t launch_run(bfx_sampler* s, RunDev& R, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(s->work, 0, sizeof(int) * ((size_t)s->C + 1), st);
  if (e != cudaSuccess) return e;
  R.work = s->work;
  int cl = R.nsteps / 16;
  R.chunk_len = cl < 1 ? 1 : (cl > 16 ? 16 : cl);
  R.grid_cap = s->C;
  if (const char* cap = std::getenv("BFX_GRID_CAP")) {  // developer knob: size of the CTA pool (occupancy experiments)
    const long long v = std::atoll(cap);
    if (v > 0 && v < R.grid_cap) R.grid_cap = v;
  }
  return launch_mcmc(R, s->m->cfg, st);
}

static int ndraws_of(int kind) { return kind == BFX_GGMC ? 2 : 1; }

static int32_t status_after(bfx_sampler* s) {
  std::vector<ChainScalars> sc;
  int32_t st = fetch_scalars(s, sc);
  if (st) return st;
  // the host mirror follows the chains that are still running (a chain stops counting at its NaN guard)
  long long ns = -1;
  int bad = -1;
  for (int c = 0; c < s->C; ++c) {
    if (sc[c].status != 0) {
      if (bad < 0) bad = c;
    } else if (sc[c].nsampled > ns) {
      ns = sc[c].nsampled;
    }
  }
  s->nsampled = ns >= 0 ? ns : sc[0].nsampled;
  if (bad >= 0) {
    char buf[128];
    if (sc[bad].status == BFX_ERR_PEER)
      std::snprintf(buf, sizeof buf, "peer exchange timed out: a rank of the minibatch-sharded chain did not arrive");
    else
      std::snprintf(buf, sizeof buf, "NaN in %s (chain %d)", sc[bad].status == BFX_ERR_NAN_G ? "g" : "θ", bad);
    return fail(sc[bad].status, buf);
  }
  return BFX_OK;
}

// continue_sampling starts from the DEVICE state: bfx_mcmc_advance is asynchronous and never touches the host
// mirror of nsampled, and a run that ended in a NaN guard leaves it behind.  Synchronise, re-read the counters and
// refuse to continue while a chain is stopped (the lock-step sample bookkeeping would no longer hold).
static int32_t refresh_counters(bfx_sampler* s) {
  if (!s->initialised) return BFX_OK;
  CK(cudaSetDevice(s->device));
  if (s->adv_event) CK(cudaEventSynchronize(s->adv_event));
  if (s->m->stream) CK(cudaStreamSynchronize(s->m->stream));
  std::vector<ChainScalars> sc;
  int32_t st = fetch_scalars(s, sc);
  if (st) return st;
  for (int c = 0; c < s->C; ++c)
    if (sc[c].status != 0) {
      char buf[128];
      std::snprintf(buf, sizeof buf, "continue_sampling: chain %d stopped at a NaN guard (status %d); re-initialise the sampler",
                    c, sc[c].status);
      return fail(BFX_ERR_STATE, buf);
    }
  for (int c = 1; c < s->C; ++c)
    if (sc[c].nsampled != sc[0].nsampled) return fail(BFX_ERR_STATE, "continue_sampling: chains are not in lock step");
  s->nsampled = sc[0].nsampled;
  return BFX_OK;
}


// ---------------------------------------------------------------------------------
// stream regime: one update! of chain c (grad on the local minibatch, all-reduce of the exchange buffer when the
// chain is minibatch-sharded, update).  Everything is enqueued on `st`; nothing synchronises.
// ---------------------------------------------------------------------------------
static bool wsplit_valid(const bfx_sampler* s) {
  const bfx_model* m = s->m;
  return m->SM.tc && s->C == 1 && m->wsplit_owner == s && m->wsplit_gstep == s->gstep;
}

static int32_t stream_grad_local(bfx_sampler* s, int c, const StreamBatch& b, float nb, bool pack, cudaStream_t st,
                                 float* gb = nullptr, StreamBucketHook* hook = nullptr, const PeerDev* pd = nullptr) {
  if (!gb) gb = s->gbuf;
  bfx_model* m = s->m;
  int32_t rc = ensure_work(m, b.B);
  if (rc) return rc;
  const float* th = s->theta + (size_t)c * s->Ps;
  const bool valid = wsplit_valid(s);
  if (!valid) m->wsplit_owner = nullptr;
  // (PULL: waits until no peer reads the previous epoch's buffer; SCATTER: the join kernel of the previous update did)
  if (pd && pd->mode == PEER_MODE_PULL) CK(stream_zero_peer(gb, m->M.P, s->ssc + c, *pd, st));
  else CK(stream_zero(gb, m->M.P, s->ssc + c, st));
  CK(stream_eval_like(m->SM, m->W, th, m->x, m->y, b, nb, true, gb, s->ssc + c, st, nullptr, valid, hook));
  if (pack && pd) CK(stream_pack_sums_signal(gb, m->M.P, s->ssc + c, *pd, st));  // (sums + the peers' ready flags)
  else if (pack && !(hook && hook->used)) CK(stream_pack_sums(gb, m->M.P, s->ssc + c, st));
  return BFX_OK;
}

// where an update records: explicit device slots (injected steps, the sharded-chain entry points) or the run's
// buffers resolved on the device from the chain's counters
struct StreamRecord {
  float* sample_dev = nullptr;
  float* kinetic_dev = nullptr;
  float* samples_base = nullptr;
  long long keep_every = 1, kept_origin = 0, ring_slots = 0;
  float* kinetic_base = nullptr;
  long long hist_origin = 0, hist_stride = 0;
};

static int32_t stream_apply(bfx_sampler* s, int c, float nb, bool from_buffer, unsigned long long seed, unsigned chainkey,
                            const float* inj, const StreamRecord& rec, cudaStream_t st, float* gb = nullptr,
                            bool device_step = false, const PeerDev* pd = nullptr) {
  bfx_model* m = s->m;
  if (!gb) gb = s->gbuf;
  float* th = s->theta + (size_t)c * s->Ps;
  if (pd) {  // the exchange IS the finish pass: sum of all ranks' buffers read over NVLink, result in gsum
    gb = s->gsum;
    if (pd->mode == PEER_MODE_SCATTER) CK(stream_finish_scatter(m->SM, th, s->ssc + c, nb, *pd, st));
    else CK(stream_finish_peer(m->SM, th, gb, s->ssc + c, nb, *pd, st));
  } else {
    CK(stream_finish(m->SM, th, gb, s->ssc + c, nb, true, from_buffer, st));
  }
  StreamUpdate U{};
  U.S = &s->S;
  U.P = m->M.P;
  U.theta = th;
  U.mom = s->mom ? s->mom + (size_t)c * s->Ps : nullptr;
  U.minv = s->minv ? s->minv + (size_t)c * s->Ps : nullptr;
  U.rmsv = s->rms_v ? s->rms_v + (size_t)c * s->Ps : nullptr;
  U.g = gb;
  U.sc = s->ssc + c;
  U.seed = seed;
  U.chain = chainkey;
  U.gstep = device_step ? -1 : s->gstep;
  U.inj = inj;
  U.sample_out = rec.sample_dev;
  U.kinetic_out = rec.kinetic_dev;
  U.samples_base = rec.samples_base;
  U.keep_every = rec.keep_every; U.kept_origin = rec.kept_origin; U.ring_slots = rec.ring_slots;
  U.kinetic_base = rec.kinetic_base; U.hist_origin = rec.hist_origin; U.hist_stride = rec.hist_stride;
  const bool keep_split = m->SM.tc && s->C == 1 && m->W.w_hi;
  U.w_hi = keep_split ? m->W.w_hi : nullptr;
  U.w_lo = keep_split ? m->W.w_lo : nullptr;
  cudaError_t e = stream_update(U, st);
  if (e != cudaSuccess) return fail(BFX_ERR_CUDA, std::string("stream_update: ") + cudaGetErrorString(e));
  if (keep_split) {
    m->wsplit_owner = s;
    m->wsplit_gstep = s->gstep + 1;  // valid for the evaluation of the next update (the caller advances gstep)
  }
  return BFX_OK;
}

// the all-reduce of the exchange buffer [g ; sum1 ; sum2 ; B] of a minibatch-sharded chain
static int32_t stream_allreduce(bfx_sampler* s, float* gb, cudaStream_t st, long long count = -1) {
  NcclApi* api = nccl_api();
  if (!api || !s->comm) return fail(BFX_ERR_STATE, "minibatch-sharded update without a communicator");
  if (count < 0) count = (long long)s->m->M.P + 3;
  const int r = api->AllReduce(gb, gb, (size_t)count, kNcclFloat32, kNcclSum, s->comm, st);
  if (r != 0) return fail(BFX_ERR_CUDA, std::string("ncclAllReduce: ") + (api->GetErrorString ? api->GetErrorString(r) : "error"));
  return BFX_OK;
}
static cudaError_t hook_reduce(void* ctx, float* ptr, long long count, cudaStream_t st) {
  return stream_allreduce(static_cast<bfx_sampler*>(ctx), ptr, st, count) == BFX_OK ? cudaSuccess : cudaErrorUnknown;
}

// one whole update! of chain c, scheduled batch
static int32_t stream_update_once(bfx_sampler* s, const bfx_run_desc* run, int c, float nbat, const StreamRecord& rec,
                                  bool device_step, cudaStream_t st) {
  bfx_model* m = s->m;
  const bool sharded = run->shard_mode == BFX_SHARD_MINIBATCH && s->comm != nullptr;
  const unsigned permkey = (unsigned)(run->chain_offset + c);
  // a sharded chain draws its local batch with its own permutation stream (chain_offset = rank) and applies the
  // update with the shared noise key 0: identical Philox streams keep theta / p / xi bit-identical on every rank
  const unsigned noisekey = sharded ? 0u : permkey;
  StreamBatch b = stream_batch(m, run, permkey, s->gstep, device_step);
  // Default: ONE all-reduce of the P + 3 floats behind the backward pass.  BFX_BUCKETS=1 issues per-layer buckets on
  // the side stream while the backward pass continues instead: measured slower since the GEMM grids fill all 148 SMs
  // in one wave (the NCCL kernels take SMs from them: 0.245 against 0.216 ms per update on 2 B200, profiles/).
  StreamBucketHook* hook = sharded && s->comm_stream && std::getenv("BFX_BUCKETS") ? &s->hook : nullptr;
  // peer-memory exchange (all ranks opened all peers at comm_init; BFX_NO_P2P=1 keeps the NCCL call)
  const PeerDev* pd = sharded && s->peer.n > 1 && !hook && !std::getenv("BFX_NO_P2P") ? &s->peer : nullptr;
  int32_t rc = stream_grad_local(s, c, b, nbat, sharded, st, nullptr, hook, pd);
  if (rc) return rc;
  if (sharded) {
    if (pd) {
      // (the ready flags went out with the likelihood sums, stream_grad_local)
    } else if (hook && hook->used) {
      CK(cudaStreamWaitEvent(st, hook->ev_join, 0));
    } else {
      rc = stream_allreduce(s, s->gbuf, st);
      if (rc) return rc;
    }
  }
  return stream_apply(s, c, nbat, sharded, run->seed, noisekey, nullptr, rec, st, nullptr, device_step, pd);
}

static int32_t stream_run(bfx_sampler* s, const bfx_run_desc* run, int64_t nsteps, int64_t base, float* samples_host,
                          int64_t nkept, float* kinetic_host, int64_t nnew, float* samples_dev, int64_t ring_slots,
                          cudaStream_t st) {
  bfx_model* m = s->m;
  const long long P = m->M.P;
  const int C = s->C;
  const int64_t ke = run->keep_every > 0 ? run->keep_every : 1;
  long long B = run->batchsize < m->n ? run->batchsize : m->n;
  long long nbat = run->partial ? (m->n + B - 1) / B : m->n / B;
  if (nbat < 1) nbat = 1;
  if (run->shard_mode == BFX_SHARD_MINIBATCH && !s->comm && run->ngpus > 1)
    return fail(BFX_ERR_STATE, "shard_mode MINIBATCH needs bfx_sampler_comm_init on every rank");
  if (run->shard_mode == BFX_SHARD_MINIBATCH && C != 1)
    return fail(BFX_ERR_BAD_ARG, "a minibatch-sharded sampler carries exactly one chain");
  float* d_kin = nullptr;
  if (kinetic_host) {
    CK(cudaMalloc(&d_kin, (size_t)C * nnew * 4));
    CK(cudaMemsetAsync(d_kin, 0, (size_t)C * nnew * 4, st));
  }
  const int64_t slots = ring_slots > 0 ? ring_slots : 1;
  auto record_of = [&](int c) {
    StreamRecord rec;
    if (samples_dev) {
      rec.samples_base = samples_dev + (size_t)c * slots * P;
      rec.keep_every = ke;
      rec.kept_origin = 0;
      rec.ring_slots = slots;
    }
    if (d_kin) {
      rec.kinetic_base = d_kin + (size_t)c * nnew;
      rec.hist_origin = base;
      rec.hist_stride = nnew;
    }
    return rec;
  };
  int32_t rc = ensure_work(m, B);
  if (rc) return rc;
  // One chain: the update is a fixed launch sequence whose per-step quantities are resolved on the device, so it is
  // captured once as a CUDA graph and replayed; only the short last batch of an epoch is launched directly.
  const bool use_graph = C == 1 && !std::getenv("BFX_NO_GRAPH");
  if (use_graph) {
    if (m->SM.tc && !wsplit_valid(s)) {  // the graph assumes current BF16 halves of theta
      CK(stream_split_theta(m->SM, m->W, s->theta, st));
      m->wsplit_owner = s;
      m->wsplit_gstep = s->gstep;
    }
    CK(stream_set_gstep(s->ssc, s->gstep, st));
    struct Key {
      long long B, nbat, n, ke, slots, base, nnew;
      int shuffle, batch_mode, shard, tc, nobuckets, pad;
      unsigned long long seed;
      long long chain_offset;
      const void *samples_dev, *d_kin, *comm, *x, *w;
    } key{B, nbat, m->n, ke, samples_dev ? slots : 0, d_kin ? base : 0, d_kin ? nnew : 0, run->shuffle, run->batch_mode,
          run->shard_mode, m->SM.tc, std::getenv("BFX_BUCKETS") ? 0 : 1, std::getenv("BFX_NO_P2P") ? 0 : s->peer.n * 4 + s->peer.mode, run->seed, run->chain_offset, samples_dev, d_kin,
          s->comm, m->x, m->W.w_hi};
    std::vector<unsigned char> kb(sizeof key);
    std::memcpy(kb.data(), &key, sizeof key);
    if (!s->upd_graph || kb != s->upd_graph_key) {
      if (s->upd_graph) {
        cudaGraphExecDestroy(s->upd_graph);
        s->upd_graph = nullptr;
      }
      // capture with a full-size batch (step index irrelevant: device_step)
      const long long g_save = s->gstep;
      s->gstep = 0;
      const void* own_save = m->wsplit_owner;
      const long long wg_save = m->wsplit_gstep;
      m->wsplit_owner = s;
      m->wsplit_gstep = 0;
      cudaGraph_t graph = nullptr;
      CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
      rc = stream_update_once(s, run, 0, (float)nbat, record_of(0), true, st);
      cudaError_t ce = cudaStreamEndCapture(st, &graph);
      s->gstep = g_save;
      m->wsplit_owner = own_save;
      m->wsplit_gstep = wg_save;
      if (rc) {
        if (graph) cudaGraphDestroy(graph);
        return rc;
      }
      if (ce != cudaSuccess) return fail(BFX_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(ce));
      ce = cudaGraphInstantiate(&s->upd_graph, graph, 0);
      cudaGraphDestroy(graph);
      if (ce != cudaSuccess) return fail(BFX_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce));
      s->upd_graph_key = kb;
    }
  }
  for (int64_t it = 0; it < nsteps; ++it) {
    const int64_t ns = base + it + 1;  // nsampled after this update (SGLD / SGNHT / SGNHTS: one sample per update)
    if (use_graph) {
      const long long k = s->gstep % nbat;
      const bool full_batch = m->n - k * B >= B;
      if (full_batch) {
        CK(cudaGraphLaunch(s->upd_graph, st));
        s->graph_replays += 1;
        if (m->SM.tc) {
          m->wsplit_owner = s;
          m->wsplit_gstep = s->gstep + 1;
        }
      } else {
        rc = stream_update_once(s, run, 0, (float)nbat, record_of(0), false, st);
        if (rc) return rc;
      }
      if (samples_host && ns % ke == 0) {
        const int64_t col = ns / ke - 1 - base / ke;
        CK(cudaMemcpyAsync(samples_host + (size_t)col * P, s->theta, (size_t)P * 4, cudaMemcpyDeviceToHost, st));
      }
    } else {
      for (int c = 0; c < C; ++c) {
        StreamRecord rec = record_of(c);
        rc = stream_update_once(s, run, c, (float)nbat, rec, false, st);
        if (rc) return rc;
        if (samples_host && ns % ke == 0) {
          const int64_t col = ns / ke - 1 - base / ke;
          CK(cudaMemcpyAsync(samples_host + ((size_t)c * nkept + col) * P, s->theta + (size_t)c * s->Ps, (size_t)P * 4,
                             cudaMemcpyDeviceToHost, st));
        }
      }
    }
    s->gstep += 1;
    s->launches += use_graph ? 1 : 16;
    if ((it & 63) == 63) CK(cudaStreamSynchronize(st));  // bound the launch queue; also paces the progress counter
    *s->progress_host = (unsigned long long)s->gstep;
  }
  if (d_kin) {
    CK(cudaMemcpyAsync(kinetic_host, d_kin, (size_t)C * nnew * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    cudaFree(d_kin);
    if (s->upd_graph) {  // the graph recorded into d_kin
      cudaGraphExecDestroy(s->upd_graph);
      s->upd_graph = nullptr;
    }
  }
  return BFX_OK;
}

int32_t bfx_sampler_step_injected(bfx_sampler* s, const int64_t* idx, int64_t B, int64_t stride, float num_batches,
                                  const float* normals, const float* uniforms, float* theta_out) {
  if (!s || !idx || !normals) return fail(BFX_ERR_BAD_ARG, "NULL argument");
  if (!s->initialised) return fail(BFX_ERR_STATE, "sampler not initialised");
  bfx_model* m = s->m;
  if (!m->x) return fail(BFX_ERR_NO_DATA, "bfx_model_set_data has not been called");
  if (B <= 0) return fail(BFX_ERR_BAD_ARG, "B must be > 0");
  const bool mh = s->S.kind == S_GGMC || s->S.kind == S_HMC;
  if (mh && !uniforms) return fail(BFX_ERR_BAD_ARG, "uniforms required for GGMC/HMC");
  int32_t st = ensure_stream(m);
  if (st) return st;
  const int P = m->M.P, C = s->C;
  const int64_t nidx = stride ? (int64_t)(C - 1) * stride + B : B;
  for (int64_t i = 0; i < nidx; ++i)
    if (idx[i] < 0 || idx[i] >= m->n) return fail(BFX_ERR_BAD_ARG, "batch index out of range");
  const size_t nn = (size_t)ndraws_of(s->S.kind) * P * C;
  float *d_n = nullptr, *d_u = nullptr;
  long long* d_idx = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_n);
    cudaFree(d_u);
    cudaFree(d_idx);
  };
  CKC(cudaMalloc(&d_n, nn * 4));
  CKC(cudaMalloc(&d_idx, (size_t)nidx * 8));
  CKC(cudaMemcpyAsync(d_n, normals, nn * 4, cudaMemcpyHostToDevice, m->stream));
  CKC(cudaMemcpyAsync(d_idx, idx, (size_t)nidx * 8, cudaMemcpyHostToDevice, m->stream));
  if (uniforms) {
    CKC(cudaMalloc(&d_u, (size_t)C * 4));
    CKC(cudaMemcpyAsync(d_u, uniforms, (size_t)C * 4, cudaMemcpyHostToDevice, m->stream));
  }
  if (m->stream_regime) {
    for (int c = 0; c < C; ++c) {
      StreamBatch b{};
      b.n = m->n;
      b.idx = d_idx + (long long)c * stride;
      b.B = B;
      int32_t rc = stream_grad_local(s, c, b, num_batches, false, m->stream);
      if (!rc) rc = stream_apply(s, c, num_batches, false, 0, (unsigned)c, d_n + (size_t)c * P, StreamRecord{}, m->stream);
      if (rc) {
        cleanup();
        return rc;
      }
    }
    s->gstep += 1;
    s->launches += 12;
    CKC(cudaStreamSynchronize(m->stream));
    if (theta_out) {
      std::vector<float> pad((size_t)s->Ps * C);
      CKC(cudaMemcpy(pad.data(), s->theta, pad.size() * 4, cudaMemcpyDeviceToHost));
      padded_to_flat(m, pad, C, theta_out);
    }
    cleanup();
    return status_after(s);
  }
  RunDev R{};
  {
    int32_t fst = fill_run_common(s, R);
    if (fst) {
      cleanup();
      return fst;
    }
  }
  R.B = B;
  R.nb = 1;
  R.num_batches = num_batches;
  R.shuffle = 0;
  R.idx = d_idx;
  R.idx_chain_stride = stride;
  R.seed = 0;
  R.gstep0 = s->gstep;
  R.nsteps = 1;
  R.inj_normals = d_n;
  R.inj_uniforms = d_u;
  CKC(launch_run(s, R, m->stream));
  s->launches += 1;
  s->gstep += 1;
  CKC(cudaStreamSynchronize(m->stream));
  if (theta_out) {
    std::vector<float> pad((size_t)s->Ps * C);
    CKC(cudaMemcpy(pad.data(), s->theta, pad.size() * 4, cudaMemcpyDeviceToHost));
    padded_to_flat(m, pad, C, theta_out);
  }
  cleanup();
  return status_after(s);
}

static int32_t run_counts(const bfx_sampler* s, const bfx_run_desc* run, int64_t* nnew, int64_t* nkept) {
  if (!s || !run) return fail(BFX_ERR_BAD_ARG, "NULL argument");
  if (run->numsamples < 0 || run->batchsize <= 0) return fail(BFX_ERR_BAD_ARG, "bad batchsize / numsamples");
  const int64_t ke = run->keep_every > 0 ? run->keep_every : 1;
  const int64_t base = run->continue_sampling ? s->nsampled : 0;
  if (run->continue_sampling && !s->initialised) return fail(BFX_ERR_STATE, "continue_sampling needs a previous run");
  int64_t nn = run->numsamples - base;
  if (nn < 0) nn = 0;
  if (nnew) *nnew = nn;
  if (nkept) *nkept = (base + nn) / ke - base / ke;
  return BFX_OK;
}

int32_t bfx_mcmc_num_kept(const bfx_sampler* s, const bfx_run_desc* run, int64_t* nnew, int64_t* nkept) {
  if (s && run && run->continue_sampling) {
    int32_t st = refresh_counters(const_cast<bfx_sampler*>(s));  // the mirror is a cache of device state
    if (st) return st;
  }
  return run_counts(s, run, nnew, nkept);
}

static int32_t prepare_schedule(bfx_sampler* s, const bfx_run_desc* run, RunDev& R) {
  bfx_model* m = s->m;
  if (!m->x) return fail(BFX_ERR_NO_DATA, "bfx_model_set_data has not been called");
  {
    int32_t fst = fill_run_common(s, R);
    if (fst) return fst;
  }
  R.B = run->batchsize < m->n ? run->batchsize : m->n;
  R.nb = run->partial ? (m->n + R.B - 1) / R.B : m->n / R.B;  // length(DataLoader)
  if (R.nb < 1) R.nb = 1;
  R.num_batches = (float)R.nb;
  R.shuffle = run->shuffle ? 1 : 0;
  R.batch_shared = run->batch_mode == BFX_BATCH_SHARED;
  R.idx = nullptr;
  R.seed = run->seed;
  R.chain0 = (unsigned)run->chain_offset;
  R.keep_every = run->keep_every > 0 ? run->keep_every : 1;
  return BFX_OK;
}

int32_t bfx_mcmc_run(bfx_sampler* s, const bfx_run_desc* run, const float* theta_start, float* samples_out,
                     float* kinetic_out, int32_t* accepted_out) {
  int64_t nnew = 0, nkept = 0;
  if (!s || !run) return fail(BFX_ERR_BAD_ARG, "NULL argument");
  int32_t st = BFX_OK;
  if (run->continue_sampling) {
    st = refresh_counters(s);
    if (st) return st;
  }
  st = run_counts(s, run, &nnew, &nkept);
  if (st) return st;
  bfx_model* m = s->m;
  st = ensure_stream(m);
  if (st) return st;
  // Chain-group pipelining (fresh runs that fit one launch): theta of group g+1 goes up and the samples of group
  // g-1 come down while the kernel of group g runs, instead of copy -> kernel -> copy for all chains at once.
  int groups = 1;
  if (!run->continue_sampling && samples_out && nkept > 0 && !m->stream_regime && s->C >= 512) {
    const int64_t ke0 = run->keep_every > 0 ? run->keep_every : 1;
    const int64_t ups0 = s->S.kind == S_HMC ? s->S.path_len : 1;
    const int64_t bytes_all = (int64_t)s->C * m->M.P * 4 * (nnew / ke0 + 1);
    // a group must still fill the persistent CTA pool (chunks of one chain run in order, so chains per launch, not
    // work items, bound the parallelism): at most 4 groups of >= 320 chains
    if (nnew * ups0 <= (1 << 16) && bytes_all <= (int64_t)(1ll << 30)) groups = std::max(1, std::min(4, s->C / 320));
    if (const char* g = std::getenv("BFX_PIPELINE_GROUPS")) groups = std::max(1, std::min(16, std::atoi(g)));
    if (groups > 1 && !(nnew * ups0 <= (1 << 16) && bytes_all <= (int64_t)(1ll << 30))) groups = 1;
  }
  if (!run->continue_sampling) {
    if (!theta_start) return fail(BFX_ERR_BAD_ARG, "theta_start is NULL");
    st = sampler_init_impl(s, theta_start, groups == 1);
    if (st) return st;
  }
  if (nnew == 0) {
    if (groups > 1) {  // nothing to run, but theta must still arrive
      st = state_h2d(s, s->theta, theta_start, 0.f);
      if (st) return st;
    }
    return BFX_OK;
  }
  if (run->continue_sampling) {
    // initialise!(...; continue_sampling=true) re-zeroes the momentum (sgnht.jl:54, sgnht-s.jl:72,
    // ggmc.jl:122, hmc.jl:92-93) and GGMC restarts its window (ggmc.jl:116); everything else is kept
    if (s->mom) CK(cudaMemsetAsync(s->mom, 0, (size_t)s->Ps * sizeof(float) * (size_t)s->C, m->stream));
    if (s->S.kind == S_GGMC) {
      std::vector<ChainScalars> sc;
      st = fetch_scalars(s, sc);
      if (st) return st;
      for (auto& c : sc) c.current_step = 1;
      st = store_scalars(s, sc);
      if (st) return st;
    }
  }
  if (m->stream_regime) {
    if (!m->x) return fail(BFX_ERR_NO_DATA, "bfx_model_set_data has not been called");
    st = stream_run(s, run, nnew, s->nsampled, samples_out, nkept, kinetic_out, nnew, nullptr, 0, m->stream);
    if (st) return st;
    CK(cudaStreamSynchronize(m->stream));
    return status_after(s);
  }
  RunDev R{};
  st = prepare_schedule(s, run, R);
  if (st) return st;
  const int P = m->M.P, C = s->C;
  const int64_t ke = R.keep_every;
  const int64_t ups = s->S.kind == S_HMC ? s->S.path_len : 1;  // update! calls per sample
  const int64_t win = s->S.kind == S_GGMC ? s->S.steps : 1;    // samples that may be rewritten together
  const int64_t base = s->nsampled;
  // chunking: device sample buffers of <= ~1 GiB each, double buffered
  int64_t chunk_samples = nnew;
  const int64_t bytes_per_kept = (int64_t)C * P * 4;
  if (samples_out) {
    int64_t max_kept = (int64_t)(1ll << 30) / bytes_per_kept;
    if (max_kept < 1) max_kept = 1;
    chunk_samples = max_kept * ke;
  }
  const int64_t max_steps_per_launch = 1 << 16;
  if (chunk_samples * ups > max_steps_per_launch) chunk_samples = max_steps_per_launch / ups;
  chunk_samples = chunk_samples / win * win;
  if (chunk_samples < win) chunk_samples = win;
  float* d_samp[2] = {nullptr, nullptr};
  float* d_kin = nullptr;
  int* d_acc = nullptr;
  cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
  std::vector<cudaEvent_t> ev_up, ev_grp;
  bool registered = false;
  auto cleanup = [&]() {
    cudaFree(d_kin);
    cudaFree(d_acc);
    for (cudaEvent_t e : ev_up) cudaEventDestroy(e);
    for (cudaEvent_t e : ev_grp) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) {
      if (ev_done[i]) cudaEventDestroy(ev_done[i]);
      if (ev_copied[i]) cudaEventDestroy(ev_copied[i]);
    }
    if (registered) cudaHostUnregister(samples_out);
  };
  int64_t slots_per_chunk = 0;
  if (samples_out && nkept > 0) {
    slots_per_chunk = chunk_samples / ke + 1;
    const int64_t need_slots = std::min<int64_t>(slots_per_chunk, nkept + 1);
    const size_t need = (size_t)need_slots * bytes_per_kept;
    if (need > s->samp_cap) {  // the device chunks are kept between calls (grow only)
      for (int i = 0; i < 2; ++i) {
        cudaFree(s->samp_cache[i]);
        s->samp_cache[i] = nullptr;
      }
      s->samp_cap = 0;
      for (int i = 0; i < 2; ++i) CKC(cudaMalloc(&s->samp_cache[i], need));
      s->samp_cap = need;
    }
    for (int i = 0; i < 2; ++i) {
      d_samp[i] = s->samp_cache[i];
      CKC(cudaEventCreateWithFlags(&ev_done[i], cudaEventDisableTiming));
      CKC(cudaEventCreateWithFlags(&ev_copied[i], cudaEventDisableTiming));
    }
    // pinning pays for long streams only: registering costs milliseconds
    if ((size_t)nkept * bytes_per_kept >= ((size_t)256 << 20)) {
      registered = cudaHostRegister(samples_out, (size_t)nkept * bytes_per_kept, cudaHostRegisterDefault) == cudaSuccess;
      if (!registered) cudaGetLastError();
    }
  }
  if (kinetic_out && (s->S.kind == S_SGNHT || s->S.kind == S_SGNHTS)) {
    CKC(cudaMalloc(&d_kin, (size_t)C * nnew * 4));
    CKC(cudaMemsetAsync(d_kin, 0, (size_t)C * nnew * 4, m->stream));
  }
  if (accepted_out && (s->S.kind == S_GGMC || s->S.kind == S_HMC)) {
    CKC(cudaMalloc(&d_acc, (size_t)C * nnew * 4));
    CKC(cudaMemsetAsync(d_acc, 0, (size_t)C * nnew * 4, m->stream));
  }
  R.kinetic_out = d_kin;
  R.accepted_out = d_acc;
  R.hist_stride = nnew;
  R.hist_origin = base;
  R.progress = s->progress_dev;
  int64_t done = 0;
  int buf = 0;
  int64_t kept_written = 0;
  while (done < nnew) {
    const int64_t cs = std::min(chunk_samples, nnew - done);
    const int64_t a = base + done, b = a + cs;
    const int64_t slots = b / ke - a / ke;
    R.gstep0 = s->gstep;
    R.nsteps = (int)(cs * ups);
    if (groups > 1 && !(d_samp[0] && slots > 0 && cs == nnew)) {
      // (window rounding left it short of one launch after all) bring theta up the plain way
      int32_t hst = state_h2d(s, s->theta, theta_start, 0.f);
      if (hst) {
        cleanup();
        return hst;
      }
      groups = 1;
    }
    if (groups > 1) {
      R.samples = d_samp[0];
      R.samples_chain_stride = slots * P;
      R.kept_origin = a / ke;
      const int Cg = (C + groups - 1) / groups;
      ev_up.resize(groups, nullptr);
      ev_grp.resize(groups, nullptr);
      for (int g = 0; g < groups; ++g) {
        CKC(cudaEventCreateWithFlags(&ev_up[g], cudaEventDisableTiming));
        CKC(cudaEventCreateWithFlags(&ev_grp[g], cudaEventDisableTiming));
      }
      auto copy_down = [&](int g) -> cudaError_t {
        const int c0 = g * Cg, cn = std::min(Cg, C - c0);
        if (cn <= 0) return cudaSuccess;
        cudaError_t e = cudaStreamWaitEvent(s->copy_stream, ev_grp[g], 0);
        if (e != cudaSuccess) return e;
        return cudaMemcpy2DAsync(samples_out + ((size_t)c0 * nkept + kept_written) * P, (size_t)nkept * P * 4,
                                 d_samp[0] + (size_t)c0 * slots * P, (size_t)slots * P * 4, (size_t)slots * P * 4, cn,
                                 cudaMemcpyDeviceToHost, s->copy_stream);
      };
      for (int g = 0; g < groups; ++g) {
        const int c0 = g * Cg, cn = std::min(Cg, C - c0);
        if (cn <= 0) break;
        CKC(cudaMemcpyAsync(s->stage + (size_t)c0 * P, theta_start + (size_t)c0 * P, (size_t)cn * P * 4,
                            cudaMemcpyHostToDevice, s->copy_stream));
        CKC(cudaEventRecord(ev_up[g], s->copy_stream));
        CKC(cudaStreamWaitEvent(m->stream, ev_up[g], 0));
        CKC(launch_flat_to_padded(s->stage + (size_t)c0 * P, s->theta + (size_t)c0 * s->Ps, m->d_flat_pad, P, s->Ps, cn,
                                  0.f, m->stream));
        R.c_begin = c0;
        R.c_count = cn;
        CKC(launch_run(s, R, m->stream));
        s->launches += 1;
        CKC(cudaEventRecord(ev_grp[g], m->stream));
        if (g > 0) CKC(copy_down(g - 1));  // blocks the host on pageable memory while group g computes
      }
      CKC(copy_down(groups - 1));
      R.c_begin = 0;
      R.c_count = 0;
      s->gstep += R.nsteps;
      kept_written += slots;
      done += cs;
      continue;
    }
    if (d_samp[buf] && slots > 0) {
      CKC(cudaStreamWaitEvent(m->stream, ev_copied[buf], 0));  // buffer free again?
      R.samples = d_samp[buf];
      R.samples_chain_stride = slots * P;
      R.kept_origin = a / ke;
    } else {
      R.samples = nullptr;
    }
    CKC(launch_run(s, R, m->stream));
    s->launches += 1;
    s->gstep += R.nsteps;
    if (R.samples) {
      CKC(cudaEventRecord(ev_done[buf], m->stream));
      CKC(cudaStreamWaitEvent(s->copy_stream, ev_done[buf], 0));
      // device [C][slots][P]  ->  host [C][nkept][P] at column kept_written
      CKC(cudaMemcpy2DAsync(samples_out + kept_written * P, (size_t)nkept * P * 4, d_samp[buf], (size_t)slots * P * 4,
                            (size_t)slots * P * 4, C, cudaMemcpyDeviceToHost, s->copy_stream));
      CKC(cudaEventRecord(ev_copied[buf], s->copy_stream));
      kept_written += slots;
      buf ^= 1;
    }
    done += cs;
  }
  CKC(cudaStreamSynchronize(m->stream));
  CKC(cudaStreamSynchronize(s->copy_stream));
  if (d_kin) {
    // device [C][nnew] -> host nnew x C column-major == the same bytes
    CKC(cudaMemcpy(kinetic_out, d_kin, (size_t)C * nnew * 4, cudaMemcpyDeviceToHost));
  }
  if (d_acc) CKC(cudaMemcpy(accepted_out, d_acc, (size_t)C * nnew * 4, cudaMemcpyDeviceToHost));
  cleanup();
  return status_after(s);
}

int32_t bfx_progress(const bfx_sampler* s, int64_t* nsampled) {
  if (!s || !nsampled) return fail(BFX_ERR_BAD_ARG, "NULL argument");
  *nsampled = (int64_t) * (volatile unsigned long long*)s->progress_host;
  return BFX_OK;
}

static int32_t mark_advance(bfx_sampler* s, cudaStream_t st) {
  if (!s->adv_event) CK(cudaEventCreateWithFlags(&s->adv_event, cudaEventDisableTiming));
  CK(cudaEventRecord(s->adv_event, st));
  return BFX_OK;
}

int32_t bfx_mcmc_advance(bfx_sampler* s, const bfx_run_desc* run, int64_t nsteps, void* stream_handle,
                         float* samples_dev, int64_t ring_slots) {
  if (!s || !run) return fail(BFX_ERR_BAD_ARG, "NULL argument");
  if (!s->initialised) return fail(BFX_ERR_STATE, "sampler not initialised");
  if (nsteps <= 0) return BFX_OK;
  bfx_model* m = s->m;
  int32_t st = ensure_stream(m);
  if (st) return st;
  if (m->stream_regime) {
    if (!m->x) return fail(BFX_ERR_NO_DATA, "bfx_model_set_data has not been called");
    cudaStream_t sst = stream_handle ? (cudaStream_t)stream_handle : m->stream;
    st = stream_run(s, run, nsteps, s->gstep, nullptr, 0, nullptr, 0, samples_dev, ring_slots, sst);
    if (st) return st;
    return mark_advance(s, sst);
  }
  RunDev R{};
  st = prepare_schedule(s, run, R);
  if (st) return st;
  cudaStream_t stream = stream_handle ? (cudaStream_t)stream_handle : m->stream;
  if (samples_dev) {
    if (ring_slots < (s->S.kind == S_GGMC ? s->S.steps : 1))
      return fail(BFX_ERR_BAD_ARG, "ring_slots must be >= 1 (>= steps for GGMC)");
    R.samples = samples_dev;
    R.samples_chain_stride = ring_slots * (int64_t)m->M.P;
    R.kept_origin = 0;
    R.ring_slots = ring_slots;
  }
  int64_t done = 0;
  while (done < nsteps) {
    const int64_t cs = std::min<int64_t>(nsteps - done, 1 << 16);
    R.gstep0 = s->gstep;
    R.nsteps = (int)cs;
    CK(launch_run(s, R, stream));
    s->launches += 1;
    s->gstep += cs;
    done += cs;
  }
  return mark_advance(s, stream);
}

// Page-lock / unlock a caller-owned host buffer that is reused over many bfx_mcmc_run calls: copies to and from
// pinned memory are plain DMA transfers (no staging through the driver's bounce buffer, no host memcpy), which is
// what keeps several processes on one box from competing for host memory bandwidth.
int32_t bfx_host_register(void* ptr, size_t bytes) {
  if (!ptr || !bytes) return fail(BFX_ERR_BAD_ARG, "NULL argument");
  cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
  if (e == cudaErrorHostMemoryAlreadyRegistered) {
    cudaGetLastError();
    return BFX_OK;
  }
  if (e != cudaSuccess) return fail(BFX_ERR_CUDA, std::string("cudaHostRegister: ") + cudaGetErrorString(e));
  return BFX_OK;
}

int32_t bfx_host_unregister(void* ptr) {
  if (!ptr) return fail(BFX_ERR_BAD_ARG, "NULL argument");
  cudaError_t e = cudaHostUnregister(ptr);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(BFX_ERR_CUDA, std::string("cudaHostUnregister: ") + cudaGetErrorString(e));
  }
  return BFX_OK;
}

int32_t bfx_sampler_launch_count(const bfx_sampler* s, int64_t* n) {
  if (!s || !n) return fail(BFX_ERR_BAD_ARG, "NULL argument");
  *n = s->launches;
  return BFX_OK;
}

// ---- minibatch-sharded single chain (SURVEY.md section 8e, configs[2]) -------------------------------
int32_t bfx_stream_grad_local(bfx_sampler* s, const bfx_run_desc* run, float num_batches, void* stream_handle,
                              float* buf_dev, int64_t count) {
  if (!s || !run || !buf_dev) return fail(BFX_ERR_BAD_ARG, "NULL argument");
  if (count != (int64_t)s->m->M.P + 3) return fail(BFX_ERR_BAD_ARG, "the exchange buffer must hold P + 3 floats");
  bfx_model* m = s->m;
  if (!m->stream_regime) return fail(BFX_ERR_UNSUPPORTED, "bfx_stream_*: the model is in the shared-memory regime");
  if (s->C != 1) return fail(BFX_ERR_BAD_ARG, "a minibatch-sharded sampler carries exactly one chain");
  if (!s->initialised) return fail(BFX_ERR_STATE, "sampler not initialised");
  if (!m->x) return fail(BFX_ERR_NO_DATA, "bfx_model_set_data has not been called");
  int32_t st = ensure_stream(m);
  if (st) return st;
  cudaStream_t cs = stream_handle ? (cudaStream_t)stream_handle : m->stream;
  // the local minibatch of this update: this rank's slice of the schedule over ITS shard of the data
  StreamBatch b = stream_batch(m, run, (unsigned)run->chain_offset, s->gstep);
  return stream_grad_local(s, 0, b, num_batches, true, cs, buf_dev);
}

int32_t bfx_stream_apply_update(bfx_sampler* s, const bfx_run_desc* run, float num_batches, void* stream_handle,
                                float* buf_dev, float* sample_dev) {
  if (!s || !run || !buf_dev) return fail(BFX_ERR_BAD_ARG, "NULL argument");
  bfx_model* m = s->m;
  if (!m->stream_regime) return fail(BFX_ERR_UNSUPPORTED, "bfx_stream_*: the model is in the shared-memory regime");
  if (s->C != 1) return fail(BFX_ERR_BAD_ARG, "a minibatch-sharded sampler carries exactly one chain");
  cudaStream_t cs = stream_handle ? (cudaStream_t)stream_handle : m->stream;
  // chain key 0 on every rank: identical Philox streams keep the replicated theta / p / xi bit-identical
  StreamRecord rec;
  rec.sample_dev = sample_dev;
  int32_t st = stream_apply(s, 0, num_batches, true, run->seed, 0u, nullptr, rec, cs, buf_dev);
  if (st) return st;
  s->gstep += 1;
  s->launches += 16;
  return BFX_OK;
}

int32_t bfx_stream_allreduce(bfx_sampler* s, void* stream_handle, float* buf_dev) {
  if (!s || !buf_dev) return fail(BFX_ERR_BAD_ARG, "NULL argument");
  if (!s->comm) return BFX_OK;  // a single rank: nothing to exchange
  int32_t st = ensure_stream(s->m);
  if (st) return st;
  cudaStream_t cs = stream_handle ? (cudaStream_t)stream_handle : s->m->stream;
  return stream_allreduce(s, buf_dev, cs);
}

// ---- communicator of a minibatch-sharded chain ----------------------------------------------------------
int32_t bfx_comm_unique_id(void* id_out) {
  if (!id_out) return fail(BFX_ERR_BAD_ARG, "NULL argument");
  NcclApi* api = nccl_api();
  if (!api) return fail(BFX_ERR_UNSUPPORTED, "NCCL could not be loaded (set BFX_NCCL_LIB to libnccl.so.2)");
  NcclId id;
  const int r = api->GetUniqueId(&id);
  if (r != 0) return fail(BFX_ERR_CUDA, std::string("ncclGetUniqueId: ") + (api->GetErrorString ? api->GetErrorString(r) : "error"));
  std::memcpy(id_out, &id, sizeof id);
  return BFX_OK;
}

// ---- peer-memory exchange: export gbuf + the flag block with CUDA IPC, gather everybody's handles through the
// communicator, open the peers'.  Used only if EVERY rank succeeded (one all-reduced vote); otherwise the ranks keep
// the NCCL all-reduce.  Nothing here is an error for the caller.
static void peers_close(bfx_sampler* s) {
  for (void*& p : s->peer_opened) {
    if (p) cudaIpcCloseMemHandle(p);
    p = nullptr;
  }
  s->peer = PeerDev{};
}

static void peers_setup(bfx_sampler* s, int nranks, int rank) {
  peers_close(s);
  NcclApi* api = nccl_api();
  int max_ranks = PEER_MAX;
  if (const char* e = std::getenv("BFX_P2P_MAX_RANKS")) max_ranks = std::atoi(e);
  if (nranks < 2 || nranks > PEER_MAX || nranks > max_ranks || !api || !api->AllGather) return;
  bfx_model* m = s->m;
  cudaStream_t st = m->stream;
  const size_t P = (size_t)m->M.P;
  struct Handles { cudaIpcMemHandle_t buf, flg, sum; };
  int ok = 1;
  if (!s->gsum && cudaMalloc(&s->gsum, (P + 4) * sizeof(float)) != cudaSuccess) ok = 0;
  if (!s->peer_flags && cudaMalloc(&s->peer_flags, PEER_WORDS * sizeof(unsigned long long)) != cudaSuccess) ok = 0;
  if (ok) cudaMemsetAsync(s->peer_flags, 0, PEER_WORDS * sizeof(unsigned long long), st);
  std::vector<Handles> all((size_t)nranks);
  Handles mine{};
  if (ok && cudaIpcGetMemHandle(&mine.buf, s->gbuf) != cudaSuccess) ok = 0;
  if (ok && cudaIpcGetMemHandle(&mine.flg, s->peer_flags) != cudaSuccess) ok = 0;
  if (ok && cudaIpcGetMemHandle(&mine.sum, s->gsum) != cudaSuccess) ok = 0;
  cudaGetLastError();
  // gather the handles (every rank takes part in the collectives whatever its own outcome)
  Handles* d_h = nullptr;
  float* d_vote = nullptr;
  if (cudaMalloc(&d_h, sizeof(Handles) * (size_t)nranks) != cudaSuccess || cudaMalloc(&d_vote, sizeof(float)) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(d_h);
    cudaFree(d_vote);
    return;  // (cannot even vote: extremely unlikely; the peers' gather would hang, so this is left to fail loudly there)
  }
  cudaMemcpyAsync(d_h + rank, &mine, sizeof mine, cudaMemcpyHostToDevice, st);
  int r1 = api->AllGather(d_h + rank, d_h, sizeof(Handles), kNcclInt8, s->comm, st);
  cudaMemcpyAsync(all.data(), d_h, sizeof(Handles) * (size_t)nranks, cudaMemcpyDeviceToHost, st);
  if (cudaStreamSynchronize(st) != cudaSuccess || r1 != 0) ok = 0;
  PeerDev pd{};
  pd.n = nranks;
  pd.self = rank;
  for (int r = 0; r < nranks && ok; ++r) {
    if (r == rank) {
      pd.buf[r] = s->gbuf;
      pd.flg[r] = s->peer_flags;
      pd.sum[r] = s->gsum;
      continue;
    }
    void *pb = nullptr, *pf = nullptr, *ps = nullptr;
    if (cudaIpcOpenMemHandle(&pb, all[r].buf, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
        cudaIpcOpenMemHandle(&pf, all[r].flg, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
        cudaIpcOpenMemHandle(&ps, all[r].sum, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      if (pb) cudaIpcCloseMemHandle(pb);
      if (pf) cudaIpcCloseMemHandle(pf);
      ok = 0;
      break;
    }
    s->peer_opened[3 * r] = pb;
    s->peer_opened[3 * r + 1] = pf;
    s->peer_opened[3 * r + 2] = ps;
    pd.buf[r] = static_cast<const float*>(pb);
    pd.flg[r] = static_cast<unsigned long long*>(pf);
    pd.sum[r] = static_cast<float*>(ps);
  }
  // protocol: pull for two ranks, scatter beyond ($BFX_P2P_MODE = 1 | 2 overrides)
  pd.mode = nranks <= 2 ? PEER_MODE_PULL : PEER_MODE_SCATTER;
  if (const char* e = std::getenv("BFX_P2P_MODE")) pd.mode = std::atoi(e) == 2 ? PEER_MODE_SCATTER : PEER_MODE_PULL;
  // the vote: the exchange is used only if all ranks can reach all peers
  const float v = ok ? 1.f : 0.f;
  float total = 0.f;
  cudaMemcpyAsync(d_vote, &v, sizeof v, cudaMemcpyHostToDevice, st);
  const int r2 = api->AllReduce(d_vote, d_vote, 1, kNcclFloat32, kNcclSum, s->comm, st);
  cudaMemcpyAsync(&total, d_vote, sizeof total, cudaMemcpyDeviceToHost, st);
  const bool all_ok = cudaStreamSynchronize(st) == cudaSuccess && r2 == 0 && total > (float)nranks - 0.5f;
  cudaFree(d_h);
  cudaFree(d_vote);
  cudaGetLastError();
  if (all_ok) s->peer = pd;
  else peers_close(s);
}

int32_t bfx_sampler_comm_init(bfx_sampler* s, const void* id, int32_t nranks, int32_t rank) {
  if (!s || !id) return fail(BFX_ERR_BAD_ARG, "NULL argument");
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(BFX_ERR_BAD_ARG, "bad rank / nranks");
  if (!s->m->stream_regime) return fail(BFX_ERR_UNSUPPORTED, "minibatch sharding is implemented for the large-P (stream) regime");
  if (s->C != 1) return fail(BFX_ERR_BAD_ARG, "a minibatch-sharded sampler carries exactly one chain");
  NcclApi* api = nccl_api();
  if (!api) return fail(BFX_ERR_UNSUPPORTED, "NCCL could not be loaded (set BFX_NCCL_LIB to libnccl.so.2)");
  CK(cudaSetDevice(s->device));
  if (s->upd_graph) {
    cudaGraphExecDestroy(s->upd_graph);
    s->upd_graph = nullptr;
    cudaDeviceSynchronize();
  }
  peers_close(s);
  if (s->comm) {
    api->CommDestroy(s->comm);
    s->comm = nullptr;
  }
  NcclId nid;
  std::memcpy(&nid, id, sizeof nid);
  void* comm = nullptr;
  const int r = api->CommInitRank(&comm, nranks, nid, rank);
  if (r != 0) return fail(BFX_ERR_CUDA, std::string("ncclCommInitRank: ") + (api->GetErrorString ? api->GetErrorString(r) : "error"));
  s->comm = comm;
  s->comm_rank = rank;
  s->comm_nranks = nranks;
  if (!s->comm_stream) {
    CK(cudaStreamCreateWithFlags(&s->comm_stream, cudaStreamNonBlocking));
    for (int i = 0; i < BFX_MAX_LAYERS; ++i) CK(cudaEventCreateWithFlags(&s->hook.ev_fork[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&s->hook.ev_join, cudaEventDisableTiming));
    s->hook.reduce = hook_reduce;
    s->hook.ctx = s;
    s->hook.side = s->comm_stream;
  }
  // One eager collective on the exchange buffer now: NCCL establishes its transports (peer mappings, host-side
  // bootstrap between the ranks) on first use, which must not fall inside the stream capture of the update graph.
  {
    int32_t st = ensure_stream(s->m);
    if (st) return st;
    CK(cudaMemsetAsync(s->gbuf, 0, ((size_t)s->m->M.P + 3) * sizeof(float), s->m->stream));
    st = stream_allreduce(s, s->gbuf, s->m->stream);
    if (st) return st;
    CK(cudaStreamSynchronize(s->m->stream));
    peers_setup(s, nranks, rank);
  }
  if (s->upd_graph) {
    cudaGraphExecDestroy(s->upd_graph);
    s->upd_graph = nullptr;
  }
  return BFX_OK;
}

int32_t bfx_sampler_comm_destroy(bfx_sampler* s) {
  if (!s) return fail(BFX_ERR_BAD_ARG, "NULL argument");
  CK(cudaSetDevice(s->device));
  if (s->m->stream) cudaStreamSynchronize(s->m->stream);
  if (s->adv_event) cudaEventSynchronize(s->adv_event);
  // the update graph holds a reference on the communicator (captured collective): it goes first, ncclCommDestroy
  // waits for those references to drop
  if (s->upd_graph) {
    cudaGraphExecDestroy(s->upd_graph);
    s->upd_graph = nullptr;
    cudaDeviceSynchronize();
  }
  peers_close(s);
  if (s->comm) {
    if (NcclApi* api = nccl_api()) api->CommDestroy(s->comm);
    s->comm = nullptr;
  }
  s->comm_rank = 0;
  s->comm_nranks = 1;
  return BFX_OK;
}

int32_t bfx_sampler_peer_ranks(const bfx_sampler* s, int32_t* n) {
  if (!s || !n) return fail(BFX_ERR_BAD_ARG, "NULL argument");
  *n = (s->peer.n > 1 && !std::getenv("BFX_NO_P2P")) ? s->peer.n : 0;
  return BFX_OK;
}

int32_t bfx_sampler_graph_replays(const bfx_sampler* s, int64_t* n) {
  if (!s || !n) return fail(BFX_ERR_BAD_ARG, "NULL argument");
  *n = s->graph_replays;
  return BFX_OK;
}

int32_t bfx_model_regime(const bfx_model* m, int32_t* regime) {
  if (!m || !regime) return fail(BFX_ERR_BAD_ARG, "NULL argument");
  *regime = m->stream_regime ? BFX_REGIME_STREAM : BFX_REGIME_SMEM;
  return BFX_OK;
}

int32_t bfx_debug_batch_indices(const bfx_sampler* s, const bfx_run_desc* run, int64_t chain, int64_t step,
                                int64_t* idx_out, int64_t* B_out) {
  if (!s || !run || !idx_out || !B_out) return fail(BFX_ERR_BAD_ARG, "NULL argument");
  bfx_model* m = s->m;
  int32_t st = ensure_stream(m);
  if (st) return st;
  RunDev R{};
  st = prepare_schedule(const_cast<bfx_sampler*>(s), run, R);
  if (st) return st;
  long long *d_out = nullptr, *d_B = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_out);
    cudaFree(d_B);
  };
  CKC(cudaMalloc(&d_out, (size_t)R.B * 8));
  CKC(cudaMalloc(&d_B, 8));
  CKC(launch_debug_batch(R, (int)chain, (unsigned long long)step, d_out, d_B, m->stream));
  long long Bk = 0;
  CKC(cudaMemcpyAsync(&Bk, d_B, 8, cudaMemcpyDeviceToHost, m->stream));
  CKC(cudaStreamSynchronize(m->stream));
  CKC(cudaMemcpy(idx_out, d_out, (size_t)Bk * 8, cudaMemcpyDeviceToHost));
  *B_out = Bk;
  cleanup();
  return BFX_OK;
}

int32_t bfx_debug_noise(const bfx_sampler* s, const bfx_run_desc* run, int64_t chain, int64_t step, int32_t slot,
                        float* normals_out, float* uniform_out) {
  if (!s || !run || !normals_out || !uniform_out) return fail(BFX_ERR_BAD_ARG, "NULL argument");
  bfx_model* m = s->m;
  int32_t st = ensure_stream(m);
  if (st) return st;
  const int P = m->M.P;
  float *d_n = nullptr, *d_u = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_n);
    cudaFree(d_u);
  };
  CKC(cudaMalloc(&d_n, (size_t)P * 4));
  CKC(cudaMalloc(&d_u, 4));
  CKC(launch_debug_noise(run->seed, (unsigned)(run->chain_offset + chain), (unsigned long long)step, (unsigned)slot, P,
                         m->d_flat_pad, d_n, d_u, m->stream));
  CKC(cudaMemcpyAsync(normals_out, d_n, (size_t)P * 4, cudaMemcpyDeviceToHost, m->stream));
  CKC(cudaMemcpyAsync(uniform_out, d_u, 4, cudaMemcpyDeviceToHost, m->stream));
  CKC(cudaStreamSynchronize(m->stream));
  cleanup();
  return BFX_OK;
}

}  // extern "C"

// error reporting for the other translation units of the library (same thread-local message as above)
namespace bfx {
int32_t report_error(int32_t code, const std::string& msg) { return ::fail(code, msg); }
}  // namespace bfx
