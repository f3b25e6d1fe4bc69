// Kernels of the large-P ("stream") regime, see bfx_stream.cuh.
#include "bfx_stream.cuh"

#include "bfx_device.cuh"

namespace bfx {

// =====================================================================================================
// tiled FP32 GEMM, column-major C[m + ldc*n] = sum_k A(m,k) * B(k,n), 128 x 128 x 16 tiles, 256 threads,
// 8 x 8 register tile per thread (two 4-wide groups in each dimension: conflict-free 128-bit smem loads)
// =====================================================================================================
enum { EPI_BIAS_ACT = 0, EPI_MUL_DACT = 1, EPI_PLAIN = 2 };

struct GemmArgs {
  int M, N, K;
  const float* A;
  long long sam, sak;  // A(m,k) = A[m*sam + k*sak]
  const float* B;
  long long sbk, sbn;  // B(k,n) = B[k*sbk + n*sbn]
  int b_ones_n;        // >= 0: column n == b_ones_n of B is the constant 1 (bias gradient), no memory read
  float* C;
  long long ldc;
  const float* bias;   // EPI_BIAS_ACT
  int act;
  const float* aux;    // EPI_MUL_DACT: C = acc * act'(aux[m + ldaux*n])
  long long ldaux;
  int accumulate;      // EPI_PLAIN: C += acc
  long long kchunk;    // split-K: blockIdx.z handles k in [z*kchunk, (z+1)*kchunk) and writes C + z*part_stride
  long long part_stride;
  __nv_bfloat16* chi;  // optional BF16 halves of the output (tensor-core consumers)
  __nv_bfloat16* clo;
};

constexpr int GBM = 128, GBN = 128, GBK = 16, GLD = 132;

template <bool A_MCONTIG, bool B_KCONTIG, int EPI>
__global__ void __launch_bounds__(256) k_sgemm(const GemmArgs g) {
  __shared__ __align__(16) float As[GBK][GLD];
  __shared__ __align__(16) float Bs[GBK][GLD];
  const int t = threadIdx.x;
  const int m0 = blockIdx.x * GBM, n0 = blockIdx.y * GBN;
  const long long kbeg = (long long)blockIdx.z * g.kchunk;
  long long kend = kbeg + g.kchunk;
  if (kend > g.K) kend = g.K;
  const int tx = t & 15, ty = t >> 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (long long k0 = kbeg; k0 < kend; k0 += GBK) {
    // ---- global -> shared (zero filled outside the matrix) -----------------------------------------
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int m, k;
      if (A_MCONTIG) { m = t & 127; k = (t >> 7) + 2 * i; }
      else { k = t & 15; m = (t >> 4) + 16 * i; }
      const long long gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < g.M && gk < kend) ? __ldg(g.A + gm * g.sam + gk * g.sak) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int n, k;
      if (B_KCONTIG) { k = t & 15; n = (t >> 4) + 16 * i; }
      else { n = t & 127; k = (t >> 7) + 2 * i; }
      const long long gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < g.N && gk < kend) v = (gn == g.b_ones_n) ? 1.f : __ldg(g.B + gk * g.sbk + gn * g.sbn);
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][tx * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + tx * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][ty * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + ty * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  // ---- epilogue ---------------------------------------------------------------------------------------
  float* C = g.C + (long long)blockIdx.z * g.part_stride;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const long long n = n0 + (j < 4 ? ty * 4 + j : 64 + ty * 4 + (j - 4));
    if (n >= g.N) continue;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const long long m = m0 + (i < 4 ? tx * 4 + i : 64 + tx * 4 + (i - 4));
      if (m >= g.M) continue;
      float v = acc[i][j];
      if (EPI == EPI_BIAS_ACT) v = act_apply(g.act, v + __ldg(g.bias + m));
      else if (EPI == EPI_MUL_DACT) v *= act_deriv_from_out(g.act, __ldg(g.aux + m + g.ldaux * n));
      else if (g.accumulate) v += C[m + g.ldc * n];
      C[m + g.ldc * n] = v;
      if (g.chi) {
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        g.chi[m + g.ldc * n] = h;
        g.clo[m + g.ldc * n] = __float2bfloat16_rn(v - __bfloat162float(h));
      }
    }
  }
}

// C[i] (+)= sum_z part[z][i]   (deterministic split-K reduction)
__global__ void k_reduce_parts(float* C, const float* part, long long count, int nparts, long long stride, int accumulate) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    float s = accumulate ? C[i] : 0.f;
    for (int z = 0; z < nparts; ++z) s += part[z * stride + i];
    C[i] = s;
  }
}

template <bool AM, bool BK, int EPI>
static cudaError_t launch_gemm(const GemmArgs& g, int ksplit, cudaStream_t st) {
  dim3 grid((g.M + GBM - 1) / GBM, (g.N + GBN - 1) / GBN, ksplit);
  k_sgemm<AM, BK, EPI><<<grid, 256, 0, st>>>(g);
  return cudaGetLastError();
}

// =====================================================================================================
// batch gather, likelihood, scalar kernels
// =====================================================================================================
// FP32 array -> BF16 hi / lo halves (theta after an update; activations produced by an FP32 layer)
__global__ void k_split(const float* __restrict__ src, long long n, __nv_bfloat16* __restrict__ hi,
                        __nv_bfloat16* __restrict__ lo) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = src[i];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

// bias gradient of a tensor-core layer: db[o] += sum_b D[o + nout*b], deterministic two-level sum
// (optionally weighted by wgt[b]: the weight gradient of a 1-wide output layer is sum_b dl[b] * act[:, b])
__global__ void k_rowsum_part(const float* __restrict__ D, int nout, int nc, const float* __restrict__ wgt,
                              float* __restrict__ part) {
  const int o = blockIdx.x * 32 + (threadIdx.x & 31);
  const int sy = threadIdx.x >> 5;                      // 8 sub-slices per block
  const int slice = blockIdx.y * 8 + sy, nslices = gridDim.y * 8;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (o < nout) {
    int b = slice;
    for (; b + 3 * nslices < nc; b += 4 * nslices) {  // four independent loads in flight
      const float d0 = D[o + (long long)nout * b], d1 = D[o + (long long)nout * (b + nslices)];
      const float d2 = D[o + (long long)nout * (b + 2 * nslices)], d3 = D[o + (long long)nout * (b + 3 * nslices)];
      s0 += wgt ? wgt[b] * d0 : d0;
      s1 += wgt ? wgt[b + nslices] * d1 : d1;
      s2 += wgt ? wgt[b + 2 * nslices] * d2 : d2;
      s3 += wgt ? wgt[b + 3 * nslices] * d3 : d3;
    }
    for (; b < nc; b += nslices) s0 += wgt ? wgt[b] * D[o + (long long)nout * b] : D[o + (long long)nout * b];
    part[(long long)slice * nout + o] = (s0 + s1) + (s2 + s3);
  }
}

// 1-wide output layer: out[b] = act(sum_i w[i] * A[i + nin*b] + b0), one warp per column
__global__ void k_out_fwd(const float* __restrict__ w, const float* __restrict__ bias, const float* __restrict__ A, int nin,
                          int nc, int act, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long col = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (col >= nc) return;
  const float* a = A + col * nin;
  float s = 0.f;
  for (int i = lane; i < nin; i += 32) s = fmaf(__ldg(w + i), a[i], s);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[col] = act_apply(act, s + __ldg(bias));
}

// its input gradient: D[i + nin*b] = w[i] * dl[b] * act'(A[i + nin*b])  (+ BF16 halves for tensor-core consumers)
__global__ void k_out_dx(const float* __restrict__ w, const float* __restrict__ dl, const float* __restrict__ A, int nin,
                         int nc, int act_prev, float* __restrict__ D, __nv_bfloat16* __restrict__ hi,
                         __nv_bfloat16* __restrict__ lo) {
  const long long total = (long long)nin * nc;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long b = e / nin;
    const int i = (int)(e - b * nin);
    const float v = __ldg(w + i) * dl[b] * act_deriv_from_out(act_prev, A[e]);
    D[e] = v;
    if (hi) {
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      hi[e] = h;
      lo[e] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}
__global__ void k_rowsum_final(const float* __restrict__ part, int nout, int nslices, float* __restrict__ g) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= nout) return;
  float s = 0.f;
  for (int i = 0; i < nslices; ++i) s += part[(long long)i * nout + o];
  g[o] += s;
}

__global__ void k_gather(const float* __restrict__ x, const float* __restrict__ y, StreamBatch b, long long c0, int nc,
                         int d, float* __restrict__ X, float* __restrict__ yb, __nv_bfloat16* __restrict__ xhi,
                         __nv_bfloat16* __restrict__ xlo) {
  // one warp per column: the d features of an observation are contiguous in x
  const int lane = threadIdx.x & 31;
  const long long col = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (col >= nc) return;
  BatchSrc bs;
  bs.idx = b.idx; bs.n = b.n; bs.first = b.first; bs.k0 = b.k0; bs.k1 = b.k1; bs.shuffle = b.shuffle; bs.full = b.full;
  const long long o = batch_obs(bs, c0 + col);
  for (int f = lane; f < d; f += 32) {
    const float v = __ldg(x + o * d + f);
    X[col * d + f] = v;
    if (xhi) {
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      xhi[col * d + f] = h;
      xlo[col * d + f] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
  if (lane == 0) yb[col] = __ldg(y + o);
}

__global__ void k_like(StreamModel M, const float* __restrict__ theta, const float* __restrict__ out, const float* __restrict__ yb,
                       int nc, float nb, int out_act, float* __restrict__ dl, StreamScalars* sc) {
  const float inv_sigma = 1.f / expf(theta[M.P - 1]);
  float s0 = 0.f, s1 = 0.f;
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nc; b += gridDim.x * blockDim.x) {
    const float yh = out[b];
    const float z = (yb[b] - yh) * inv_sigma;
    float d;
    if (M.like_kind == LK_NORMAL) {
      s0 += z * z;
      d = nb * z * inv_sigma;
    } else {
      const float zz = z * z;
      const float w = (M.nu + 1.f) * z / (M.nu + zz);
      s0 += log1pf(zz / M.nu);
      s1 += w * z;
      d = nb * w * inv_sigma;
    }
    dl[b] = d * act_deriv_from_out(out_act, yh);
  }
  __shared__ double r0[32], r1[32];
  double d0 = s0, d1 = s1;
  for (int o = 16; o > 0; o >>= 1) {
    d0 += __shfl_xor_sync(0xffffffffu, d0, o);
    d1 += __shfl_xor_sync(0xffffffffu, d1, o);
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { r0[w] = d0; r1[w] = d1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    d0 = 0; d1 = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { d0 += r0[i]; d1 += r1[i]; }
    atomicAdd(&sc->ds0, d0);
    atomicAdd(&sc->ds1, d1);
    if (blockIdx.x == 0) atomicAdd(&sc->bcount, (double)nc);
  }
}

__global__ void k_zero(float* g, long long P, StreamScalars* sc) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < P; i += (long long)gridDim.x * blockDim.x) g[i] = 0.f;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    sc->ds0 = sc->ds1 = sc->bcount = 0.0;
    sc->th2 = sc->gn2 = sc->acc0 = sc->acc1 = 0.0;
  }
}

__global__ void k_pack_sums(float* g, long long P, StreamScalars* sc, int to_buffer) {
  if (to_buffer) {
    g[P] = (float)sc->ds0;
    g[P + 1] = (float)sc->ds1;
    g[P + 2] = (float)sc->bcount;
  } else {
    sc->ds0 = g[P];
    sc->ds1 = g[P + 1];
    sc->bcount = g[P + 2];
  }
}

// sigma prior + d/d theta_like from the (globally reduced) likelihood sums: SURVEY.md 8a rows a8 / a9
__global__ void k_scalar_like(StreamModel M, const float* theta, StreamScalars* sc, float nb) {
  const float sl = theta[M.P - 1];
  const double sigma = (double)expf(sl);
  double lps, dlps;
  if (M.sprior_kind == SP_GAMMA) {
    lps = M.sprior_const + ((double)M.sprior_a - 1.0) * (double)sl - sigma / (double)M.sprior_b + (double)sl;
    dlps = (double)M.sprior_a - sigma / (double)M.sprior_b;
  } else {
    lps = M.sprior_const - ((double)M.sprior_a + 1.0) * (double)sl - (double)M.sprior_b / sigma + (double)sl;
    dlps = -(double)M.sprior_a + (double)M.sprior_b / sigma;
  }
  const double B = sc->bcount;
  double like, dlike;
  if (M.like_kind == LK_NORMAL) {
    like = -0.5 * B * 1.8378770664093453 - 0.5 * sc->ds0;
    dlike = sc->ds0;
  } else {
    like = B * M.tdist_const - 0.5 * ((double)M.nu + 1.0) * sc->ds0;
    dlike = sc->ds1;
  }
  like += -B * (double)sl + lps;
  dlike += -B + dlps;
  sc->glike = (float)((double)nb * dlike);
  sc->value = (double)nb * like;  // the prior is added by k_scalar_value
}

// g_net -= c2 * theta_net; g_like = glike; sums of theta_net^2 and g^2
__global__ void k_finish(StreamModel M, const float* __restrict__ theta, float* __restrict__ g, StreamScalars* sc, int want_grad) {
  float a0 = 0.f, a1 = 0.f;
  const float c2 = M.prior_c2, glike = sc->glike;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < M.P; i += (long long)gridDim.x * blockDim.x) {
    const float th = theta[i];
    float gg;
    if (i < M.Pnet) {
      a0 += th * th;
      gg = g[i] - c2 * th;
    } else {
      gg = glike;
    }
    if (want_grad) {
      g[i] = gg;
      a1 += gg * gg;
    }
  }
  __shared__ double r0[32], r1[32];
  double d0 = a0, d1 = a1;
  for (int o = 16; o > 0; o >>= 1) {
    d0 += __shfl_xor_sync(0xffffffffu, d0, o);
    d1 += __shfl_xor_sync(0xffffffffu, d1, o);
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { r0[w] = d0; r1[w] = d1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    d0 = 0; d1 = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { d0 += r0[i]; d1 += r1[i]; }
    atomicAdd(&sc->th2, d0);
    atomicAdd(&sc->gn2, d1);
  }
}

__global__ void k_scalar_value(StreamModel M, StreamScalars* sc) {
  sc->value += M.prior_c0n - 0.5 * (double)M.prior_c2 * sc->th2;
}

// =====================================================================================================
// host orchestration of one evaluation
// =====================================================================================================
cudaError_t stream_work_alloc(const StreamModel& M, long long Bc, StreamWork& W) {
  W = StreamWork{};
  W.Bc = Bc;
  cudaError_t e = cudaSuccess;
  int maxw = 1;
  for (int l = 0; l < M.nlayers; ++l) maxw = M.L[l].nout > maxw ? M.L[l].nout : maxw;
  e = cudaMalloc(&W.act[0], (size_t)M.d_in * Bc * sizeof(float));
  for (int l = 0; l < M.nlayers && e == cudaSuccess; ++l) e = cudaMalloc(&W.act[l + 1], (size_t)M.L[l].nout * Bc * sizeof(float));
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaMalloc(&W.delta[i], (size_t)maxw * Bc * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&W.yb, (size_t)Bc * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&W.dl, (size_t)Bc * sizeof(float));
  // split-K partials: up to 32 partial copies of the largest [W | b] block
  long long big = 0;
  for (int l = 0; l < M.nlayers; ++l) {
    long long c = (long long)M.L[l].nout * (M.L[l].nin + 1);
    big = c > big ? c : big;
  }
  W.part_floats = big * 32;
  if (e == cudaSuccess) e = cudaMalloc(&W.part, (size_t)W.part_floats * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&W.rowpart, (size_t)64 * maxw * sizeof(float));
  if (M.tc) {
    if (e == cudaSuccess) e = cudaMalloc(&W.act_hi[0], (size_t)M.d_in * Bc * 2);
    if (e == cudaSuccess) e = cudaMalloc(&W.act_lo[0], (size_t)M.d_in * Bc * 2);
    for (int l = 0; l < M.nlayers && e == cudaSuccess; ++l) {
      e = cudaMalloc(&W.act_hi[l + 1], (size_t)M.L[l].nout * Bc * 2);
      if (e == cudaSuccess) e = cudaMalloc(&W.act_lo[l + 1], (size_t)M.L[l].nout * Bc * 2);
    }
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
      e = cudaMalloc(&W.delta_hi[i], (size_t)maxw * Bc * 2);
      if (e == cudaSuccess) e = cudaMalloc(&W.delta_lo[i], (size_t)maxw * Bc * 2);
    }
    if (e == cudaSuccess) e = cudaMalloc(&W.w_hi, (size_t)(M.P + 8) * 2);
    if (e == cudaSuccess) e = cudaMalloc(&W.w_lo, (size_t)(M.P + 8) * 2);
  }
  if (e != cudaSuccess) stream_work_free(W);
  return e;
}

void stream_work_free(StreamWork& W) {
  for (int l = 0; l <= BFX_MAX_LAYERS; ++l) cudaFree(W.act[l]);
  cudaFree(W.delta[0]);
  cudaFree(W.delta[1]);
  cudaFree(W.yb);
  cudaFree(W.dl);
  cudaFree(W.part);
  cudaFree(W.rowpart);
  for (int l = 0; l <= BFX_MAX_LAYERS; ++l) {
    cudaFree(W.act_hi[l]);
    cudaFree(W.act_lo[l]);
  }
  for (int i = 0; i < 2; ++i) {
    cudaFree(W.delta_hi[i]);
    cudaFree(W.delta_lo[i]);
  }
  cudaFree(W.w_hi);
  cudaFree(W.w_lo);
  W = StreamWork{};
}

#define SCK(call)                      \
  do {                                 \
    cudaError_t e__ = (call);          \
    if (e__ != cudaSuccess) return e__; \
  } while (0)

cudaError_t stream_zero(float* g, long long P, StreamScalars* sc, cudaStream_t st) {
  k_zero<<<296, 256, 0, st>>>(g, P, sc);
  return cudaGetLastError();
}

cudaError_t stream_pack_sums(float* g, long long P, StreamScalars* sc, bool to_buffer, cudaStream_t st) {
  k_pack_sums<<<1, 1, 0, st>>>(g, P, sc, to_buffer ? 1 : 0);
  return cudaGetLastError();
}

cudaError_t stream_eval_like(const StreamModel& M, StreamWork& W, const float* theta, const float* x, const float* y,
                             const StreamBatch& b, float nb, bool want_grad, float* g, StreamScalars* sc,
                             cudaStream_t st, float* yhat_out) {
  const int nl = M.nlayers;
  if (M.tc) {  // BF16 halves of theta for the tensor-core layers
    k_split<<<592, 256, 0, st>>>(theta, M.P, W.w_hi, W.w_lo);
    SCK(cudaGetLastError());
  }
  for (long long c0 = 0; c0 < b.B; c0 += W.Bc) {
    const int nc = (int)((b.B - c0) < W.Bc ? (b.B - c0) : W.Bc);
    {
      const long long threads = (long long)nc * 32;
      k_gather<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(x, y, b, c0, nc, M.d_in, W.act[0], W.yb,
                                                               M.tc ? W.act_hi[0] : nullptr, M.tc ? W.act_lo[0] : nullptr);
      SCK(cudaGetLastError());
    }
    // forward: act[l+1] (nout x nc) = act(W_l act[l] + b_l)
    for (int l = 0; l < nl; ++l) {
      const StreamLayer& L = M.L[l];
      if (L.tc) {
        TcGemmArgs t{};
        t.M = L.nout; t.N = nc; t.K = L.nin;
        t.A = TcOperand{W.w_hi + L.w_off, W.w_lo + L.w_off, L.nout};
        t.B = TcOperand{W.act_hi[l], W.act_lo[l], L.nin};
        t.C = W.act[l + 1]; t.Chi = W.act_hi[l + 1]; t.Clo = W.act_lo[l + 1]; t.ldc = L.nout;
        t.bias = theta + L.b_off; t.act = L.act;
        t.kchunk = L.nin; t.part_stride = 0;
        SCK(tc_gemm_forward(t, st));
        continue;
      }
      if (L.nout == 1) {
        const long long threads = (long long)nc * 32;
        k_out_fwd<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(theta + L.w_off, theta + L.b_off, W.act[l], L.nin, nc,
                                                                  L.act, W.act[l + 1]);
        SCK(cudaGetLastError());
        continue;
      }
      GemmArgs a{};
      a.M = L.nout; a.N = nc; a.K = L.nin;
      a.A = theta + L.w_off; a.sam = 1; a.sak = L.nout;
      a.B = W.act[l]; a.sbk = 1; a.sbn = L.nin; a.b_ones_n = -1;
      a.C = W.act[l + 1]; a.ldc = L.nout;
      a.bias = theta + L.b_off; a.act = L.act;
      a.kchunk = L.nin; a.part_stride = 0;
      if (M.tc && l + 1 < nl && M.L[l + 1].tc) { a.chi = W.act_hi[l + 1]; a.clo = W.act_lo[l + 1]; }
      SCK((launch_gemm<true, true, EPI_BIAS_ACT>(a, 1, st)));
    }
    k_like<<<148, 256, 0, st>>>(M, theta, W.act[nl], W.yb, nc, nb, M.L[nl - 1].act, W.dl, sc);
    SCK(cudaGetLastError());
    if (yhat_out) SCK(cudaMemcpyAsync(yhat_out + c0, W.act[nl], (size_t)nc * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (!want_grad) continue;
    // backward: delta of layer l in D (nout x nc); the output layer's delta is dl (1 x nc)
    const float* D = W.dl;
    const __nv_bfloat16 *Dhi = nullptr, *Dlo = nullptr;
    int ping = 0;
    for (int l = nl - 1; l >= 0; --l) {
      const StreamLayer& L = M.L[l];
      if (L.tc) {
        // dW (nout x nin) += D * act[l]^T on the tensor cores, split over the batch; db by a deterministic row sum
        const long long cnt = (long long)L.nout * L.nin;
        int ksplit = 1;
        const int tiles = ((L.nout + 127) / 128) * ((L.nin + 127) / 128);
        while (ksplit < 32 && tiles * ksplit < 148 && (nc / (ksplit * 2)) >= 512) ksplit *= 2;
        TcGemmArgs t{};
        t.M = L.nout; t.N = L.nin; t.K = nc;
        t.A = TcOperand{Dhi, Dlo, L.nout};
        t.B = TcOperand{W.act_hi[l], W.act_lo[l], L.nin};
        t.ldc = L.nout;
        if (ksplit == 1) {
          t.C = g + L.w_off; t.accumulate = 1; t.kchunk = nc; t.part_stride = 0;
          SCK(tc_gemm_dw(t, 1, st));
        } else {
          long long kch = (nc + ksplit - 1) / ksplit;
          kch = (kch + 31) / 32 * 32;
          t.C = W.part; t.accumulate = 0; t.kchunk = kch; t.part_stride = cnt;
          SCK(tc_gemm_dw(t, ksplit, st));
          k_reduce_parts<<<(unsigned)((cnt + 255) / 256 < 1184 ? (cnt + 255) / 256 : 1184), 256, 0, st>>>(g + L.w_off, W.part, cnt, ksplit, cnt, 1);
          SCK(cudaGetLastError());
        }
        {
          dim3 grid((L.nout + 31) / 32, 8);
          k_rowsum_part<<<grid, 256, 0, st>>>(D, L.nout, nc, nullptr, W.rowpart);
          SCK(cudaGetLastError());
          k_rowsum_final<<<(L.nout + 127) / 128, 128, 0, st>>>(W.rowpart, L.nout, 64, g + L.b_off);
          SCK(cudaGetLastError());
        }
        if (l > 0) {
          // delta_{l-1} (nin x nc) = (W_l^T D) .* act'(act[l])
          TcGemmArgs u{};
          u.M = L.nin; u.N = nc; u.K = L.nout;
          u.A = TcOperand{W.w_hi + L.w_off, W.w_lo + L.w_off, L.nout};
          u.B = TcOperand{Dhi, Dlo, L.nout};
          u.C = W.delta[ping]; u.Chi = W.delta_hi[ping]; u.Clo = W.delta_lo[ping]; u.ldc = L.nin;
          u.aux = W.act[l]; u.ldaux = L.nin; u.act = M.L[l - 1].act;
          u.kchunk = L.nout; u.part_stride = 0;
          SCK(tc_gemm_dx(u, st));
          D = W.delta[ping]; Dhi = W.delta_hi[ping]; Dlo = W.delta_lo[ping];
          ping ^= 1;
        }
        continue;
      }
      if (L.nout == 1 && W.rowpart) {
        // 1-wide output layer: dW[i] += sum_b dl[b] act[l][i, b], db += sum_b dl[b]; delta = w dl^T .* act'
        dim3 grid((L.nin + 31) / 32, 8);
        k_rowsum_part<<<grid, 256, 0, st>>>(W.act[l], L.nin, nc, D, W.rowpart);
        SCK(cudaGetLastError());
        k_rowsum_final<<<(L.nin + 127) / 128, 128, 0, st>>>(W.rowpart, L.nin, 64, g + L.w_off);
        SCK(cudaGetLastError());
        k_rowsum_part<<<dim3(1, 8), 256, 0, st>>>(D, 1, nc, nullptr, W.rowpart);
        SCK(cudaGetLastError());
        k_rowsum_final<<<1, 32, 0, st>>>(W.rowpart, 1, 64, g + L.b_off);
        SCK(cudaGetLastError());
        if (l > 0) {
          const bool tcn = M.tc && M.L[l - 1].tc;
          k_out_dx<<<1184, 256, 0, st>>>(theta + L.w_off, D, W.act[l], L.nin, nc, M.L[l - 1].act, W.delta[ping],
                                        tcn ? W.delta_hi[ping] : nullptr, tcn ? W.delta_lo[ping] : nullptr);
          SCK(cudaGetLastError());
          D = W.delta[ping]; Dhi = W.delta_hi[ping]; Dlo = W.delta_lo[ping];
          ping ^= 1;
        }
        continue;
      }
      // [dW | db] (nout x (nin+1), exactly the flat layout of the layer) += D * [act[l] ; 1]^T, split over the batch
      {
        const long long cnt = (long long)L.nout * (L.nin + 1);
        int ksplit = 1;
        const int tiles = ((L.nout + GBM - 1) / GBM) * ((L.nin + 1 + GBN - 1) / GBN);
        while (ksplit < 32 && tiles * ksplit < 148 && (nc / (ksplit * 2)) >= 256) ksplit *= 2;
        GemmArgs a{};
        a.M = L.nout; a.N = L.nin + 1; a.K = nc;
        a.A = D; a.sam = 1; a.sak = L.nout;
        a.B = W.act[l]; a.sbk = L.nin; a.sbn = 1; a.b_ones_n = L.nin;
        a.ldc = L.nout;
        if (ksplit == 1) {
          a.C = g + L.w_off; a.accumulate = 1; a.kchunk = nc; a.part_stride = 0;
          SCK((launch_gemm<true, false, EPI_PLAIN>(a, 1, st)));
        } else {
          a.C = W.part; a.accumulate = 0; a.kchunk = (nc + ksplit - 1) / ksplit; a.part_stride = cnt;
          SCK((launch_gemm<true, false, EPI_PLAIN>(a, ksplit, st)));
          k_reduce_parts<<<(unsigned)((cnt + 255) / 256 < 1184 ? (cnt + 255) / 256 : 1184), 256, 0, st>>>(g + L.w_off, W.part, cnt, ksplit, cnt, 1);
          SCK(cudaGetLastError());
        }
      }
      if (l > 0) {
        // delta_{l-1} (nin x nc) = (W_l^T D) .* act'(act[l])
        GemmArgs a{};
        a.M = L.nin; a.N = nc; a.K = L.nout;
        a.A = theta + L.w_off; a.sam = L.nout; a.sak = 1;
        a.B = D; a.sbk = 1; a.sbn = L.nout; a.b_ones_n = -1;
        a.C = W.delta[ping]; a.ldc = L.nin;
        a.aux = W.act[l]; a.ldaux = L.nin; a.act = M.L[l - 1].act;
        a.kchunk = L.nout; a.part_stride = 0;
        if (M.tc && M.L[l - 1].tc) { a.chi = W.delta_hi[ping]; a.clo = W.delta_lo[ping]; }
        SCK((launch_gemm<false, true, EPI_MUL_DACT>(a, 1, st)));
        D = W.delta[ping]; Dhi = W.delta_hi[ping]; Dlo = W.delta_lo[ping];
        ping ^= 1;
      }
    }
  }
  return cudaSuccess;
}

cudaError_t stream_finish(const StreamModel& M, const float* theta, float* g, StreamScalars* sc, float nb, bool want_grad,
                          cudaStream_t st) {
  k_scalar_like<<<1, 1, 0, st>>>(M, theta, sc, nb);
  SCK(cudaGetLastError());
  k_finish<<<296, 256, 0, st>>>(M, theta, g, sc, want_grad ? 1 : 0);
  SCK(cudaGetLastError());
  k_scalar_value<<<1, 1, 0, st>>>(M, sc);
  return cudaGetLastError();
}

// =====================================================================================================
// sampler updates (flat order, Philox quad q = J / 4)
// =====================================================================================================
struct UpdDev {
  SamplerDev S;
  long long P;
  float *theta, *mom, *minv, *rmsv, *g;
  StreamScalars* sc;
  NoiseKey key;
  unsigned long long gstep;
  const float* inj;
  float* sample_out;
  float* kinetic_out;
};

__device__ __forceinline__ float noise_flat(const UpdDev& U, unsigned slot, long long J) {
  if (U.inj) return U.inj[J];
  const float4 z = normal_quad(U.key, U.gstep, slot, (unsigned)(J >> 2));
  const int e = (int)(J & 3);
  return e == 0 ? z.x : e == 1 ? z.y : e == 2 ? z.z : z.w;
}

template <class F>
__device__ __forceinline__ void for_flat(long long P, F f) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < P; i += (long long)gridDim.x * blockDim.x) f(i);
}

__device__ __forceinline__ void grid_acc(float a0, float a1, StreamScalars* sc) {
  __shared__ double r0[32], r1[32];
  double d0 = a0, d1 = a1;
  for (int o = 16; o > 0; o >>= 1) {
    d0 += __shfl_xor_sync(0xffffffffu, d0, o);
    d1 += __shfl_xor_sync(0xffffffffu, d1, o);
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { r0[w] = d0; r1[w] = d1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    d0 = 0; d1 = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { d0 += r0[i]; d1 += r1[i]; }
    atomicAdd(&sc->acc0, d0);
    atomicAdd(&sc->acc1, d1);
  }
}

// SGLD (sgld.jl:96-112)
__global__ void k_upd_sgld(UpdDev U) {
  if (U.sc->status) return;
  const SamplerDev& S = U.S;
  float alpha = S.stepsize_a * powf(S.stepsize_b + (float)U.sc->t, -S.stepsize_gamma);
  alpha = fmaxf(alpha, S.min_stepsize);
  const float ng = sqrtf((float)U.sc->gn2);
  const float ha = 0.5f * alpha * (ng > S.maxnorm ? S.maxnorm / ng : 1.f), sa = sqrtf(alpha);
  for_flat(U.P, [&](long long i) {
    const float t = U.theta[i] + ha * U.g[i] + sa * noise_flat(U, 0, i);
    U.theta[i] = t;
    if (U.sample_out) U.sample_out[i] = t;
  });
}

// SGNHT (sgnht.jl:64-90)
__global__ void k_upd_sgnht(UpdDev U) {
  if (U.sc->status) return;
  const SamplerDev& S = U.S;
  const float l = S.l, xi = U.sc->xi, cn = sqrtf(l * 2.f * S.A);
  float a0 = 0.f, a1 = 0.f;
  for_flat(U.P, [&](long long i) {
    float p = U.mom[i];
    p = p - xi * l * p + l * U.g[i] + cn * noise_flat(U, 0, i);
    U.mom[i] = p;
    const float t = U.theta[i] + l * p;
    U.theta[i] = t;
    if (U.sample_out) U.sample_out[i] = t;
    a0 += p * p;
    a1 += (t != t) ? 1.f : 0.f;
  });
  grid_acc(a0, a1, U.sc);
}

// RMSPropMassAdapter (rmspropmassadapter.jl:26-46) on the gradient already in g
__global__ void k_upd_rmsprop(UpdDev U) {
  if (U.sc->status || U.sc->ma_t > U.S.ma_adapt_steps) return;
  const bool first = U.sc->ma_t == 1;
  const float inv_ng = 1.f / sqrtf((float)U.sc->gn2);
  for_flat(U.P, [&](long long i) {
    const float gg = U.g[i] * inv_ng;
    float V = first ? 1.f : U.rmsv[i];
    V = U.S.ma_alpha * V + (1.f - U.S.ma_alpha) * gg * gg;
    U.rmsv[i] = V;
    U.minv[i] = 1.f / (U.S.ma_lambda + sqrtf(V));
  });
}

// SGNHTS first half (sgnht-s.jl:95-97)
__global__ void k_upd_sgnhts_a(UpdDev U) {
  if (U.sc->status) return;
  const float hl = 0.5f * U.S.l;
  float a0 = 0.f;
  for_flat(U.P, [&](long long i) {
    const float mi = U.minv ? U.minv[i] : 1.f;
    const float p = U.mom[i] + hl * U.g[i];
    U.mom[i] = p;
    U.theta[i] += hl * mi * p;
    a0 += p * p * mi;
  });
  grid_acc(a0, 0.f, U.sc);
}

// SGNHTS second half (sgnht-s.jl:98-104)
__global__ void k_upd_sgnhts_b(UpdDev U) {
  if (U.sc->status) return;
  const float hl = 0.5f * U.S.l, e1 = U.sc->e1, scl = U.sc->sc;
  float a0 = 0.f, a1 = 0.f;
  for_flat(U.P, [&](long long i) {
    const float mi = U.minv ? U.minv[i] : 1.f;
    const float p = e1 * U.mom[i] + scl * rsqrtf(mi) * noise_flat(U, 0, i);
    U.mom[i] = p;
    const float t = U.theta[i] + hl * mi * p;
    U.theta[i] = t;
    if (U.sample_out) U.sample_out[i] = t;
    a0 += p * p * mi;
    a1 += (t != t) ? 1.f : 0.f;
  });
  grid_acc(a0, a1, U.sc);
}

// scalar bookkeeping between / after the passes.  phase: 0 NaN-in-g check, 1 after SGLD, 2 after SGNHT,
// 3 between the SGNHTS halves, 4 after SGNHTS, 5 after RMSProp
__global__ void k_upd_scalar(UpdDev U, int phase) {
  StreamScalars* sc = U.sc;
  const SamplerDev& S = U.S;
  if (sc->status) return;
  const float n = (float)U.P;
  if (phase == 0) {
    if (sc->gn2 != sc->gn2) sc->status = 4;  // error("NaN in g")
  } else if (phase == 1) {
    sc->nsampled += 1;
    sc->t += 1;
  } else if (phase == 2) {
    const float kin = (float)sc->acc0 / n;
    if (U.kinetic_out) *U.kinetic_out = kin;
    sc->kin = kin;
    sc->xi = sc->xi + (kin - 1.f) * S.l;
    if (sc->acc1 > 0.0) { sc->status = 5; return; }
    sc->nsampled += 1;
    sc->t += 1;
  } else if (phase == 3) {
    const float xi = sc->xi + S.l / (2.f * S.mu) * ((float)sc->acc0 - n);
    sc->xi = xi;
    if (xi != 0.f) {
      sc->e1 = expf(-xi * S.l);
      sc->sc = S.sigmaA * sqrtf((1.f - expf(-2.f * xi * S.l)) / (2.f * xi));
    } else {
      sc->e1 = 1.f;
      sc->sc = sqrtf(S.l) * S.sigmaA;
    }
    sc->acc0 = 0.0;
    sc->acc1 = 0.0;
  } else if (phase == 4) {
    sc->xi = sc->xi + S.l / (2.f * S.mu) * ((float)sc->acc0 - n);
    if (sc->acc1 > 0.0) { sc->status = 5; return; }
    const float kin = (float)sc->acc0 / n;
    if (U.kinetic_out) *U.kinetic_out = kin;
    sc->kin = kin;
    sc->nsampled += 1;
    sc->t += 1;
  } else if (phase == 5) {
    if (sc->ma_t <= S.ma_adapt_steps) sc->ma_t += 1;
  }
}

cudaError_t stream_update(const StreamUpdate& su, cudaStream_t st) {
  UpdDev U{};
  U.S = *su.S;
  U.P = su.P;
  U.theta = su.theta; U.mom = su.mom; U.minv = su.minv; U.rmsv = su.rmsv; U.g = su.g;
  U.sc = su.sc;
  U.key.seed = su.seed;
  U.key.chain = su.chain;
  U.gstep = su.gstep;
  U.inj = su.inj;
  U.sample_out = su.sample_out;
  U.kinetic_out = su.kinetic_out;
  const int grid = 592;
  switch (U.S.kind) {
    case S_SGLD:
      k_upd_sgld<<<grid, 256, 0, st>>>(U);
      k_upd_scalar<<<1, 1, 0, st>>>(U, 1);
      break;
    case S_SGNHT:
      k_upd_scalar<<<1, 1, 0, st>>>(U, 0);
      k_upd_sgnht<<<grid, 256, 0, st>>>(U);
      k_upd_scalar<<<1, 1, 0, st>>>(U, 2);
      break;
    case S_SGNHTS:
      if (U.S.madapter_kind == MA_RMSPROP) {
        k_upd_rmsprop<<<grid, 256, 0, st>>>(U);
        k_upd_scalar<<<1, 1, 0, st>>>(U, 5);
      }
      k_upd_scalar<<<1, 1, 0, st>>>(U, 0);
      k_upd_sgnhts_a<<<grid, 256, 0, st>>>(U);
      k_upd_scalar<<<1, 1, 0, st>>>(U, 3);
      k_upd_sgnhts_b<<<grid, 256, 0, st>>>(U);
      k_upd_scalar<<<1, 1, 0, st>>>(U, 4);
      break;
    default:
      return cudaErrorNotSupported;
  }
  return cudaGetLastError();
}

}  // namespace bfx
