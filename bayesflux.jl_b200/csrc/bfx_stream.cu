// Kernels of the large-P ("stream") regime, see bfx_stream.cuh.
#include "bfx_stream.cuh"

#include <algorithm>

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


// =====================================================================================================
// deterministic grid-wide reductions: every block leaves its partial sums in sc->part[block], the last block to
// arrive folds them in block order.  The ticket wraps to zero by itself (atomicInc), so consecutive kernels on the
// stream reuse it.  Results are bit-identical from run to run and across the replicas of a sharded chain.
// =====================================================================================================
__device__ __forceinline__ bool grid_last(unsigned* ticket) {
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;
  }
  __syncthreads();
  return s_last != 0;
}

__device__ __forceinline__ bool grid_sum2(double v0, double v1, StreamScalars* sc, unsigned* ticket, double (&out)[2]) {
  __shared__ double r0[32], r1[32], tot[2];
  for (int o = 16; o > 0; o >>= 1) {
    v0 += __shfl_xor_sync(0xffffffffu, v0, o);
    v1 += __shfl_xor_sync(0xffffffffu, v1, o);
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { r0[w] = v0; r1[w] = v1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double d0 = 0, d1 = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { d0 += r0[i]; d1 += r1[i]; }
    sc->part[blockIdx.x][0] = d0;
    sc->part[blockIdx.x][1] = d1;
  }
  if (!grid_last(ticket)) return false;
  __threadfence();
  if (w == 0) {
    // all loads first (independent, in flight together), then a fixed-order sum: lane l owns slots l, l + 32, ...
    constexpr int NS = (STREAM_RED_BLOCKS + 31) / 32;
    double a0[NS], a1[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      const int i = lane + 32 * j;
      const bool ok = i < (int)gridDim.x;
      a0[j] = ok ? __ldcg(&sc->part[i][0]) : 0.0;
      a1[j] = ok ? __ldcg(&sc->part[i][1]) : 0.0;
    }
    double d0 = 0, d1 = 0;
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      d0 += a0[j];
      d1 += a1[j];
    }
    for (int o = 16; o > 0; o >>= 1) {
      d0 += __shfl_xor_sync(0xffffffffu, d0, o);
      d1 += __shfl_xor_sync(0xffffffffu, d1, o);
    }
    if (lane == 0) { tot[0] = d0; tot[1] = d1; }
  }
  __syncthreads();
  out[0] = tot[0];
  out[1] = tot[1];
  return true;
}

// permutation key of (seed, chain-or-shared, epoch): same derivation as perm_keys in bfx_kernels_impl.cuh
__device__ __forceinline__ void stream_perm_keys(unsigned long long seed, unsigned chainkey, unsigned long long epoch,
                                                 unsigned& k0, unsigned& k1) {
  const uint4 r = philox4x32_10(make_uint4((unsigned)epoch, (unsigned)(epoch >> 32), chainkey, 0x5eedba7cu),
                                make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
  k0 = r.x;
  k1 = r.y;
}

// the observations of this evaluation; returns the number of valid columns (a scheduled batch may be the short
// last one of its epoch)
__device__ __forceinline__ long long resolve_batch(const StreamBatch& b, const StreamScalars* sc, BatchSrc& bs) {
  bs.idx = b.idx;
  bs.n = b.n;
  bs.full = b.full;
  bs.shuffle = b.shuffle;
  bs.first = 0;
  bs.k0 = bs.k1 = 0;
  if (!b.scheduled) return b.B;
  const unsigned long long g = b.gstep >= 0 ? (unsigned long long)b.gstep : (unsigned long long)sc->gstep;
  const unsigned long long epoch = g / (unsigned long long)b.nbat, k = g % (unsigned long long)b.nbat;
  bs.first = (long long)k * b.Bsched;
  const long long rem = b.n - bs.first;
  if (b.shuffle) stream_perm_keys(b.seed, b.chainkey, epoch, bs.k0, bs.k1);
  return rem < b.Bsched ? rem : b.Bsched;
}

__global__ void k_gather(const float* __restrict__ x, const float* __restrict__ y, StreamBatch b, const StreamScalars* sc,
                         long long c0, int nc, int d, float* __restrict__ X, float* __restrict__ yb,
                         __nv_bfloat16* __restrict__ xhi, __nv_bfloat16* __restrict__ xlo) {
  // one warp per column: the d features of an observation are contiguous in x
  const int lane = threadIdx.x & 31;
  const long long col = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (col >= nc) return;
  BatchSrc bs;
  const long long Bk = resolve_batch(b, sc, bs);
  const bool valid = c0 + col < Bk;  // columns past a short batch are zero (they carry no likelihood weight either)
  const long long o = valid ? batch_obs(bs, c0 + col) : 0;
  for (int f = lane; f < d; f += 32) {
    const float v = valid ? __ldg(x + o * d + f) : 0.f;
    if (X) X[col * d + f] = v;
    if (xhi) {
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      xhi[col * d + f] = h;
      xlo[col * d + f] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
  if (lane == 0) yb[col] = valid ? __ldg(y + o) : 0.f;
}

// likelihood on the network output (FP32 path; the tensor-core path fuses this into k_head)
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
  double tot[2];
  if (grid_sum2((double)s0, (double)s1, sc, &sc->ticket2, tot) && threadIdx.x == 0) {
    sc->ds0 += tot[0];
    sc->ds1 += tot[1];
    sc->bcount += (double)nc;
  }
}

// -----------------------------------------------------------------------------------------------------
// Fused head of the tensor-core path: the 1-wide output layer on BF16-split activations, the likelihood, and
// the delta of the last hidden layer.  Per batch column b (one warp per column, 8 columns per warp, 64 per block):
//   yhat = act(w . a_b + b0),  dl = num_batches * dloglike/dyhat * act'(yhat)        (rows a8 / a9 of SURVEY.md 8a)
//   D[i, b] = w[i] * dl * act_prev'(a[i, b])                                         (BF16 halves, for the GEMMs)
// and per block the partial sums  dW_out[i] = sum_b dl a[i, b],  rs[i] = sum_b D[i, b] (bias gradient of the last
// hidden layer),  db_out = sum_b dl  into part[block][2 nin + 1]; the likelihood sums go through the grid reduction.
// -----------------------------------------------------------------------------------------------------
constexpr int HEAD_COLS = 32, HEAD_CPW = 4, HEAD_R = 8;  // 8 warps x 4 columns; lane l owns i = 2 l + 64 r + {0, 1}, r < HEAD_R

__host__ __device__ inline long long head_part_stride(int nin) { return 2LL * nin + 4; }  // [dW_out | row sums | db_out, pad]

__global__ void __launch_bounds__(256) k_head(StreamModel M, const float* __restrict__ theta, const __nv_bfloat16* __restrict__ ahi,
                                              const __nv_bfloat16* __restrict__ alo, const float* __restrict__ yb, int nin, int nc,
                                              long long nvalid_host, StreamBatch bdesc, long long c0, float nb, int out_act,
                                              int act_prev, long long w_off, long long b_off, int want_grad, float* __restrict__ yhat,
                                              __nv_bfloat16* __restrict__ dhi, __nv_bfloat16* __restrict__ dlo,
                                              float* __restrict__ part, StreamScalars* sc) {
  extern __shared__ float red[];  // [8 warps][2 * 512]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const float* wv = theta + w_off;
  const float b0 = theta[b_off];
  const float inv_sigma = 1.f / expf(theta[M.P - 1]);
  long long nvalid = nvalid_host;
  if (bdesc.scheduled) {  // a scheduled batch may be the short last one of its epoch
    BatchSrc bs;
    nvalid = resolve_batch(bdesc, sc, bs) - c0;
  }
  const int bw = blockIdx.x * HEAD_COLS + HEAD_CPW * w;  // first column of this warp
  long long colbase[HEAD_CPW];
#pragma unroll
  for (int j = 0; j < HEAD_CPW; ++j) colbase[j] = (long long)(bw + j < nc ? bw + j : nc - 1) * nin;  // clamped: loads stay in bounds
  float dlc[HEAD_CPW], s0 = 0.f, s1 = 0.f, sdl = 0.f;
  // ---- phase A: network output and likelihood delta of this warp's columns (their loads in flight together)
  {
    float acc[HEAD_CPW];
#pragma unroll
    for (int j = 0; j < HEAD_CPW; ++j) acc[j] = 0.f;
    for (int i2 = lane; i2 < nin / 2; i2 += 32) {
      const float2 ww = *reinterpret_cast<const float2*>(wv + 2 * i2);
      float2 h[HEAD_CPW], l[HEAD_CPW];
#pragma unroll
      for (int j = 0; j < HEAD_CPW; ++j) {
        h[j] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(ahi + colbase[j] + 2 * i2));
        l[j] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(alo + colbase[j] + 2 * i2));
      }
#pragma unroll
      for (int j = 0; j < HEAD_CPW; ++j) {
        acc[j] = fmaf(ww.x, h[j].x + l[j].x, acc[j]);
        acc[j] = fmaf(ww.y, h[j].y + l[j].y, acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < HEAD_CPW; ++j) {
      float a = acc[j];
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      const int b = bw + j;
      dlc[j] = 0.f;
      if (b < nc) {
        const float yh = act_apply(out_act, a + b0);
        if (lane == 0 && yhat) yhat[b] = yh;
        if (b < nvalid) {
          const float z = (yb[b] - yh) * inv_sigma;
          float d;
          if (M.like_kind == LK_NORMAL) {
            s0 += z * z;
            d = nb * z * inv_sigma;
          } else {
            const float zz = z * z;
            const float tw = (M.nu + 1.f) * z / (M.nu + zz);
            s0 += log1pf(zz / M.nu);
            s1 += tw * z;
            d = nb * tw * inv_sigma;
          }
          dlc[j] = d * act_deriv_from_out(out_act, yh);
          sdl += dlc[j];
        }
      }
    }
  }
  // ---- phase B: delta of the last hidden layer and the partial gradient sums, 512 inputs at a time
  if (want_grad) {
    float* mypart = part + (long long)blockIdx.x * head_part_stride(nin);
    for (int ic = 0; ic < nin; ic += 64 * HEAD_R) {
      float dw[2 * HEAD_R], rs[2 * HEAD_R];
#pragma unroll
      for (int r = 0; r < HEAD_R; ++r) {
        dw[2 * r] = dw[2 * r + 1] = rs[2 * r] = rs[2 * r + 1] = 0.f;
        const int i = ic + 2 * lane + 64 * r;
        if (i < nin) {
          const float2 ww = *reinterpret_cast<const float2*>(wv + i);
          float2 h[HEAD_CPW], l[HEAD_CPW];
#pragma unroll
          for (int j = 0; j < HEAD_CPW; ++j) {
            h[j] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(ahi + colbase[j] + i));
            l[j] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(alo + colbase[j] + i));
          }
#pragma unroll
          for (int j = 0; j < HEAD_CPW; ++j) {
            if (bw + j < nc) {
              const float dl = dlc[j];
              const float a0 = h[j].x + l[j].x, a1 = h[j].y + l[j].y;
              const float d0 = ww.x * dl * act_deriv_from_out(act_prev, a0), d1 = ww.y * dl * act_deriv_from_out(act_prev, a1);
              const __nv_bfloat162 dh = __floats2bfloat162_rn(d0, d1);
              const float2 dhf = __bfloat1622float2(dh);
              *reinterpret_cast<__nv_bfloat162*>(dhi + colbase[j] + i) = dh;
              *reinterpret_cast<__nv_bfloat162*>(dlo + colbase[j] + i) = __floats2bfloat162_rn(d0 - dhf.x, d1 - dhf.y);
              dw[2 * r] = fmaf(dl, a0, dw[2 * r]);
              dw[2 * r + 1] = fmaf(dl, a1, dw[2 * r + 1]);
              rs[2 * r] += d0;
              rs[2 * r + 1] += d1;
            }
          }
        }
      }
      // cross-warp sums in warp order
#pragma unroll
      for (int r = 0; r < HEAD_R; ++r) {
        const int il = 2 * lane + 64 * r;
        *reinterpret_cast<float2*>(red + w * 1024 + il) = make_float2(dw[2 * r], dw[2 * r + 1]);
        *reinterpret_cast<float2*>(red + w * 1024 + 512 + il) = make_float2(rs[2 * r], rs[2 * r + 1]);
      }
      __syncthreads();
      for (int e = threadIdx.x; e < 1024; e += 256) {
        const int il = e & 511, which = e >> 9;
        if (ic + il < nin) {
          float s = 0.f;
#pragma unroll
          for (int ww2 = 0; ww2 < 8; ++ww2) s += red[ww2 * 1024 + e];
          mypart[which * nin + ic + il] = s;
        }
      }
      __syncthreads();
    }
    // db_out of the block (every lane of a warp carries the same dl values)
    if (lane == 0) red[w] = sdl;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int i = 0; i < 8; ++i) s += red[i];
      mypart[2 * nin] = s;
    }
  }
  // likelihood sums: identical on the 32 lanes of a warp (all lanes ran the same scalar code), lane 0 contributes
  double tot[2];
  const bool mine = lane == 0;
  if (grid_sum2(mine ? (double)s0 : 0.0, mine ? (double)s1 : 0.0, sc, &sc->ticket2, tot) && threadIdx.x == 0) {
    sc->ds0 += tot[0];
    sc->ds1 += tot[1];
    const long long v = nvalid < (long long)nc ? (nvalid < 0 ? 0 : nvalid) : (long long)nc;  // valid columns of the chunk
    sc->bcount += (double)v;
  }
}

__global__ void k_zero(float* g, long long P, StreamScalars* sc) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < P; i += (long long)gridDim.x * blockDim.x) g[i] = 0.f;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    sc->ds0 = sc->ds1 = sc->bcount = 0.0;
    sc->th2 = sc->gn2 = sc->acc0 = sc->acc1 = 0.0;
  }
}

__global__ void k_pack_sums(float* g, long long P, const StreamScalars* sc) {
  g[P] = (float)sc->ds0;
  g[P + 1] = (float)sc->ds1;
  g[P + 2] = (float)sc->bcount;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v);
// the same, then the local gradient is complete: publish the epoch in every peer's ready[self] (one launch less)
__global__ void k_pack_sums_signal(float* g, long long P, const StreamScalars* sc, const __grid_constant__ PeerDev pd) {
  if (threadIdx.x == 0) {
    g[P] = (float)sc->ds0;
    g[P + 1] = (float)sc->ds1;
    g[P + 2] = (float)sc->bcount;
    __threadfence_system();
  }
  __syncthreads();
  const int r = threadIdx.x;
  if (r < pd.n && r != pd.self) st_release_sys(pd.flg[r] + PEER_READY + pd.self, pd.flg[pd.self][PEER_EPOCH] + 1ull);
}

__global__ void k_set_gstep(StreamScalars* sc, long long gstep) { sc->gstep = gstep; }

// prior, sigma prior, d/d theta_like, the value and the norms in one pass:
//   g_net -= c2 * theta_net;  g_like = glike;  value = nb * loglike + logprior
// (PEER: the gradient and the likelihood sums are the rank-ordered sums over all ranks' exchange buffers, read over
// NVLink behind the peers' ready flags; the result goes to the local buffer g)
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// thread 0 of the block waits until word[j] >= e for every peer j; false after about a minute (a peer died; ranks that
// merely enter their first update seconds apart must not trip it): the caller flags the chain instead of hanging
__device__ __forceinline__ bool peer_wait(const unsigned long long* words, int n, int self, unsigned long long e) {
  __shared__ int s_ok;
  if (threadIdx.x == 0) {
    int ok = 1;
    const long long t0 = clock64();
    for (int j = 0; j < n && ok; ++j) {
      if (j == self) continue;
      while (ld_acquire_sys(words + j) < e) {
        if (clock64() - t0 > 120000000000LL) {
          ok = 0;
          break;
        }
        __nanosleep(64);
      }
    }
    s_ok = ok;
  }
  __syncthreads();
  return s_ok != 0;
}

template <bool PEER>
__device__ __forceinline__ void finish_body(const StreamModel& M, const float* __restrict__ theta, float* __restrict__ g,
                                            StreamScalars* sc, float nb, int want_grad, int from_buffer, const PeerDev& pd) {
  double ds0, ds1, B;
  unsigned long long epoch = 0;
  if (PEER) {
    epoch = pd.flg[pd.self][PEER_EPOCH] + 1ull;
    if (!peer_wait(pd.flg[pd.self] + PEER_READY, pd.n, pd.self, epoch)) {
      if (grid_last(&sc->ticket) && threadIdx.x == 0) sc->status = 8;  // peer exchange timed out
      return;
    }
    // the three likelihood sums: one thread per block fetches the ranks' tails (remote reads), the block shares them
    __shared__ double s_sum[3];
    if (threadIdx.x == 0) {
      double a = 0.0, b = 0.0, c = 0.0;
      for (int r = 0; r < pd.n; ++r) {
        a += (double)__ldcg(pd.buf[r] + M.P);
        b += (double)__ldcg(pd.buf[r] + M.P + 1);
        c += (double)__ldcg(pd.buf[r] + M.P + 2);
      }
      s_sum[0] = a; s_sum[1] = b; s_sum[2] = c;
    }
    __syncthreads();
    ds0 = s_sum[0]; ds1 = s_sum[1]; B = s_sum[2];
  } else if (from_buffer) {
    ds0 = (double)g[M.P];
    ds1 = (double)g[M.P + 1];
    B = (double)g[M.P + 2];
  } else {
    ds0 = sc->ds0;
    ds1 = sc->ds1;
    B = sc->bcount;
  }
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
  double like, dlike;
  if (M.like_kind == LK_NORMAL) {
    like = -0.5 * B * 1.8378770664093453 - 0.5 * ds0;
    dlike = ds0;
  } else {
    like = B * M.tdist_const - 0.5 * ((double)M.nu + 1.0) * ds0;
    dlike = ds1;
  }
  like += -B * (double)sl + lps;
  dlike += -B + dlps;
  const float glike = (float)((double)nb * dlike);
  float a0 = 0.f, a1 = 0.f;
  const float c2 = M.prior_c2;
  const long long nq = (M.P + 3) >> 2;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < nq; q += (long long)gridDim.x * blockDim.x) {
    const long long i = q << 2;
    const int cnt = M.P - i < 4 ? (int)(M.P - i) : 4;
    float th[4], gg[4];
    if (cnt == 4) {
      const float4 t = *reinterpret_cast<const float4*>(theta + i);
      th[0] = t.x; th[1] = t.y; th[2] = t.z; th[3] = t.w;
      if (PEER) {
        float4 gv = __ldcg(reinterpret_cast<const float4*>(pd.buf[0] + i));
        for (int r = 1; r < pd.n; ++r) {
          const float4 o = __ldcg(reinterpret_cast<const float4*>(pd.buf[r] + i));
          gv.x += o.x; gv.y += o.y; gv.z += o.z; gv.w += o.w;
        }
        gg[0] = gv.x; gg[1] = gv.y; gg[2] = gv.z; gg[3] = gv.w;
      } else {
        const float4 gv = *reinterpret_cast<const float4*>(g + i);
        gg[0] = gv.x; gg[1] = gv.y; gg[2] = gv.z; gg[3] = gv.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        th[j] = j < cnt ? theta[i + j] : 0.f;
        if (PEER) {
          float a = 0.f;
          if (j < cnt)
            for (int r = 0; r < pd.n; ++r) a += __ldcg(pd.buf[r] + i + j);
          gg[j] = a;
        } else {
          gg[j] = j < cnt ? g[i + j] : 0.f;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < cnt) {
        if (i + j < M.Pnet) {
          a0 += th[j] * th[j];
          gg[j] = gg[j] - c2 * th[j];
        } else {
          gg[j] = glike;
        }
        if (want_grad) a1 += gg[j] * gg[j];
      }
    }
    if (want_grad) {
      if (cnt == 4) {
        *reinterpret_cast<float4*>(g + i) = make_float4(gg[0], gg[1], gg[2], gg[3]);
      } else {
        for (int j = 0; j < cnt; ++j) g[i + j] = gg[j];
      }
    }
  }
  double tot[2];
  if (grid_sum2((double)a0, (double)a1, sc, &sc->ticket, tot) && threadIdx.x == 0) {
    sc->th2 = tot[0];
    sc->gn2 = tot[1];
    sc->glike = glike;
    sc->value = (double)nb * like + M.prior_c0n - 0.5 * (double)M.prior_c2 * tot[0];
    if (from_buffer || PEER) {
      sc->ds0 = ds0;
      sc->ds1 = ds1;
      sc->bcount = B;
    }
    if (PEER) {  // every block of this grid has finished reading the peers' buffers (it arrived at the reduction)
      __threadfence_system();
      for (int r = 0; r < pd.n; ++r)
        if (r != pd.self) st_release_sys(pd.flg[r] + PEER_DONE + pd.self, epoch);
      pd.flg[pd.self][PEER_EPOCH] = epoch;
    }
  }
}

__global__ void k_finish(StreamModel M, const float* __restrict__ theta, float* __restrict__ g, StreamScalars* sc, float nb,
                         int want_grad, int from_buffer) {
  finish_body<false>(M, theta, g, sc, nb, want_grad, from_buffer, PeerDev{});
}
__global__ void k_finish_peer(StreamModel M, const float* __restrict__ theta, float* __restrict__ g, StreamScalars* sc,
                              float nb, const __grid_constant__ PeerDev pd) {
  finish_body<true>(M, theta, g, sc, nb, 1, 1, pd);
}
// ---- SCATTER protocol (bfx_stream.cuh) --------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_finish_scatter(StreamModel M, const float* __restrict__ theta, StreamScalars* sc,
                                                        float nb, const __grid_constant__ PeerDev pd) {
  const unsigned long long epoch = pd.flg[pd.self][PEER_EPOCH] + 1ull;
  if (!peer_wait(pd.flg[pd.self] + PEER_READY, pd.n, pd.self, epoch)) {
    if (grid_last(&sc->ticket) && threadIdx.x == 0) sc->status = 8;  // peer exchange timed out
    return;
  }
  __shared__ double s_sum[3];
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0, c = 0.0;
    for (int r = 0; r < pd.n; ++r) {
      a += (double)__ldcg(pd.buf[r] + M.P);
      b += (double)__ldcg(pd.buf[r] + M.P + 1);
      c += (double)__ldcg(pd.buf[r] + M.P + 2);
    }
    s_sum[0] = a; s_sum[1] = b; s_sum[2] = c;
  }
  __syncthreads();
  const double ds0 = s_sum[0], ds1 = s_sum[1], B = s_sum[2];
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
  double like, dlike;
  if (M.like_kind == LK_NORMAL) {
    like = -0.5 * B * 1.8378770664093453 - 0.5 * ds0;
    dlike = ds0;
  } else {
    like = B * M.tdist_const - 0.5 * ((double)M.nu + 1.0) * ds0;
    dlike = ds1;
  }
  like += -B * (double)sl + lps;
  dlike += -B + dlps;
  const float glike = (float)((double)nb * dlike);
  // this rank's slice of the quads
  const long long nq = (M.P + 3) >> 2, per = (nq + pd.n - 1) / pd.n;
  const long long q0 = per * pd.self, q1 = (q0 + per < nq) ? q0 + per : nq;
  float a0 = 0.f, a1 = 0.f;
  const float c2 = M.prior_c2;
  for (long long q = q0 + blockIdx.x * (long long)blockDim.x + threadIdx.x; q < q1; q += (long long)gridDim.x * blockDim.x) {
    const long long i = q << 2;
    const int cnt = M.P - i < 4 ? (int)(M.P - i) : 4;
    float th[4], gg[4];
    if (cnt == 4) {
      const float4 t = *reinterpret_cast<const float4*>(theta + i);
      th[0] = t.x; th[1] = t.y; th[2] = t.z; th[3] = t.w;
      float4 gv = __ldcg(reinterpret_cast<const float4*>(pd.buf[0] + i));
      for (int r = 1; r < pd.n; ++r) {
        const float4 o = __ldcg(reinterpret_cast<const float4*>(pd.buf[r] + i));
        gv.x += o.x; gv.y += o.y; gv.z += o.z; gv.w += o.w;
      }
      gg[0] = gv.x; gg[1] = gv.y; gg[2] = gv.z; gg[3] = gv.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        th[j] = j < cnt ? theta[i + j] : 0.f;
        float a = 0.f;
        if (j < cnt)
          for (int r = 0; r < pd.n; ++r) a += __ldcg(pd.buf[r] + i + j);
        gg[j] = a;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < cnt) {
        if (i + j < M.Pnet) {
          a0 += th[j] * th[j];
          gg[j] = gg[j] - c2 * th[j];
        } else {
          gg[j] = glike;
        }
        a1 += gg[j] * gg[j];
      }
    }
    for (int r = 0; r < pd.n; ++r) {  // the finished slice goes to every rank's result buffer (remote stores)
      if (cnt == 4) {
        *reinterpret_cast<float4*>(pd.sum[r] + i) = make_float4(gg[0], gg[1], gg[2], gg[3]);
      } else {
        for (int j = 0; j < cnt; ++j) pd.sum[r][i + j] = gg[j];
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) __threadfence_system();  // this block's remote stores are ordered before the flag below
  double tot[2];
  if (grid_sum2((double)a0, (double)a1, sc, &sc->ticket, tot) && threadIdx.x == 0) {
    sc->glike = glike;
    sc->value = (double)nb * like + M.prior_c0n;  // (the join subtracts the prior's quadratic term)
    sc->ds0 = ds0;
    sc->ds1 = ds1;
    sc->bcount = B;
    for (int r = 0; r < pd.n; ++r) {
      double* nr = reinterpret_cast<double*>(pd.flg[r] + PEER_NORMS);
      nr[2 * pd.self] = tot[0];
      nr[2 * pd.self + 1] = tot[1];
    }
    __threadfence_system();
    for (int r = 0; r < pd.n; ++r) st_release_sys(pd.flg[r] + PEER_SLICED + pd.self, epoch);
    pd.flg[pd.self][PEER_EPOCH] = epoch;
  }
}
// one block: every slice has arrived (and every peer has finished reading this rank's buffer); norms in rank order
__global__ void k_peer_join(StreamModel M, StreamScalars* sc, const __grid_constant__ PeerDev pd) {
  const unsigned long long epoch = pd.flg[pd.self][PEER_EPOCH];
  const bool ok = peer_wait(pd.flg[pd.self] + PEER_SLICED, pd.n, pd.self, epoch);
  if (threadIdx.x == 0) {
    if (!ok) {
      sc->status = 8;
    } else {
      const double* nr = reinterpret_cast<const double*>(pd.flg[pd.self] + PEER_NORMS);
      double th2 = 0.0, gn2 = 0.0;
      for (int r = 0; r < pd.n; ++r) {
        th2 += __ldcg(nr + 2 * r);
        gn2 += __ldcg(nr + 2 * r + 1);
      }
      sc->th2 = th2;
      sc->gn2 = gn2;
      sc->value -= 0.5 * (double)M.prior_c2 * th2;
    }
  }
}

// behind the local gradient (stream order): publish the epoch in every peer's ready[self]
__global__ void k_peer_signal(const __grid_constant__ PeerDev pd) {
  const int r = threadIdx.x;
  if (r < pd.n && r != pd.self) {
    const unsigned long long e = pd.flg[pd.self][PEER_EPOCH] + 1ull;
    __threadfence_system();
    st_release_sys(pd.flg[r] + PEER_READY + pd.self, e);
  }
}
// k_zero behind the peers' done flags: nobody reads this rank's buffer of the previous epoch any more
__global__ void k_zero_peer(float* g, long long P, StreamScalars* sc, const __grid_constant__ PeerDev pd) {
  const unsigned long long e = pd.flg[pd.self][PEER_EPOCH];
  const bool ok = peer_wait(pd.flg[pd.self] + PEER_DONE, pd.n, pd.self, e);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < P; i += (long long)gridDim.x * blockDim.x) g[i] = 0.f;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    sc->ds0 = sc->ds1 = sc->bcount = 0.0;
    sc->th2 = sc->gn2 = sc->acc0 = sc->acc1 = 0.0;
    if (!ok) sc->status = 8;
  }
}



// g[dst + e] += sum_z src[z * stride + e] for a list of vectors (split-K partials of the weight gradients, bias-
// gradient row sums, the head's partial sums), summed in slot order
struct ReduceItem {
  long long dst, count, stride;
  const float* src;
  int nparts;
};
struct ReduceList {
  int n;
  long long total;
  ReduceItem it[3 * BFX_MAX_LAYERS + 4];
};

// (the list is a __grid_constant__: indexing a by-value parameter dynamically would copy it to local memory per thread)
__global__ void __launch_bounds__(256) k_reduce_list(float* __restrict__ g, const __grid_constant__ ReduceList R) {
  const ReduceItem& I = R.it[blockIdx.y];
  const bool vec = (I.count & 3) == 0 && (I.stride & 3) == 0 && (I.dst & 3) == 0 &&
                   (reinterpret_cast<unsigned long long>(I.src) & 15ull) == 0;
  if (vec && I.nparts >= 32) {
    // many partials of a short vector (bias row sums: 64 slots, the head's partial sums: one per 32 columns): a thread
    // walking all of them is a chain of dependent memory round trips, so eight threads share a quad (slots g, g + 8,
    // ...), and their sums are folded in group order through shared memory
    __shared__ float4 sh[8][32];
    const long long nq = I.count >> 2;
    const int pg = threadIdx.x >> 5, ql = threadIdx.x & 31;
    const long long sq = I.stride >> 2;
    for (long long q0 = (long long)blockIdx.x * 32; q0 < nq; q0 += (long long)gridDim.x * 32) {
      const long long q = q0 + ql;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q < nq) {
        const float4* src = reinterpret_cast<const float4*>(I.src) + q;
        int z = pg;
        for (; z + 56 < I.nparts; z += 64) {
          float4 v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = __ldcs(src + (long long)(z + 8 * j) * sq);
#pragma unroll
          for (int j = 0; j < 8; ++j) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
        }
        for (; z < I.nparts; z += 8) {
          const float4 v = __ldcs(src + (long long)z * sq);
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
      }
      sh[pg][ql] = acc;
      __syncthreads();
      if (pg == 0 && q < nq) {
        float4 t = *reinterpret_cast<const float4*>(g + I.dst + 4 * q);
#pragma unroll
        for (int k = 0; k < 8; ++k) { t.x += sh[k][ql].x; t.y += sh[k][ql].y; t.z += sh[k][ql].z; t.w += sh[k][ql].w; }
        *reinterpret_cast<float4*>(g + I.dst + 4 * q) = t;
      }
      __syncthreads();
    }
  } else if (vec) {
    const long long nq = I.count >> 2;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < nq; q += (long long)gridDim.x * blockDim.x) {
      const float4* src = reinterpret_cast<const float4*>(I.src) + q;
      const long long sq = I.stride >> 2;
      float4 acc = *reinterpret_cast<const float4*>(g + I.dst + 4 * q);
      int z = 0;
      for (; z + 7 < I.nparts; z += 8) {  // eight independent 16-byte loads in flight, summed in slot order
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldcs(src + (long long)(z + j) * sq);
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
      }
      for (; z < I.nparts; ++z) {
        const float4 v = __ldcs(src + (long long)z * sq);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      *reinterpret_cast<float4*>(g + I.dst + 4 * q) = acc;
    }
  } else if (I.count <= 8) {
    // a handful of scalars with many partials (the head's db): one warp per element, lanes take slots l, l + 32, ...
    if (blockIdx.x != 0) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (long long r = w; r < I.count; r += blockDim.x >> 5) {
      float acc = 0.f;
      for (int z = lane; z < I.nparts; z += 32) acc += I.src[(long long)z * I.stride + r];
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) g[I.dst + r] += acc;
    }
  } else {
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < I.count; r += (long long)gridDim.x * blockDim.x) {
      float acc = g[I.dst + r];
      for (int z = 0; z < I.nparts; ++z) acc += I.src[(long long)z * I.stride + r];
      g[I.dst + r] = acc;
    }
  }
}

// =====================================================================================================
// host orchestration of one evaluation
// =====================================================================================================
static bool head_on_tc(const StreamModel& M) {  // 1-wide output layer fed by a tensor-core layer: fused head kernel
  const int nl = M.nlayers;
  return M.tc && nl >= 2 && M.L[nl - 1].nout == 1 && M.L[nl - 2].tc && M.L[nl - 1].nin % 8 == 0;
}

// distance between the split-K partials of a weight gradient: NOT the power of two the matrix size usually is (16
// partial copies of a 512 x 512 matrix exactly 1 MB apart fall onto the same L2 slices / DRAM channels and the
// list-reduce that reads them slot by slot ran at 0.6 TB/s), 16-byte aligned
static long long tc_part_stride(const StreamLayer& L) { return (long long)L.nout * L.nin + 1056; }

static int tc_ksplit_for(const StreamLayer& L, long long nc) {
  const int tiles = ((L.nout + 255) / 256) * ((L.nin + 255) / 256);
  int ks = 1;
  while (ks < 32 && tiles * ks * 2 <= 74 && (nc / (ks * 2)) >= 256) ks *= 2;
  return ks;
}

cudaError_t stream_work_alloc(const StreamModel& M, long long Bc, StreamWork& W) {
  W = StreamWork{};
  W.Bc = Bc;
  cudaError_t e = cudaSuccess;
  const int nl = M.nlayers;
  int maxw = 1;
  for (int l = 0; l < nl; ++l) maxw = M.L[l].nout > maxw ? M.L[l].nout : maxw;
  const bool htc = head_on_tc(M);
  // FP32 copies of the activations only where an FP32 kernel reads them
  // (an FP32 layer reads its input in the forward, weight-gradient and act' kernels and always writes its output)
  for (int j = 0; j < nl; ++j) W.need_f32[j] = !(M.L[j].tc || (htc && j == nl - 1)) || (j > 0 && !M.L[j - 1].tc);
  W.need_f32[nl] = 1;
  bool any_f32_delta = false;
  for (int l = 0; l < nl; ++l) any_f32_delta |= !M.L[l].tc;
  for (int j = 0; j <= nl && e == cudaSuccess; ++j) {
    const long long rows = j == 0 ? M.d_in : M.L[j - 1].nout;
    if (W.need_f32[j]) e = cudaMalloc(&W.act[j], (size_t)rows * Bc * sizeof(float));
  }
  if (any_f32_delta || !M.tc)
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaMalloc(&W.delta[i], (size_t)maxw * Bc * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&W.yb, (size_t)Bc * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&W.dl, (size_t)Bc * sizeof(float));
  // split-K partials of the FP32 weight-gradient GEMMs: up to 32 partial copies of the largest [W | b] block
  long long big = 0;
  for (int l = 0; l < nl; ++l) {
    if (M.L[l].tc) continue;
    long long c = (long long)M.L[l].nout * (M.L[l].nin + 1);
    big = c > big ? c : big;
  }
  W.part_floats = big * 32;
  if (e == cudaSuccess && big) e = cudaMalloc(&W.part, (size_t)W.part_floats * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&W.rowpart, (size_t)64 * maxw * sizeof(float));
  if (M.tc) {
    for (int j = 0; j < nl && e == cudaSuccess; ++j) {
      const long long rows = j == 0 ? M.d_in : M.L[j - 1].nout;
      // halves exist for every activation a tensor-core kernel reads
      if (!(M.L[j].tc || (htc && j == nl - 1))) continue;
      e = cudaMalloc(&W.act_hi[j], (size_t)rows * Bc * 2);
      if (e == cudaSuccess) e = cudaMalloc(&W.act_lo[j], (size_t)rows * Bc * 2);
      if (e == cudaSuccess && j > 0 && M.L[j - 1].tc && M.L[j - 1].act == A_RELU)
        e = cudaMalloc(&W.mask[j], (size_t)((Bc + 31) / 32) * rows * sizeof(unsigned));
    }
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
      e = cudaMalloc(&W.delta_hi[i], (size_t)maxw * Bc * 2);
      if (e == cudaSuccess) e = cudaMalloc(&W.delta_lo[i], (size_t)maxw * Bc * 2);
    }
    if (e == cudaSuccess) e = cudaMalloc(&W.w_hi, (size_t)(M.P + 8) * 2);
    if (e == cudaSuccess) e = cudaMalloc(&W.w_lo, (size_t)(M.P + 8) * 2);
    for (int l = 0; l < nl && e == cudaSuccess; ++l) {
      if (!M.L[l].tc) continue;
      W.tc_ksplit[l] = tc_ksplit_for(M.L[l], Bc);
      e = cudaMalloc(&W.tc_part[l], (size_t)W.tc_ksplit[l] * tc_part_stride(M.L[l]) * sizeof(float));
      W.tc_rowslots[l] = tcg::ROWSLOTS * (int)((Bc + 223) / 224);  // (tiles are 224 or 256 columns wide: tcg::tile_cols_kmajor_b)
      if (e == cudaSuccess) e = cudaMalloc(&W.tc_rowpart[l], (size_t)W.tc_rowslots[l] * M.L[l].nout * sizeof(float));
    }
    if (htc && e == cudaSuccess) {
      W.head_blocks = (int)((Bc + HEAD_COLS - 1) / HEAD_COLS);
      e = cudaMalloc(&W.head_part, (size_t)W.head_blocks * head_part_stride(M.L[nl - 1].nin) * sizeof(float));
    }
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
    cudaFree(W.mask[l]);
  }
  for (int l = 0; l < BFX_MAX_LAYERS; ++l) {
    cudaFree(W.tc_part[l]);
    cudaFree(W.tc_rowpart[l]);
  }
  for (int i = 0; i < 2; ++i) {
    cudaFree(W.delta_hi[i]);
    cudaFree(W.delta_lo[i]);
  }
  cudaFree(W.w_hi);
  cudaFree(W.w_lo);
  cudaFree(W.head_part);
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

cudaError_t stream_zero_peer(float* g, long long P, StreamScalars* sc, const PeerDev& pd, cudaStream_t st) {
  k_zero_peer<<<296, 256, 0, st>>>(g, P, sc, pd);
  return cudaGetLastError();
}
cudaError_t stream_peer_signal(const PeerDev& pd, cudaStream_t st) {
  k_peer_signal<<<1, 32, 0, st>>>(pd);
  return cudaGetLastError();
}
cudaError_t stream_finish_scatter(const StreamModel& M, const float* theta, StreamScalars* sc, float nb, const PeerDev& pd,
                                  cudaStream_t st) {
  const long long nq = (M.P + 3) >> 2, per = (nq + pd.n - 1) / pd.n;
  long long grid = (per + 255) / 256;
  if (grid > STREAM_RED_BLOCKS) grid = STREAM_RED_BLOCKS;
  if (grid < 1) grid = 1;
  k_finish_scatter<<<(unsigned)grid, 256, 0, st>>>(M, theta, sc, nb, pd);
  k_peer_join<<<1, 32, 0, st>>>(M, sc, pd);
  return cudaGetLastError();
}
cudaError_t stream_pack_sums_signal(float* g, long long P, StreamScalars* sc, const PeerDev& pd, cudaStream_t st) {
  k_pack_sums_signal<<<1, 32, 0, st>>>(g, P, sc, pd);
  return cudaGetLastError();
}
cudaError_t stream_finish_peer(const StreamModel& M, const float* theta, float* g_out, StreamScalars* sc, float nb,
                               const PeerDev& pd, cudaStream_t st) {
  k_finish_peer<<<STREAM_RED_BLOCKS, 256, 0, st>>>(M, theta, g_out, sc, nb, pd);
  return cudaGetLastError();
}

cudaError_t stream_pack_sums(float* g, long long P, StreamScalars* sc, cudaStream_t st) {
  k_pack_sums<<<1, 1, 0, st>>>(g, P, sc);
  return cudaGetLastError();
}

cudaError_t stream_split_theta(const StreamModel& M, StreamWork& W, const float* theta, cudaStream_t st) {
  if (!M.tc || !W.w_hi) return cudaSuccess;
  k_split<<<592, 256, 0, st>>>(theta, M.P, W.w_hi, W.w_lo);
  return cudaGetLastError();
}

cudaError_t stream_set_gstep(StreamScalars* sc, long long gstep, cudaStream_t st) {
  k_set_gstep<<<1, 1, 0, st>>>(sc, gstep);
  return cudaGetLastError();
}

cudaError_t stream_eval_like(const StreamModel& M, StreamWork& W, const float* theta, const float* x, const float* y,
                             const StreamBatch& b, float nb, bool want_grad, float* g, StreamScalars* sc,
                             cudaStream_t st, float* yhat_out, bool theta_split_valid, StreamBucketHook* hook) {
  const int nl = M.nlayers;
  const bool htc = head_on_tc(M);
  // bucketed exchange: every layer below the fused head on the tensor cores, one batch chunk
  bool bucketed = hook != nullptr && want_grad && htc && b.B <= W.Bc;
  for (int l = 0; l < nl - 1; ++l) bucketed = bucketed && M.L[l].tc;
  if (hook) hook->used = bucketed ? 1 : 0;
  int nbuckets = 0;
  if (M.tc && !theta_split_valid) {  // BF16 halves of theta for the tensor-core layers
    k_split<<<592, 256, 0, st>>>(theta, M.P, W.w_hi, W.w_lo);
    SCK(cudaGetLastError());
  }
  for (long long c0 = 0; c0 < b.B; c0 += W.Bc) {
    const int nc = (int)((b.B - c0) < W.Bc ? (b.B - c0) : W.Bc);
    {
      const long long threads = (long long)nc * 32;
      k_gather<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(x, y, b, sc, c0, nc, M.d_in, W.act[0], W.yb, W.act_hi[0],
                                                               W.act_lo[0]);
      SCK(cudaGetLastError());
    }
    // forward: act[l+1] (nout x nc) = act(W_l act[l] + b_l)
    for (int l = 0; l < nl; ++l) {
      const StreamLayer& L = M.L[l];
      if (L.tc) {
        tcg::Args t{};
        t.M = L.nout; t.N = nc; t.K = L.nin;
        t.C = W.need_f32[l + 1] ? W.act[l + 1] : nullptr; t.ldc = L.nout;
        t.bias = theta + L.b_off; t.act = L.act;
        t.mask_out = W.mask[l + 1];
        t.kchunk = (L.nin + tcg::BK - 1) / tcg::BK * tcg::BK; t.part_stride = 0;
        SCK(tc2_forward(t, tcg::Operand{W.w_hi + L.w_off, W.w_lo + L.w_off, L.nout}, tcg::Operand{W.act_hi[l], W.act_lo[l], L.nin},
                        tcg::Output{W.act_hi[l + 1], W.act_lo[l + 1]}, st));
        continue;
      }
      if (l == nl - 1 && htc) {
        k_head<<<(nc + HEAD_COLS - 1) / HEAD_COLS, 256, 8 * 1024 * sizeof(float), st>>>(
            M, theta, W.act_hi[l], W.act_lo[l], W.yb, L.nin, nc, b.B - c0, b, c0, nb, L.act, M.L[l - 1].act, L.w_off, L.b_off,
            want_grad ? 1 : 0, W.act[nl], W.delta_hi[0], W.delta_lo[0], W.head_part, sc);
        SCK(cudaGetLastError());
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
      if (W.act_hi[l + 1]) { a.chi = W.act_hi[l + 1]; a.clo = W.act_lo[l + 1]; }
      SCK((launch_gemm<true, true, EPI_BIAS_ACT>(a, 1, st)));
    }
    if (!htc) {
      k_like<<<148, 256, 0, st>>>(M, theta, W.act[nl], W.yb, nc, nb, M.L[nl - 1].act, W.dl, sc);
      SCK(cudaGetLastError());
    }
    if (yhat_out) SCK(cudaMemcpyAsync(yhat_out + c0, W.act[nl], (size_t)nc * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (!want_grad) continue;
    // backward: delta of layer l in D (nout x nc); the output layer's delta is dl (1 x nc)
    const float* D = W.dl;
    const __nv_bfloat16 *Dhi = nullptr, *Dlo = nullptr;
    int ping = 0;
    int bias_slots = 0;  // > 0: the producer of the current delta left bias-gradient partials in W.tc_rowpart[l]
    ReduceList R{};
    auto add_item = [&](long long dst, long long count, const float* src, int nparts, long long stride) {
      ReduceItem& I = R.it[R.n++];
      I.dst = dst; I.count = count; I.src = src; I.nparts = nparts; I.stride = stride;
      R.total += count;
    };
    auto flush_items = [&]() -> cudaError_t {
      if (R.n == 0) return cudaSuccess;
      long long maxq = 1;
      for (int k = 0; k < R.n; ++k) maxq = std::max(maxq, (R.it[k].count + 3) / 4);
      long long blocks = (maxq + 255) / 256;
      if (blocks > 296) blocks = 296;
      k_reduce_list<<<dim3((unsigned)blocks, (unsigned)R.n), 256, 0, st>>>(g, R);
      R = ReduceList{};
      return cudaGetLastError();
    };
    for (int l = nl - 1; l >= 0; --l) {
      const StreamLayer& L = M.L[l];
      if (l == nl - 1 && htc) {
        // the fused head has produced the delta of layer nl-2 (BF16 halves) and its own partial sums
        const int hb = (nc + HEAD_COLS - 1) / HEAD_COLS;
        const long long hs = head_part_stride(L.nin);
        add_item(L.w_off, L.nin, W.head_part, hb, hs);
        add_item(M.L[l - 1].b_off, L.nin, W.head_part + L.nin, hb, hs);
        add_item(L.b_off, 1, W.head_part + 2 * L.nin, hb, hs);
        D = nullptr; Dhi = W.delta_hi[0]; Dlo = W.delta_lo[0];
        ping = 1;
        bias_slots = -1;  // bias gradient of layer nl-2 already listed
        continue;
      }
      if (L.tc) {
        // dW (nout x nin) = D * act[l]^T on the tensor cores, split over the batch into partials
        const long long cnt = (long long)L.nout * L.nin;
        int ksplit = tc_ksplit_for(L, nc);
        if (ksplit > W.tc_ksplit[l]) ksplit = W.tc_ksplit[l];
        int kch = (nc + ksplit - 1) / ksplit;
        kch = (kch + tcg::BK - 1) / tcg::BK * tcg::BK;
        ksplit = (nc + kch - 1) / kch;
        tcg::Args t{};
        t.M = L.nout; t.N = L.nin; t.K = nc;
        t.C = W.tc_part[l]; t.ldc = L.nout; t.accumulate = 0; t.kchunk = kch; t.part_stride = tc_part_stride(L);
        SCK(tc2_dw(t, tcg::Operand{Dhi, Dlo, L.nout}, tcg::Operand{W.act_hi[l], W.act_lo[l], L.nin}, ksplit, st));
        add_item(L.w_off, cnt, W.tc_part[l], ksplit, tc_part_stride(L));
        if (bias_slots > 0) add_item(L.b_off, L.nout, W.tc_rowpart[l], bias_slots, L.nout);
        if (bucketed) {
          // the gradient of this layer (and, for the last hidden layer, of the head with the likelihood sums) is
          // complete: reduce its partials now and start its all-reduce beside the remaining GEMMs
          if (l == nl - 2) {
            k_pack_sums<<<1, 1, 0, st>>>(g, M.P, sc);
            SCK(cudaGetLastError());
          }
          SCK(flush_items());
          const long long lo = L.w_off, hi = l == nl - 2 ? M.P + 3 : M.L[l + 1].w_off;
          SCK(cudaEventRecord(hook->ev_fork[nbuckets], st));
          SCK(cudaStreamWaitEvent(hook->side, hook->ev_fork[nbuckets], 0));
          SCK(hook->reduce(hook->ctx, g + lo, hi - lo, hook->side));
          ++nbuckets;
        }
        if (bias_slots == 0) {  // delta came from an FP32 kernel: deterministic two-level row sum of its FP32 copy
          dim3 grid((L.nout + 31) / 32, 8);
          k_rowsum_part<<<grid, 256, 0, st>>>(D, L.nout, nc, nullptr, W.rowpart);
          SCK(cudaGetLastError());
          k_rowsum_final<<<(L.nout + 127) / 128, 128, 0, st>>>(W.rowpart, L.nout, 64, g + L.b_off);
          SCK(cudaGetLastError());
        }
        bias_slots = 0;
        if (l > 0) {
          // delta_{l-1} (nin x nc) = (W_l^T D) .* act'(act[l])
          const bool prev_tc = M.L[l - 1].tc;
          tcg::Args u{};
          u.M = L.nin; u.N = nc; u.K = L.nout;
          u.C = prev_tc ? nullptr : W.delta[ping]; u.ldc = L.nin;
          u.aux_hi = W.act_hi[l]; u.aux_lo = W.act_lo[l]; u.ldaux = L.nin; u.act = M.L[l - 1].act;
          u.mask_in = W.mask[l];
          u.rowpart = prev_tc ? W.tc_rowpart[l - 1] : nullptr;
          u.kchunk = (L.nout + tcg::BK - 1) / tcg::BK * tcg::BK; u.part_stride = 0;
          SCK(tc2_dx(u, tcg::Operand{W.w_hi + L.w_off, W.w_lo + L.w_off, L.nout}, tcg::Operand{Dhi, Dlo, L.nout},
                     tcg::Output{prev_tc ? W.delta_hi[ping] : nullptr, prev_tc ? W.delta_lo[ping] : nullptr}, st));
          D = W.delta[ping]; Dhi = W.delta_hi[ping]; Dlo = W.delta_lo[ping];
          {
            const int bn = tcg::tile_cols_kmajor_b(u.M, u.N);
            bias_slots = prev_tc ? tcg::ROWSLOTS * ((nc + bn - 1) / bn) : 0;
          }
          ping ^= 1;
        }
        continue;
      }
      const bool prev_tc = l > 0 && M.L[l - 1].tc;
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
          k_out_dx<<<1184, 256, 0, st>>>(theta + L.w_off, D, W.act[l], L.nin, nc, M.L[l - 1].act, W.delta[ping],
                                        prev_tc ? W.delta_hi[ping] : nullptr, prev_tc ? W.delta_lo[ping] : nullptr);
          SCK(cudaGetLastError());
          D = W.delta[ping]; Dhi = W.delta_hi[ping]; Dlo = W.delta_lo[ping];
          bias_slots = 0;
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
        if (prev_tc) { a.chi = W.delta_hi[ping]; a.clo = W.delta_lo[ping]; }
        SCK((launch_gemm<false, true, EPI_MUL_DACT>(a, 1, st)));
        D = W.delta[ping]; Dhi = W.delta_hi[ping]; Dlo = W.delta_lo[ping];
        bias_slots = 0;
        ping ^= 1;
      }
    }
    SCK(flush_items());
    if (bucketed && nbuckets > 0) SCK(cudaEventRecord(hook->ev_join, hook->side));
  }
  return cudaSuccess;
}

cudaError_t stream_finish(const StreamModel& M, const float* theta, float* g, StreamScalars* sc, float nb, bool want_grad,
                          bool from_buffer, cudaStream_t st) {
  k_finish<<<STREAM_RED_BLOCKS, 256, 0, st>>>(M, theta, g, sc, nb, want_grad ? 1 : 0, from_buffer ? 1 : 0);
  return cudaGetLastError();
}

// =====================================================================================================
// sampler updates (flat order, Philox quad q = J / 4).  Scalars shared by all blocks (step counters, thermostat)
// are read at the start of a pass and advanced by the last block to finish it.
// =====================================================================================================
struct UpdDev {
  SamplerDev S;
  long long P;
  float *theta, *mom, *minv, *rmsv, *g;
  StreamScalars* sc;
  NoiseKey key;
  long long gstep;
  const float* inj;
  float* sample_out;
  float* samples_base;
  long long keep_every, kept_origin, ring_slots;
  float* kinetic_base;
  long long hist_origin, hist_stride;
  float* kinetic_out;
  __nv_bfloat16 *w_hi, *w_lo;
};

__device__ __forceinline__ unsigned long long upd_step(const UpdDev& U) {
  return U.gstep >= 0 ? (unsigned long long)U.gstep : (unsigned long long)U.sc->gstep;
}

// ---- quad-wise access: thread = four consecutive parameters (one Philox call, 16-byte loads / stores); the last
// quad of an odd-sized theta is handled element-wise ------------------------------------------------------------
__device__ __forceinline__ void ldq(const float* p, long long i, int cnt, float (&v)[4]) {
  if (cnt == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p + i);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = j < cnt ? p[i + j] : 0.f;
  }
}
__device__ __forceinline__ void stq(float* p, long long i, int cnt, const float (&v)[4]) {
  if (cnt == 4 && (reinterpret_cast<unsigned long long>(p + i) & 15ull) == 0) {
    *reinterpret_cast<float4*>(p + i) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < cnt) p[i + j] = v[j];
  }
}
__device__ __forceinline__ void noise_quad(const UpdDev& U, unsigned long long step, unsigned slot, long long i, int cnt,
                                           float (&z)[4]) {
  if (U.inj) {
#pragma unroll
    for (int j = 0; j < 4; ++j) z[j] = j < cnt ? U.inj[i + j] : 0.f;
    return;
  }
  const float4 n = normal_quad(U.key, step, slot, (unsigned)(i >> 2));
  z[0] = n.x; z[1] = n.y; z[2] = n.z; z[3] = n.w;
}

// where the sample recorded by this update goes (ns = nsampled after the update), or nullptr
__device__ __forceinline__ float* sample_dst(const UpdDev& U, long long ns) {
  if (U.sample_out) return U.sample_out;
  if (!U.samples_base) return nullptr;
  const long long ke = U.keep_every > 0 ? U.keep_every : 1;
  if (ns % ke != 0) return nullptr;
  long long slot = ns / ke - 1 - U.kept_origin;
  if (slot < 0) return nullptr;
  if (U.ring_slots > 0) slot %= U.ring_slots;
  return U.samples_base + slot * U.P;
}

__device__ __forceinline__ void record_kinetic(const UpdDev& U, long long t, float kin) {
  if (U.kinetic_out) *U.kinetic_out = kin;
  if (U.kinetic_base) {
    const long long k = t - 1 - U.hist_origin;
    if (k >= 0 && k < U.hist_stride) U.kinetic_base[k] = kin;
  }
}

// the updated theta: state, optional recorded sample, optional BF16 halves for the tensor-core layers
__device__ __forceinline__ void store_theta(const UpdDev& U, float* sdst, long long i, int cnt, const float (&t)[4]) {
  stq(U.theta, i, cnt, t);
  if (sdst) stq(sdst, i, cnt, t);
  if (U.w_hi) {
    if (cnt == 4) {
      const __nv_bfloat162 h0 = __floats2bfloat162_rn(t[0], t[1]), h1 = __floats2bfloat162_rn(t[2], t[3]);
      const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
      const __nv_bfloat162 l0 = __floats2bfloat162_rn(t[0] - f0.x, t[1] - f0.y), l1 = __floats2bfloat162_rn(t[2] - f1.x, t[3] - f1.y);
      uint2 hw, lw;
      hw.x = *reinterpret_cast<const unsigned*>(&h0); hw.y = *reinterpret_cast<const unsigned*>(&h1);
      lw.x = *reinterpret_cast<const unsigned*>(&l0); lw.y = *reinterpret_cast<const unsigned*>(&l1);
      *reinterpret_cast<uint2*>(U.w_hi + i) = hw;
      *reinterpret_cast<uint2*>(U.w_lo + i) = lw;
    } else {
      for (int j = 0; j < cnt; ++j) {
        const __nv_bfloat16 h = __float2bfloat16_rn(t[j]);
        U.w_hi[i + j] = h;
        U.w_lo[i + j] = __float2bfloat16_rn(t[j] - __bfloat162float(h));
      }
    }
  }
}

template <class F>
__device__ __forceinline__ void for_quads(long long P, F f) {
  const long long nq = (P + 3) >> 2;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < nq; q += (long long)gridDim.x * blockDim.x) {
    const long long i = q << 2;
    const long long rem = P - i;
    f(i, rem < 4 ? (int)rem : 4);
  }
}

// SGLD (sgld.jl:96-112)
__global__ void __launch_bounds__(256) k_upd_sgld(UpdDev U) {
  StreamScalars* sc = U.sc;
  const unsigned long long step = upd_step(U);
  if (sc->status) {  // stopped chain: only the step counter moves on (the host mirror counts every update!)
    if (grid_last(&sc->ticket) && threadIdx.x == 0) sc->gstep = (long long)step + 1;
    return;
  }
  const SamplerDev& S = U.S;
  const long long t0 = sc->t, ns = sc->nsampled + 1;
  float alpha = S.stepsize_a * powf(S.stepsize_b + (float)t0, -S.stepsize_gamma);
  alpha = fmaxf(alpha, S.min_stepsize);
  const float ng = sqrtf((float)sc->gn2);
  const float ha = 0.5f * alpha * (ng > S.maxnorm ? S.maxnorm / ng : 1.f), sa = sqrtf(alpha);
  float* sdst = sample_dst(U, ns);
  for_quads(U.P, [&](long long i, int cnt) {
    float th[4], g[4], z[4];
    ldq(U.theta, i, cnt, th);
    ldq(U.g, i, cnt, g);
    noise_quad(U, step, 0, i, cnt, z);
#pragma unroll
    for (int j = 0; j < 4; ++j) th[j] = th[j] + ha * g[j] + sa * z[j];
    store_theta(U, sdst, i, cnt, th);
  });
  if (grid_last(&sc->ticket) && threadIdx.x == 0) {
    sc->nsampled = ns;
    sc->t = t0 + 1;
    sc->gstep = (long long)step + 1;
  }
}

// SGNHT (sgnht.jl:64-90)
__global__ void __launch_bounds__(256) k_upd_sgnht(UpdDev U) {
  StreamScalars* sc = U.sc;
  const unsigned long long step = upd_step(U);
  if (sc->status) {
    if (grid_last(&sc->ticket) && threadIdx.x == 0) sc->gstep = (long long)step + 1;
    return;
  }
  const SamplerDev& S = U.S;
  const long long t0 = sc->t, ns = sc->nsampled + 1;
  const bool bad_g = sc->gn2 != sc->gn2;  // error("NaN in g") sgnht.jl:70-72: the state stays untouched
  const float l = S.l, xi = sc->xi, cn = sqrtf(l * 2.f * S.A);
  float a0 = 0.f, a1 = 0.f;
  float* sdst = sample_dst(U, ns);
  if (!bad_g)
    for_quads(U.P, [&](long long i, int cnt) {
      float th[4], g[4], p[4], z[4];
      ldq(U.theta, i, cnt, th);
      ldq(U.g, i, cnt, g);
      ldq(U.mom, i, cnt, p);
      noise_quad(U, step, 0, i, cnt, z);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        p[j] = p[j] - xi * l * p[j] + l * g[j] + cn * z[j];
        th[j] = th[j] + l * p[j];
        if (j < cnt) {
          a0 += p[j] * p[j];
          a1 += (th[j] != th[j]) ? 1.f : 0.f;
        }
      }
      stq(U.mom, i, cnt, p);
      store_theta(U, sdst, i, cnt, th);
    });
  double tot[2];
  if (grid_sum2((double)a0, (double)a1, sc, &sc->ticket, tot) && threadIdx.x == 0) {
    if (bad_g) {
      sc->status = 4;
    } else {
      const float kin = (float)tot[0] / (float)U.P;
      record_kinetic(U, t0, kin);
      sc->kin = kin;
      sc->xi = xi + (kin - 1.f) * S.l;
      if (tot[1] > 0.0) {
        sc->status = 5;  // error("NaN in θ") sgnht.jl:81-83
      } else {
        sc->nsampled = ns;
        sc->t = t0 + 1;
      }
    }
    sc->gstep = (long long)step + 1;
  }
}

// RMSPropMassAdapter (rmspropmassadapter.jl:26-46) on the gradient already in g
__global__ void __launch_bounds__(256) k_upd_rmsprop(UpdDev U) {
  StreamScalars* sc = U.sc;
  if (sc->status || sc->ma_t > U.S.ma_adapt_steps) return;
  const int ma_t = sc->ma_t;
  const bool first = ma_t == 1;
  const float inv_ng = 1.f / sqrtf((float)sc->gn2);
  for_quads(U.P, [&](long long i, int cnt) {
    float g[4], V[4], mi[4];
    ldq(U.g, i, cnt, g);
    ldq(U.rmsv, i, cnt, V);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gg = g[j] * inv_ng;
      const float v0 = first ? 1.f : V[j];
      V[j] = U.S.ma_alpha * v0 + (1.f - U.S.ma_alpha) * gg * gg;
      mi[j] = 1.f / (U.S.ma_lambda + sqrtf(V[j]));
    }
    stq(U.rmsv, i, cnt, V);
    stq(U.minv, i, cnt, mi);
  });
  if (grid_last(&sc->ticket) && threadIdx.x == 0) sc->ma_t = ma_t + 1;
}

// SGNHTS first half: B, A, C half steps (sgnht-s.jl:95-97)
__global__ void __launch_bounds__(256) k_upd_sgnhts_a(UpdDev U) {
  StreamScalars* sc = U.sc;
  if (sc->status) return;
  const SamplerDev& S = U.S;
  const bool bad_g = sc->gn2 != sc->gn2;  // error("NaN in g") sgnht-s.jl:90-92
  const float hl = 0.5f * S.l, xi0 = sc->xi;
  float a0 = 0.f;
  if (!bad_g)
    for_quads(U.P, [&](long long i, int cnt) {
      float th[4], g[4], p[4], mi[4] = {1.f, 1.f, 1.f, 1.f};
      ldq(U.theta, i, cnt, th);
      ldq(U.g, i, cnt, g);
      ldq(U.mom, i, cnt, p);
      if (U.minv) ldq(U.minv, i, cnt, mi);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        p[j] = p[j] + hl * g[j];
        th[j] += hl * mi[j] * p[j];
        if (j < cnt) a0 += p[j] * p[j] * mi[j];
      }
      stq(U.mom, i, cnt, p);
      stq(U.theta, i, cnt, th);
    });
  double tot[2];
  if (grid_sum2((double)a0, 0.0, sc, &sc->ticket, tot) && threadIdx.x == 0) {
    if (bad_g) {
      sc->status = 4;
    } else {
      const float xi = xi0 + S.l / (2.f * S.mu) * ((float)tot[0] - (float)U.P);
      sc->xi = xi;
      if (xi != 0.f) {
        sc->e1 = expf(-xi * S.l);
        sc->sc = S.sigmaA * sqrtf((1.f - expf(-2.f * xi * S.l)) / (2.f * xi));
      } else {
        sc->e1 = 1.f;
        sc->sc = sqrtf(S.l) * S.sigmaA;
      }
    }
  }
}

// SGNHTS second half: O, C, A (sgnht-s.jl:98-104), recording (:110-111)
__global__ void __launch_bounds__(256) k_upd_sgnhts_b(UpdDev U) {
  StreamScalars* sc = U.sc;
  const unsigned long long step = upd_step(U);
  if (sc->status) {
    if (grid_last(&sc->ticket) && threadIdx.x == 0) sc->gstep = (long long)step + 1;
    return;
  }
  const SamplerDev& S = U.S;
  const long long t0 = sc->t, ns = sc->nsampled + 1;
  const float hl = 0.5f * S.l, e1 = sc->e1, scl = sc->sc, xi0 = sc->xi;
  float a0 = 0.f, a1 = 0.f;
  float* sdst = sample_dst(U, ns);
  for_quads(U.P, [&](long long i, int cnt) {
    float th[4], p[4], z[4], mi[4] = {1.f, 1.f, 1.f, 1.f};
    ldq(U.theta, i, cnt, th);
    ldq(U.mom, i, cnt, p);
    if (U.minv) ldq(U.minv, i, cnt, mi);
    noise_quad(U, step, 0, i, cnt, z);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      p[j] = e1 * p[j] + scl * rsqrtf(mi[j]) * z[j];
      th[j] = th[j] + hl * mi[j] * p[j];
      if (j < cnt) {
        a0 += p[j] * p[j] * mi[j];
        a1 += (th[j] != th[j]) ? 1.f : 0.f;
      }
    }
    stq(U.mom, i, cnt, p);
    store_theta(U, sdst, i, cnt, th);
  });
  double tot[2];
  if (grid_sum2((double)a0, (double)a1, sc, &sc->ticket, tot) && threadIdx.x == 0) {
    sc->xi = xi0 + S.l / (2.f * S.mu) * ((float)tot[0] - (float)U.P);
    if (tot[1] > 0.0) {
      sc->status = 5;
    } else {
      const float kin = (float)tot[0] / (float)U.P;
      record_kinetic(U, t0, kin);
      sc->kin = kin;
      sc->nsampled = ns;
      sc->t = t0 + 1;
    }
    sc->gstep = (long long)step + 1;
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
  U.samples_base = su.samples_base;
  U.keep_every = su.keep_every; U.kept_origin = su.kept_origin; U.ring_slots = su.ring_slots;
  U.kinetic_base = su.kinetic_base; U.hist_origin = su.hist_origin; U.hist_stride = su.hist_stride;
  U.kinetic_out = su.kinetic_out;
  U.w_hi = su.w_hi; U.w_lo = su.w_lo;
  const int grid = STREAM_RED_BLOCKS;  // 592 x 256 threads = one quad per thread for P up to 606 208, grid-stride beyond
  switch (U.S.kind) {
    case S_SGLD:
      k_upd_sgld<<<grid, 256, 0, st>>>(U);
      break;
    case S_SGNHT:
      k_upd_sgnht<<<grid, 256, 0, st>>>(U);
      break;
    case S_SGNHTS:
      if (U.S.madapter_kind == MA_RMSPROP) k_upd_rmsprop<<<grid, 256, 0, st>>>(U);
      k_upd_sgnhts_a<<<grid, 256, 0, st>>>(U);
      k_upd_sgnhts_b<<<grid, 256, 0, st>>>(U);
      break;
    default:
      return cudaErrorNotSupported;
  }
  return cudaGetLastError();
}

}  // namespace bfx
