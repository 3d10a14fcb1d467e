// Row kernels of the soft quantizer VQ_Payam_GSSoft (Autoencoder_VQVAE_model.py:1304-1433), sm_100a.
//
// The layer is four dense products (mean_layer, logvar_layer, the distance contraction m E^T, p E) around a
// row-wise soft assignment; the products run on the tensor cores (g2v_gemm.cu), the rest is here:
//
//   soft_assign_kernel   d = |m|^2 + |e|^2 - 2 m.e ; s = 1 / exp(lv)^2 ; p~ = exp(-(d / 400)(0.5 s)) / sqrt(s) ;
//                        p = p~ / sum_k p~ (soft_prob, :1349-1372) + the column sums of p for the perplexity --
//                        one pass over the [N, K] products, d and p written once, no other [N, K] temporary
//                        (the reference materialises eight)
//   soft_tail_kernel     out = x + (q - x), sum (q - x)^2                                   (:1412-1424)
//   soft_scalars_kernel  loss = mse + beta mse, perplexity = exp(-sum avg log(avg + 1e-10))  (:1418-1426)
//   soft_bwd_kernel      closed-form backward of the soft assignment (oracle/gssoft_oracle.py):
//                        t = sum_k p dp ; g = p (dp - t) ; gd = -g s / 800 ; glv = g (d s / 400 + 1), with the
//                        row sums of gd and the column sums of gd and glv the dense products of the backward need
//   soft_gx_kernel       gx = c gmW + a (x - q) + g_out
// One warp per row; lanes stride the codes (coalesced 128-byte accesses).
#include "g2v_common.cuh"

#include <math.h>

#include <algorithm>

namespace g2v {
namespace {

constexpr int SW = 8;                        // warps per block
constexpr float kSoftScale = 400.f;          // the literal `dist / 400` of soft_prob (:1351)

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double wsumd(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// dot [N,K] holds m E^T on entry and the distances d on exit
__global__ void __launch_bounds__(SW * 32) soft_assign_kernel(const float* __restrict__ m, float* __restrict__ dot,
                                                              const float* __restrict__ lv, const float* __restrict__ e2,
                                                              long long N, int K, int D, float* __restrict__ p,
                                                              float* colsum) {
  extern __shared__ float cs[];                 // [K] column sums of this block
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = threadIdx.x; k < K; k += blockDim.x) cs[k] = 0.f;
  __syncthreads();
  for (long long row = (long long)blockIdx.x * SW + warp; row < N; row += (long long)gridDim.x * SW) {
    const float* mr = m + (size_t)row * D;
    float m2 = 0.f;
    for (int j = lane; j < D; j += 32) m2 = fmaf(mr[j], mr[j], m2);
    m2 = wsum(m2);
    float* dr = dot + (size_t)row * K;
    const float* lr = lv + (size_t)row * K;
    float* pr = p + (size_t)row * K;
    float sum = 0.f;
    for (int k = lane; k < K; k += 32) {
      const float d = (m2 + e2[k]) - 2.f * dr[k];               // the reference's order: (|m|^2 + |e|^2) - 2 m.e
      const float ex = expf(lr[k]);
      const float s = 1.f / (ex * ex);                          // smooth = 1 / exp(logvar)^2      (:1399)
      const float prob = expf(-((d / kSoftScale) * (0.5f * s))) / sqrtf(s);
      dr[k] = d;
      pr[k] = prob;
      sum += prob;
    }
    sum = wsum(sum);
    for (int k = lane; k < K; k += 32) {
      const float v = pr[k] / sum;
      pr[k] = v;
      atomicAdd(&cs[k], v);
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x)
    if (cs[k] != 0.f) atomicAdd(colsum + k, cs[k]);
}

// K <= 32 KI: a lane keeps its K / 32 codes of the row AND its share of the column sums in registers -- p is written
// once (already normalised) and the per-element shared-memory atomics of the generic kernel (one per (row, code):
// they were 4/5 of its time) become one per (warp, code).
template <int KI>
__global__ void __launch_bounds__(SW * 32) soft_assign_reg_kernel(const float* __restrict__ m, float* __restrict__ dot,
                                                                  const float* __restrict__ lv, const float* __restrict__ e2,
                                                                  long long N, int K, int D, float* __restrict__ p,
                                                                  float* colsum) {
  extern __shared__ float cs[];                 // [K] column sums of this block
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = threadIdx.x; k < K; k += blockDim.x) cs[k] = 0.f;
  __syncthreads();
  float csum[KI], e2r[KI];
#pragma unroll
  for (int i = 0; i < KI; ++i) {
    csum[i] = 0.f;
    e2r[i] = (lane + 32 * i < K) ? e2[lane + 32 * i] : 0.f;
  }
  for (long long row = (long long)blockIdx.x * SW + warp; row < N; row += (long long)gridDim.x * SW) {
    const float* mr = m + (size_t)row * D;
    float* dr = dot + (size_t)row * K;
    const float* lr = lv + (size_t)row * K;
    float* pr = p + (size_t)row * K;
    float dv[KI], lvv[KI];
#pragma unroll
    for (int i = 0; i < KI; ++i) {
      const int k = lane + 32 * i;
      dv[i] = k < K ? __ldcs(dr + k) : 0.f;
      lvv[i] = k < K ? __ldcs(lr + k) : 0.f;
    }
    float m2 = 0.f;
    for (int j = lane; j < D; j += 32) m2 = fmaf(mr[j], mr[j], m2);
    m2 = wsum(m2);
    float prob[KI];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < KI; ++i) {
      const int k = lane + 32 * i;
      prob[i] = 0.f;
      if (k < K) {
        const float d = (m2 + e2r[i]) - 2.f * dv[i];             // the reference's order: (|m|^2 + |e|^2) - 2 m.e
        const float ex = expf(lvv[i]);
        const float s = 1.f / (ex * ex);                         // smooth = 1 / exp(logvar)^2      (:1399)
        prob[i] = expf(-((d / kSoftScale) * (0.5f * s))) / sqrtf(s);
        dr[k] = d;
        sum += prob[i];                                          // (same lane order as the generic kernel)
      }
    }
    sum = wsum(sum);
#pragma unroll
    for (int i = 0; i < KI; ++i) {
      const int k = lane + 32 * i;
      if (k < K) {
        const float v = prob[i] / sum;
        pr[k] = v;
        csum[i] += v;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < KI; ++i)
    if (lane + 32 * i < K && csum[i] != 0.f) atomicAdd(&cs[lane + 32 * i], csum[i]);
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x)
    if (cs[k] != 0.f) atomicAdd(colsum + k, cs[k]);
}

__global__ void __launch_bounds__(256) soft_tail_kernel(const float* __restrict__ x, const float* __restrict__ q,
                                                        long long n, float* __restrict__ out, double* sse) {
  __shared__ double sh[8];
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float xv = x[i], d = q[i] - xv;
    out[i] = xv + d;                                             // inputs + (quantized - inputs).detach()
    acc += (double)d * (double)d;
  }
  acc = wsumd(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    atomicAdd(sse, t);
  }
}

__global__ void __launch_bounds__(256) soft_scalars_kernel(const double* __restrict__ sse, const float* __restrict__ colsum,
                                                           long long N, int K, int D, float beta, float* loss, float* ppl) {
  __shared__ double sh[8];
  double h = 0.0;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float avg = colsum[k] / (float)N;                      // torch.mean(encodings, dim=0)
    h += (double)(avg * logf(avg + 1e-10f));
  }
  h = wsumd(h);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = h;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    *ppl = expf(-(float)t);
    const float mse = (float)(*sse / ((double)N * (double)D));
    *loss = __fadd_rn(mse, __fmul_rn(beta, mse));                // q_latent_loss + commitment_cost * e_latent_loss
  }
}

__global__ void __launch_bounds__(SW * 32) soft_bwd_kernel(const float* __restrict__ p, const float* __restrict__ dp,
                                                           const float* __restrict__ d, const float* __restrict__ lv,
                                                           long long N, int K, float* __restrict__ gd,
                                                           float* __restrict__ glv, float* __restrict__ rowsum_gd,
                                                           float* col_gd, float* col_glv) {
  extern __shared__ float cs[];                 // [2K]: column sums of gd, glv of this block
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = threadIdx.x; k < 2 * K; k += blockDim.x) cs[k] = 0.f;
  __syncthreads();
  for (long long row = (long long)blockIdx.x * SW + warp; row < N; row += (long long)gridDim.x * SW) {
    const size_t o = (size_t)row * K;
    float t = 0.f;
    for (int k = lane; k < K; k += 32) t = fmaf(p[o + k], dp[o + k], t);
    t = wsum(t);
    float rs = 0.f;
    for (int k = lane; k < K; k += 32) {
      const float g = p[o + k] * (dp[o + k] - t);               // d/d(log p~) of sum_k p_k dp_k
      const float ex = expf(lv[o + k]);
      const float s = 1.f / (ex * ex);
      const float a = -g * s / (2.f * kSoftScale);              // d log p~ / d d   = -s / 800
      const float b = g * (d[o + k] * s / kSoftScale + 1.f);    // d log p~ / d lv  = d s / 400 + 1
      gd[o + k] = a;
      glv[o + k] = b;
      rs += a;
      atomicAdd(&cs[k], a);
      atomicAdd(&cs[K + k], b);
    }
    rs = wsum(rs);
    if (lane == 0) rowsum_gd[row] = rs;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    if (cs[k] != 0.f) atomicAdd(col_gd + k, cs[k]);
    if (cs[K + k] != 0.f) atomicAdd(col_glv + k, cs[K + k]);
  }
}

template <int KI>
__global__ void __launch_bounds__(SW * 32) soft_bwd_reg_kernel(const float* __restrict__ p, const float* __restrict__ dp,
                                                               const float* __restrict__ d, const float* __restrict__ lv,
                                                               long long N, int K, float* __restrict__ gd,
                                                               float* __restrict__ glv, float* __restrict__ rowsum_gd,
                                                               float* col_gd, float* col_glv) {
  extern __shared__ float cs[];                 // [2K]: column sums of gd, glv of this block
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = threadIdx.x; k < 2 * K; k += blockDim.x) cs[k] = 0.f;
  __syncthreads();
  float ca[KI], cb[KI];
#pragma unroll
  for (int i = 0; i < KI; ++i) ca[i] = cb[i] = 0.f;
  for (long long row = (long long)blockIdx.x * SW + warp; row < N; row += (long long)gridDim.x * SW) {
    const size_t o = (size_t)row * K;
    float pv[KI], dpv[KI], dv[KI], lvv[KI];
#pragma unroll
    for (int i = 0; i < KI; ++i) {
      const int k = lane + 32 * i;
      const bool in = k < K;
      pv[i] = in ? __ldcs(p + o + k) : 0.f;
      dpv[i] = in ? __ldcs(dp + o + k) : 0.f;
      dv[i] = in ? __ldcs(d + o + k) : 0.f;
      lvv[i] = in ? __ldcs(lv + o + k) : 0.f;
    }
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < KI; ++i) t = fmaf(pv[i], dpv[i], t);
    t = wsum(t);
    float rs = 0.f;
#pragma unroll
    for (int i = 0; i < KI; ++i) {
      const int k = lane + 32 * i;
      if (k < K) {
        const float g = pv[i] * (dpv[i] - t);                    // d/d(log p~) of sum_k p_k dp_k
        const float ex = expf(lvv[i]);
        const float s = 1.f / (ex * ex);
        const float a = -g * s / (2.f * kSoftScale);             // d log p~ / d d   = -s / 800
        const float b = g * (dv[i] * s / kSoftScale + 1.f);      // d log p~ / d lv  = d s / 400 + 1
        __stcs(gd + o + k, a);
        __stcs(glv + o + k, b);
        rs += a;
        ca[i] += a;
        cb[i] += b;
      }
    }
    rs = wsum(rs);
    if (lane == 0) rowsum_gd[row] = rs;
  }
#pragma unroll
  for (int i = 0; i < KI; ++i) {
    const int k = lane + 32 * i;
    if (k < K) {
      if (ca[i] != 0.f) atomicAdd(&cs[k], ca[i]);
      if (cb[i] != 0.f) atomicAdd(&cs[K + k], cb[i]);
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    if (cs[k] != 0.f) atomicAdd(col_gd + k, cs[k]);
    if (cs[K + k] != 0.f) atomicAdd(col_glv + k, cs[K + k]);
  }
}

// gx = c[0] * gmw + a[0] * (x - q) + g_out (g_out optional); c, a are device scalars
__global__ void __launch_bounds__(256) soft_gx_kernel(const float* __restrict__ gmw, const float* __restrict__ x,
                                                      const float* __restrict__ q, const float* __restrict__ g_out,
                                                      const float* __restrict__ c, const float* __restrict__ a, long long n,
                                                      float* __restrict__ gx) {
  const float cc = *c, aa = *a;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    gx[i] = fmaf(cc, gmw[i], fmaf(aa, x[i] - q[i], g_out ? g_out[i] : 0.f));
}

inline int grid_rows(long long rows, int per_block) {
  long long g = (rows + per_block - 1) / per_block;
  const long long cap = (long long)num_sms() * 8;
  return (int)(g < 1 ? 1 : (g < cap ? g : cap));
}

}  // namespace

int launch_soft_assign(const float* m, float* dot, const float* lv, const float* e2, int64_t N, int K, int D, float* p,
                       float* colsum, cudaStream_t st) {
  // a persistent grid: the per-warp column sums are flushed once per warp
  const int g = (int)std::max<long long>(1, std::min<long long>((N + SW - 1) / SW, (long long)num_sms() * 4));
  if (K <= 256) soft_assign_reg_kernel<8><<<g, SW * 32, (size_t)K * sizeof(float), st>>>(m, dot, lv, e2, N, K, D, p, colsum);
  else if (K <= 512) soft_assign_reg_kernel<16><<<g, SW * 32, (size_t)K * sizeof(float), st>>>(m, dot, lv, e2, N, K, D, p, colsum);
  else if (K <= 1024) soft_assign_reg_kernel<32><<<g, SW * 32, (size_t)K * sizeof(float), st>>>(m, dot, lv, e2, N, K, D, p, colsum);
  else soft_assign_kernel<<<grid_rows(N, SW), SW * 32, (size_t)K * sizeof(float), st>>>(m, dot, lv, e2, N, K, D, p, colsum);
  G2V_LAUNCH_CHECK("soft_assign_kernel");
  return G2V_OK;
}

int launch_soft_tail(const float* x, const float* q, int64_t N, int K, int D, float beta, float* out, double* sse,
                     const float* colsum, float* loss, float* ppl, cudaStream_t st) {
  const long long n = (long long)N * D;
  soft_tail_kernel<<<grid_rows(n, 1024), 256, 0, st>>>(x, q, n, out, sse);
  G2V_LAUNCH_CHECK("soft_tail_kernel");
  soft_scalars_kernel<<<1, 256, 0, st>>>(sse, colsum, N, K, D, beta, loss, ppl);
  G2V_LAUNCH_CHECK("soft_scalars_kernel");
  return G2V_OK;
}

int launch_soft_bwd(const float* p, const float* dp, const float* d, const float* lv, int64_t N, int K, float* gd, float* glv,
                    float* rowsum_gd, float* col_gd, float* col_glv, cudaStream_t st) {
  const int g = (int)std::max<long long>(1, std::min<long long>((N + SW - 1) / SW, (long long)num_sms() * 4));
  const size_t sm = (size_t)2 * K * sizeof(float);
  if (K <= 256) soft_bwd_reg_kernel<8><<<g, SW * 32, sm, st>>>(p, dp, d, lv, N, K, gd, glv, rowsum_gd, col_gd, col_glv);
  else if (K <= 512) soft_bwd_reg_kernel<16><<<g, SW * 32, sm, st>>>(p, dp, d, lv, N, K, gd, glv, rowsum_gd, col_gd, col_glv);
  else soft_bwd_kernel<<<grid_rows(N, SW), SW * 32, sm, st>>>(p, dp, d, lv, N, K, gd, glv, rowsum_gd, col_gd, col_glv);
  G2V_LAUNCH_CHECK("soft_bwd_kernel");
  return G2V_OK;
}

int launch_soft_gx(const float* gmw, const float* x, const float* q, const float* g_out, const float* c, const float* a,
                   int64_t n, float* gx, cudaStream_t st) {
  soft_gx_kernel<<<grid_rows(n, 1024), 256, 0, st>>>(gmw, x, q, g_out, c, a, n, gx);
  G2V_LAUNCH_CHECK("soft_gx_kernel");
  return G2V_OK;
}

}  // namespace g2v
