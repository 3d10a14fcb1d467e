// CUDA-core (fp32 / fp64) kernels of the vector-quantizer path, sm_100a.
//
//   search_simt_kernel   register-tiled fp32 distance sweep with a fused running top-2 and an
//                        in-kernel exact (fp64) re-rank of the rows whose top-2 gap is inside the
//                        fp32 error bound.  General fallback (any K, D), and the second stage of
//                        the tensor-core path for the rows it cannot certify.
//   apply_kernel         gather + straight-through value + SSE + histogram + EMA residual sums.
//   backward_kernel      g_x = g_out + c*(x - E[idx]).
//   ema_* / stats_*      codebook update and scalar outputs.
//   onehot_kernel        dense one-hot the reference returns.
//
// Everything here is bandwidth- or CUDA-core-bound by nature; the tensor-core search lives in
// g2v_tc.cu.
#include "g2v_common.cuh"

#include <cooperative_groups.h>

#include <stdlib.h>

#include <algorithm>

#include <float.h>
#include <math.h>

namespace g2v {

namespace {

constexpr float kU32 = 5.9604645e-8f;  // 2^-24, fp32 unit roundoff

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// row streams (read once): evict-first, so they do not push the codebook rows the same kernel gathers out of L1 / L2
__device__ __forceinline__ float4 lds4(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------------------------------
// fp32 search with fused top-2 and exact re-rank
// ------------------------------------------------------------------------------------------
constexpr int BM = 128, BN = 128, BK = 16, NT = 256, LDS = BM + 4;

struct SearchSmem {
  float As[2][BK][LDS];
  float Bs[2][BK][LDS];
  float e2s[BN];
  float z2p[2][BM];
  int rows[BM];
};

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }

// 8 consecutive elements of a row (fp32 / fp16 / bf16 storage) as fp32; zero beyond D or for row < 0
template <bool VEC, typename T>
__device__ __forceinline__ void load8(const T* __restrict__ base, long long row, int D, int col,
                                      float (&v)[8]) {
  if (row < 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
    return;
  }
  const T* p = base + (size_t)row * D + col;
  if (VEC && col + 8 <= D) {
    if constexpr (sizeof(T) == 4) {
      float4 a = ldg4(reinterpret_cast<const float*>(p)), b = ldg4(reinterpret_cast<const float*>(p) + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
      v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
      uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
      const T* h = reinterpret_cast<const T*>(&raw);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = to_f32(h[i]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (col + i < D) ? to_f32(p[i]) : 0.f;
  }
}

template <bool VEC, typename ZT>
__global__ void __launch_bounds__(NT) search_simt_kernel(
    const ZT* __restrict__ z, const float* __restrict__ E, const float* __restrict__ e2,
    const float* __restrict__ ntab, long long N, int K, int D, int* __restrict__ full_list,
    int* __restrict__ full_count, int* __restrict__ idx_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SearchSmem& S = *reinterpret_cast<SearchSmem*>(smem_raw);
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int lrow = t & (BM - 1), lhalf = t >> 7;  // loader mapping: row, which 8 of the 16 columns
  const long long n_rows = N;
  const int nk = (D + BK - 1) / BK;
  const int n_ctile = (K + BN - 1) / BN;

  for (long long tile = blockIdx.x; tile * BM < n_rows; tile += gridDim.x) {
    __syncthreads();  // previous tile fully done with smem
    if (t < BM) {
      long long r = tile * BM + t;
      S.rows[t] = (r < n_rows) ? (int)r : -1;
    }
    __syncthreads();
    const long long grow = S.rows[lrow];

    float m1[8], m2[8];
    int i1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { m1[i] = INFINITY; m2[i] = INFINITY; i1[i] = 0; }
    float zsq = 0.f;

    for (int ct = 0; ct < n_ctile; ++ct) {
      const int kb = ct * BN;
      const long long crow = (kb + lrow < K) ? (kb + lrow) : -1;
      float acc[8][8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

      float ra[8], rb[8];
      load8<VEC>(z, grow, D, lhalf * 8, ra);
      load8<VEC>(E, crow, D, lhalf * 8, rb);
      if (t < BN) S.e2s[t] = (kb + t < K) ? e2[kb + t] : INFINITY;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        S.As[0][lhalf * 8 + i][lrow] = ra[i];
        S.Bs[0][lhalf * 8 + i][lrow] = rb[i];
      }
      if (ct == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) zsq = fmaf(ra[i], ra[i], zsq);
      }
      __syncthreads();

      for (int kc = 0; kc < nk; ++kc) {
        const int buf = kc & 1;
        if (kc + 1 < nk) {
          load8<VEC>(z, grow, D, (kc + 1) * BK + lhalf * 8, ra);
          load8<VEC>(E, crow, D, (kc + 1) * BK + lhalf * 8, rb);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
          float4 a0 = *reinterpret_cast<const float4*>(&S.As[buf][k][ty * 4]);
          float4 a1 = *reinterpret_cast<const float4*>(&S.As[buf][k][64 + ty * 4]);
          float4 b0 = *reinterpret_cast<const float4*>(&S.Bs[buf][k][tx * 4]);
          float4 b1 = *reinterpret_cast<const float4*>(&S.Bs[buf][k][64 + tx * 4]);
          float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
          float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kc + 1 < nk) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            S.As[buf ^ 1][lhalf * 8 + i][lrow] = ra[i];
            S.Bs[buf ^ 1][lhalf * 8 + i][lrow] = rb[i];
          }
          if (ct == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) zsq = fmaf(ra[i], ra[i], zsq);
          }
        }
        __syncthreads();
      }

      // fused running top-2 (ascending code order inside a thread, strict '<' keeps the first)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = (j < 4) ? (tx * 4 + j) : (64 + tx * 4 + j - 4);
        const float ek = S.e2s[c];
        const int k = kb + c;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float d = fmaf(-2.f, acc[i][j], ek);
          bool lt = d < m1[i];
          m2[i] = fminf(m2[i], fmaxf(d, m1[i]));
          i1[i] = lt ? k : i1[i];
          m1[i] = fminf(m1[i], d);
        }
      }
      __syncthreads();  // e2s / smem buffers are rewritten by the next code tile
    }

    // row norms for the error bound
    S.z2p[lhalf][lrow] = zsq;
    // merge the 16 column-threads of each row (they are one half-warp)
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float om1 = __shfl_xor_sync(0xffffffffu, m1[i], o);
        float om2 = __shfl_xor_sync(0xffffffffu, m2[i], o);
        int oi1 = __shfl_xor_sync(0xffffffffu, i1[i], o);
        bool take = (om1 < m1[i]) || (om1 == m1[i] && oi1 < i1[i]);
        float nm2 = take ? fminf(m1[i], om2) : fminf(m2[i], om1);
        m1[i] = take ? om1 : m1[i];
        i1[i] = take ? oi1 : i1[i];
        m2[i] = nm2;
      }
    }
    __syncthreads();
    if (tx == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = (i < 4) ? (ty * 4 + i) : (64 + ty * 4 + i - 4);
        const int g = S.rows[r];
        if (g < 0) continue;
        idx_out[g] = i1[i];
        // |d_hat - d| <= 2*D*u*|z||e| + u*e2 + u*|d_hat| per code; the gap of two codes can be
        // off by twice that.
        // (|e| bounded by the largest code norm that can still win this row, see reachable_norm)
        const float z2 = S.z2p[0][r] + S.z2p[1][r], zn = sqrtf(z2) * 1.0001f;
        const float cu = reachable_norm(ntab, zn, z2 + m1[i]);
        const float tau = 4.f * (float)(D + 2) * kU32 * zn * cu + 4.f * kU32 * (cu * cu + fabsf(m1[i]));
        // uncertified rows get their whole distance row recomputed in fp64 (full_recheck_kernel)
        if (!(m2[i] - m1[i] > tau)) full_list[atomicAdd(full_count, 1)] = g;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// exact re-rank of the listed rows over the whole codebook, first index on exact ties.
// One CTA takes FR_ROWS rows at a time (held in shared memory) so each codebook row fetched from
// L2 serves all of them.  Thread = code: a thread walks its code's D values once (128-bit loads)
// and the latent values arrive as shared-memory broadcasts, so there is no cross-lane reduction
// in the inner loop.  Phase A evaluates every distance in fp32 and certifies the rows whose top-2
// gap exceeds the fp32 error bound; phase B re-scans an uncertified row with the same arithmetic
// and evaluates in fp64 only the codes within the bound of its minimum.
// ------------------------------------------------------------------------------------------
constexpr int FR_THREADS = 256;

// fp32 dot products of one code row with R latent rows.  NP independent partial sums per row (column
// j goes to partial j % NP, NP in {1, 4}) shorten the rounding chain to D/NP + NP steps; the order is
// fixed, so re-evaluating a row gives bit-identical values.
template <int R, int NP, bool VEC>
__device__ __forceinline__ void fr_dots(const float* __restrict__ er, const float* __restrict__ zs, int Dp4, int D,
                                        float (&out)[R]) {
  float acc[R][NP];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int p = 0; p < NP; ++p) acc[r][p] = 0.f;
  if (VEC) {
#pragma unroll 2
    for (int j = 0; j < D; j += 4) {
      const float4 e = ldg4(er + j);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float4 zv = *reinterpret_cast<const float4*>(zs + r * Dp4 + j);
        acc[r][0] = fmaf(zv.x, e.x, acc[r][0]);
        acc[r][1 % NP] = fmaf(zv.y, e.y, acc[r][1 % NP]);
        acc[r][2 % NP] = fmaf(zv.z, e.z, acc[r][2 % NP]);
        acc[r][3 % NP] = fmaf(zv.w, e.w, acc[r][3 % NP]);
      }
    }
  } else {
    for (int j = 0; j < D; ++j) {
      const float e = __ldg(er + j);
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r][0] = fmaf(zs[r * Dp4 + j], e, acc[r][0]);
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) out[r] = (NP == 4) ? (acc[r][0] + acc[r][1 % NP]) + (acc[r][2 % NP] + acc[r][3 % NP]) : acc[r][0];
}

template <typename ZT, bool VEC, int FR_ROWS, int NP>
__global__ void __launch_bounds__(FR_THREADS) full_recheck_kernel(
    const ZT* __restrict__ z, const float* __restrict__ E, const float* __restrict__ e2,
    const float* __restrict__ ntab, int K, int D, const int* __restrict__ list,
    const int* __restrict__ count, int skip, int* __restrict__ idx_out, unsigned long long* stats, int Dz) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* zs = reinterpret_cast<float*>(smem_raw);                 // [FR_ROWS][Dp4]
  const int Dp4 = (D + 3) & ~3;
  constexpr int NW = FR_THREADS / 32;
  list += skip;                                                   // the first `skip` rows are handled by full64_kernel
  __shared__ double bv[NW];
  __shared__ float b1[NW][FR_ROWS], b2[NW][FR_ROWS];
  __shared__ int bi[NW][FR_ROWS];
  __shared__ int rows[FR_ROWS];
  __shared__ int need64[FR_ROWS];
  __shared__ float z2s[FR_ROWS], v1s[FR_ROWS], taus[FR_ROWS];
  __shared__ int n64;
  const int n = *count - skip;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // rounding steps on a dot product: the FMA chain of one partial, the partial merge, e2 and the final fma
  const float nround = (VEC && NP == 4) ? (float)(D / 4 + 6) : (float)(D + 2);
  if (threadIdx.x == 0) n64 = 0;
  for (int b0 = blockIdx.x * FR_ROWS; b0 < n; b0 += gridDim.x * FR_ROWS) {
    __syncthreads();
    if (threadIdx.x < FR_ROWS) rows[threadIdx.x] = (b0 + threadIdx.x < n) ? list[b0 + threadIdx.x] : -1;
    __syncthreads();
    for (int i = threadIdx.x; i < FR_ROWS * Dp4; i += blockDim.x) {
      const int r = i / Dp4, j = i - r * Dp4;
      zs[i] = (rows[r] >= 0 && j < Dz) ? to_f32(z[(size_t)rows[r] * Dz + j]) : 0.f;
    }
    __syncthreads();
    for (int r = warp; r < FR_ROWS; r += NW) {              // row norms for the fp32 error bound
      float s = 0.f;
      for (int j = lane; j < Dp4; j += 32) s = fmaf(zs[r * Dp4 + j], zs[r * Dp4 + j], s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) z2s[r] = s;
    }
    // ---- phase A: fp32 distances, running top-2 per row (codes ascend inside a thread) ----
    float m1[FR_ROWS], m2[FR_ROWS];
    int i1[FR_ROWS];
#pragma unroll
    for (int r = 0; r < FR_ROWS; ++r) { m1[r] = INFINITY; m2[r] = INFINITY; i1[r] = 0x7fffffff; }
    for (int k = threadIdx.x; k < K; k += FR_THREADS) {
      float acc[FR_ROWS];
      fr_dots<FR_ROWS, NP, VEC>(E + (size_t)k * D, zs, Dp4, D, acc);
      const float ek = __ldg(e2 + k);
#pragma unroll
      for (int r = 0; r < FR_ROWS; ++r) {
        const float d = fmaf(-2.f, acc[r], ek);
        const bool lt = d < m1[r];
        m2[r] = fminf(m2[r], fmaxf(d, m1[r]));
        i1[r] = lt ? k : i1[r];
        m1[r] = fminf(m1[r], d);
      }
    }
#pragma unroll
    for (int r = 0; r < FR_ROWS; ++r) {                     // warp merge (value, then lower index)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float o1 = __shfl_xor_sync(0xffffffffu, m1[r], o), o2 = __shfl_xor_sync(0xffffffffu, m2[r], o);
        const int oi = __shfl_xor_sync(0xffffffffu, i1[r], o);
        const bool take = (o1 < m1[r]) || (o1 == m1[r] && oi < i1[r]);
        const float nv2 = take ? fminf(m1[r], o2) : fminf(m2[r], o1);
        m1[r] = take ? o1 : m1[r];
        i1[r] = take ? oi : i1[r];
        m2[r] = nv2;
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < FR_ROWS; ++r) { b1[warp][r] = m1[r]; b2[warp][r] = m2[r]; bi[warp][r] = i1[r]; }
    }
    __syncthreads();
    if (threadIdx.x < FR_ROWS) {
      const int r = threadIdx.x;
      float v1 = b1[0][r], v2 = b2[0][r];
      int id = bi[0][r];
      for (int w = 1; w < NW; ++w) {
        const float o1 = b1[w][r], o2 = b2[w][r];
        const int oi = bi[w][r];
        const bool take = (o1 < v1) || (o1 == v1 && oi < id);
        const float nv2 = take ? fminf(v1, o2) : fminf(v2, o1);
        v1 = take ? o1 : v1;
        id = take ? oi : id;
        v2 = nv2;
      }
      // per-code error: `nround` roundings of at most u*|z||e| each; a gap can be off by twice that
      const float zn = sqrtf(z2s[r]) * 1.0001f;
      const float cu = reachable_norm(ntab, zn, z2s[r] + v1);
      const float tau = 4.f * nround * kU32 * zn * cu + 4.f * kU32 * (cu * cu + fabsf(v1));
      int need = 0;
      if (rows[r] >= 0) {
        idx_out[rows[r]] = id;
        need = !(v2 - v1 > tau);
      }
      need64[r] = need;
      v1s[r] = v1;
      taus[r] = tau;
      if (need) atomicAdd(&n64, 1);
    }
    __syncthreads();
    // ---- phase B (rare): same fp32 scan of the one row, fp64 for the codes within tau of its minimum ----
    for (int r = 0; r < FR_ROWS; ++r) {
      if (!need64[r]) continue;                            // block-uniform
      const float lim = v1s[r] + taus[r];
      double best = INFINITY;
      int besti = 0x7fffffff;
      for (int k = threadIdx.x; k < K; k += FR_THREADS) {
        float acc[1];
        fr_dots<1, NP, VEC>(E + (size_t)k * D, zs + r * Dp4, Dp4, D, acc);
        const float d = fmaf(-2.f, acc[0], __ldg(e2 + k));
        if (d <= lim) {
          const float* er = E + (size_t)k * D;
          double s = 0.0;
          for (int j = 0; j < D; ++j) {
            const double df = (double)zs[r * Dp4 + j] - (double)__ldg(er + j);
            s = fma(df, df, s);
          }
          if (s < best) { best = s; besti = k; }            // ascending k inside a thread: first wins
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (ov < best || (ov == best && oi < besti)) { best = ov; besti = oi; }
      }
      if (lane == 0) { bv[warp] = best; bi[warp][0] = besti; }
      __syncthreads();
      if (threadIdx.x == 0) {
        double v = bv[0];
        int id = bi[0][0];
        for (int w = 1; w < NW; ++w)
          if (bv[w] < v || (bv[w] == v && bi[w][0] < id)) { v = bv[w]; id = bi[w][0]; }
        if (id != 0x7fffffff) idx_out[rows[r]] = id;
      }
      __syncthreads();
    }
  }
  __syncthreads();
  if (stats && threadIdx.x == 0 && n64) atomicAdd(stats + G2V_STAT_FULL_RECHECK, (unsigned long long)n64);
}

// ------------------------------------------------------------------------------------------
// gather + STE value + statistics
// ------------------------------------------------------------------------------------------
constexpr int APPLY_WARPS = 8;

template <bool VEC>
__global__ void __launch_bounds__(APPLY_WARPS * 32) apply_kernel(
    const float* __restrict__ x, const float* __restrict__ zs, const float* __restrict__ E,
    const int* __restrict__ idx, long long N, int K, int D, float* __restrict__ out, double* sse,
    int* counts, float* dwr, int dwr_replicas, int use_hist) {
  extern __shared__ int hist[];
  __shared__ double wsum[APPLY_WARPS];
  if (dwr) dwr += (size_t)(blockIdx.x % dwr_replicas) * K * D;      // this block's private copy
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (use_hist) {
    for (int k = threadIdx.x; k < K; k += blockDim.x) hist[k] = 0;
    __syncthreads();
  }
  float acc = 0.f;
  const long long stride = (long long)gridDim.x * APPLY_WARPS;
  for (long long row = (long long)blockIdx.x * APPLY_WARPS + warp; row < N; row += stride) {
    int k = __ldg(idx + row);
    k = min(max(k, 0), K - 1);
    const float* xr = x + (size_t)row * D;
    const float* er = E + (size_t)k * D;
    const float* zr = zs ? zs + (size_t)row * D : nullptr;
    float* orow = out ? out + (size_t)row * D : nullptr;
    float* drow = dwr ? dwr + (size_t)k * D : nullptr;
    if (VEC) {
      for (int c = lane * 4; c < D; c += 128) {
        float4 xv = lds4(xr + c), ev = ldg4(er + c);
        float4 d = make_float4(ev.x - xv.x, ev.y - xv.y, ev.z - xv.z, ev.w - xv.w);
        acc = fmaf(d.x, d.x, acc); acc = fmaf(d.y, d.y, acc);
        acc = fmaf(d.z, d.z, acc); acc = fmaf(d.w, d.w, acc);
        if (orow) {
          float4 o = make_float4(xv.x + d.x, xv.y + d.y, xv.z + d.z, xv.w + d.w);
          __stcs(reinterpret_cast<float4*>(orow + c), o);
        }
        if (drow) {
          if (zr) {
            float4 zv = lds4(zr + c);
            red_add_v4(drow + c, zv.x - ev.x, zv.y - ev.y, zv.z - ev.z, zv.w - ev.w);
          } else {
            red_add_v4(drow + c, -d.x, -d.y, -d.z, -d.w);
          }
        }
      }
    } else {
      for (int c = lane; c < D; c += 32) {
        float xv = __ldg(xr + c), ev = __ldg(er + c);
        float d = ev - xv;
        acc = fmaf(d, d, acc);
        if (orow) orow[c] = xv + d;
        if (drow) atomicAdd(drow + c, (zr ? __ldg(zr + c) : xv) - ev);
      }
    }
    if (lane == 0 && counts) {
      if (use_hist) atomicAdd(&hist[k], 1);
      else atomicAdd(counts + k, 1);
    }
  }
  if (sse) {
    double s = warp_sum((double)acc);
    if (lane == 0) wsum[warp] = s;
  }
  __syncthreads();
  if (sse && threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < APPLY_WARPS; ++w) s += wsum[w];
    atomicAdd(sse, s);
  }
  if (use_hist && counts) {
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      int v = hist[k];
      if (v) atomicAdd(counts + k, v);
    }
  }
}

// Same pass with the EMA sums aggregated before they reach memory.  A warp takes 32 consecutive rows, sorts
// their (code, row) pairs with a shuffle network and walks the rows code by code: the residuals of a run of
// rows with the same code are summed in registers and leave as ONE vector reduction per run (and the code
// is fetched once per run).  Usage is Zipf-like in practice, so the hot codes -- the ones whose per-row
// atomics serialise in L2 -- collapse the most.  Lane l owns the float4 columns l, l+32, l+64, l+96 of a row
// (D % 4 == 0, D <= 512).
constexpr int RUNS_MAX_D = 512;

__device__ __forceinline__ unsigned warp_sort_u32(unsigned v, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const unsigned o = __shfl_xor_sync(0xffffffffu, v, j);
      const bool up = ((lane & k) == 0) == ((lane & j) == 0);     // keep the smaller of the pair?
      v = up ? min(v, o) : max(v, o);
    }
  }
  return v;
}

// The rows of a batch are fetched ahead of use with cp.async into a per-warp ring in shared memory (RUNS_PF rows
// in flight per warp: 16 warps x 4 x 1.6 KB per SM covers the HBM latency-bandwidth product; one row held in
// registers did not).  A lane reads back exactly the 16-byte slots it wrote itself, so the ring needs no barrier --
// cp.async.wait_group orders a lane's own copies, and a slot is re-filled one full iteration after its last read.
#ifndef G2V_RUNS_PF
#define G2V_RUNS_PF 4
#endif
constexpr int RUNS_PF = G2V_RUNS_PF;       // rows in flight per warp
constexpr int RUNS_SLOTS = RUNS_PF + 1;
constexpr int RUNS_SLOT_F4 = RUNS_MAX_D / 4;      // float4 per ring slot (128)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

size_t apply_runs_smem(int K, int use_hist) {
  const size_t hist = use_hist ? ((size_t)K * sizeof(int) + 15) / 16 * 16 : 0;
  return hist + (size_t)APPLY_WARPS * RUNS_SLOTS * RUNS_SLOT_F4 * sizeof(float4);
}

// SORTED: the warp's 32 rows are 32 consecutive positions of `order`, the rows grouped by code on the device
// beforehand (code_sort_kernel) -- runs then span whole batches and the per-run vector reductions, which are what
// bounds the unsorted pass (1 M row-sized reductions cost ~0.77 ms of L2 atomic throughput), all but disappear.
template <bool HAS_ZS, bool SORTED>
__global__ void __launch_bounds__(APPLY_WARPS * 32, 2) apply_runs_kernel(
    const float* __restrict__ x, const float* __restrict__ zs, const float* __restrict__ E,
    const int* __restrict__ idx, const int* __restrict__ order, long long N, int K, int D, float* __restrict__ out,
    double* sse, int* counts, float* dwr, int dwr_replicas, int use_hist) {
  extern __shared__ __align__(16) unsigned char runs_smem[];
  int* hist = reinterpret_cast<int*>(runs_smem);
  __shared__ double wsum[APPLY_WARPS];
  dwr += (size_t)(blockIdx.x % dwr_replicas) * K * D;               // this block's private copy
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4* ring = reinterpret_cast<float4*>(runs_smem + (use_hist ? ((size_t)K * sizeof(int) + 15) / 16 * 16 : 0)) +
                 (size_t)warp * RUNS_SLOTS * RUNS_SLOT_F4;
  if (use_hist) {
    for (int k = threadIdx.x; k < K; k += blockDim.x) hist[k] = 0;
    __syncthreads();
  }
  const int nq = D >> 2;                                            // float4 columns per row
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float acc = 0.f;
  const long long stride = (long long)gridDim.x * APPLY_WARPS * 32;
  for (long long base = ((long long)blockIdx.x * APPLY_WARPS + warp) * 32; base < N; base += stride) {
    const int nvalid = (int)min((long long)32, N - base);
    int k = K;                                                      // rows past the end sort last
    long long myrow = base + lane;
    if (lane < nvalid) {
      if (SORTED) myrow = __ldg(order + base + lane);
      k = min(max(__ldg(idx + myrow), 0), K - 1);
      if (!SORTED && counts) {                                      // (the sort has counted the rows already)
        if (use_hist) atomicAdd(&hist[k], 1);
        else atomicAdd(counts + k, 1);
      }
    }
    const unsigned key = SORTED ? (((unsigned)k << 5) | (unsigned)lane) : warp_sort_u32(((unsigned)k << 5) | (unsigned)lane, lane);
    float4 a[4], ev[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = ev[u] = zero4;
    int cur = -1;
    // fetch row i of the sorted batch into its ring slot (a group is committed even when there is no row i, so
    // that wait_group counts the same for every i)
    auto fetch = [&](int i) {
      if (i < nvalid) {
        const unsigned ki = __shfl_sync(0xffffffffu, key, i);
        const long long ri = __shfl_sync(0xffffffffu, myrow, (int)(ki & 31u));
        const float* r = x + (size_t)ri * D;
        float4* slot = ring + (i % RUNS_SLOTS) * RUNS_SLOT_F4;
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (lane + 32 * u < nq) cp_async16(slot + lane + 32 * u, r + 4 * (lane + 32 * u));
      }
      cp_async_commit();
    };
    auto flush = [&]() {
      float* drow = dwr + (size_t)cur * D;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (lane + 32 * u < nq) red_add_v4(drow + 4 * (lane + 32 * u), a[u].x, a[u].y, a[u].z, a[u].w);
    };
#pragma unroll
    for (int i = 0; i < RUNS_PF; ++i) fetch(i);
    for (int i = 0; i < nvalid; ++i) {
      const unsigned ki = __shfl_sync(0xffffffffu, key, i);
      const int code = (int)(ki >> 5);
      const long long row = __shfl_sync(0xffffffffu, myrow, (int)(ki & 31u));
      float4 zv[4];
      if (HAS_ZS) {
        const float* r = zs + (size_t)row * D;
#pragma unroll
        for (int u = 0; u < 4; ++u) zv[u] = (lane + 32 * u < nq) ? lds4(r + 4 * (lane + 32 * u)) : zero4;
      }
      if (code != cur) {                                            // warp-uniform
        if (cur >= 0) flush();
        cur = code;
        const float* er = E + (size_t)code * D;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          ev[u] = (lane + 32 * u < nq) ? ldg4(er + 4 * (lane + 32 * u)) : zero4;
          a[u] = zero4;
        }
      }
      cp_async_wait<RUNS_PF - 1>();                                 // row i has landed (this lane's part of it)
      const float4* slot = ring + (i % RUNS_SLOTS) * RUNS_SLOT_F4;
      float4 xv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) xv[u] = (lane + 32 * u < nq) ? slot[lane + 32 * u] : zero4;
      fetch(i + RUNS_PF);                                           // into the slot read one iteration ago
      float* orow = out ? out + (size_t)row * D : nullptr;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 d = make_float4(ev[u].x - xv[u].x, ev[u].y - xv[u].y, ev[u].z - xv[u].z, ev[u].w - xv[u].w);
        acc = fmaf(d.x, d.x, acc); acc = fmaf(d.y, d.y, acc);
        acc = fmaf(d.z, d.z, acc); acc = fmaf(d.w, d.w, acc);
        if (orow && lane + 32 * u < nq)
          __stcs(reinterpret_cast<float4*>(orow + 4 * (lane + 32 * u)),
                 make_float4(xv[u].x + d.x, xv[u].y + d.y, xv[u].z + d.z, xv[u].w + d.w));
        if (HAS_ZS) {
          a[u].x += zv[u].x - ev[u].x; a[u].y += zv[u].y - ev[u].y;
          a[u].z += zv[u].z - ev[u].z; a[u].w += zv[u].w - ev[u].w;
        } else {
          a[u].x -= d.x; a[u].y -= d.y; a[u].z -= d.z; a[u].w -= d.w;
        }
      }
    }
    cp_async_wait<0>();                                             // (only empty groups are left)
    if (cur >= 0) flush();
  }
  if (sse) {
    double s = warp_sum((double)acc);
    if (lane == 0) wsum[warp] = s;
  }
  __syncthreads();
  if (sse && threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < APPLY_WARPS; ++w) s += wsum[w];
    atomicAdd(sse, s);
  }
  if (use_hist && counts) {
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      int v = hist[k];
      if (v) atomicAdd(counts + k, v);
    }
  }
}

// Group the rows by code: order[0 .. N) lists the row indices code by code (within a code in no particular order).
// One cooperative launch: (1) per-block histograms of contiguous row chunks -> total[K]; (2) after a grid sync every
// block scans total[] into the segment starts; (3) each block claims, per code, a range inside the segment with ONE
// global atomic (cursor[k] += its count) and places its rows there with shared-memory counters -- a hot code costs a
// block one global atomic, not one per row.  total and cursor are zero on entry.
constexpr int SORT_THREADS = 256;
constexpr int SORT_MAX_K = 4096;             // 2 K ints of shared memory

__global__ void __launch_bounds__(SORT_THREADS) code_sort_kernel(const int* __restrict__ idx, long long N, int K, int* total,
                                                                 int* cursor, int* __restrict__ order, int* counts) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  extern __shared__ int sort_sm[];
  int* hist = sort_sm;                         // [K] this block's rows per code, later its write position per code
  int* segs = sort_sm + K;                     // [K] first position of every code
  __shared__ int wsum[SORT_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long per = (N + gridDim.x - 1) / gridDim.x;
  const long long r0 = (long long)blockIdx.x * per, r1 = min(N, r0 + per);
  for (int k = tid; k < K; k += SORT_THREADS) hist[k] = 0;
  __syncthreads();
  for (long long r = r0 + tid; r < r1; r += SORT_THREADS) atomicAdd(&hist[min(max(__ldg(idx + r), 0), K - 1)], 1);
  __syncthreads();
  for (int k = tid; k < K; k += SORT_THREADS)
    if (hist[k]) atomicAdd(total + k, hist[k]);
  grid.sync();
  // exclusive scan of total[0 .. K): thread t owns a contiguous piece
  const int piece = (K + SORT_THREADS - 1) / SORT_THREADS;
  const int k0 = min(tid * piece, K), k1 = min(k0 + piece, K);
  int mine = 0;
  for (int k = k0; k < k1; ++k) mine += __ldcg(total + k);
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  int before = incl - mine;
  for (int w = 0; w < warp; ++w) before += wsum[w];
  for (int k = k0; k < k1; ++k) {
    const int c = __ldcg(total + k);
    segs[k] = before;
    before += c;
    if (blockIdx.x == 0 && counts && c) counts[k] += c;
  }
  __syncthreads();
  for (int k = tid; k < K; k += SORT_THREADS) {
    const int c = hist[k];
    hist[k] = c ? segs[k] + atomicAdd(cursor + k, c) : 0;
  }
  __syncthreads();
  for (long long r = r0 + tid; r < r1; r += SORT_THREADS) {
    const int k = min(max(__ldg(idx + r), 0), K - 1);
    order[atomicAdd(&hist[k], 1)] = (int)r;
  }
}

size_t sorted_ws_bytes(int64_t N, int K) {
  if (N < 16384 || K > SORT_MAX_K || N >= (1ll << 31)) return 0;
  return ((size_t)2 * K * sizeof(int) + 255) / 256 * 256 + (size_t)N * sizeof(int);
}

// rows [N, D] -> [N, Dp] with zeros in the extra columns (the raw rows a folded codebook is searched with)
__global__ void __launch_bounds__(256) pad_rows_kernel(const float* __restrict__ x, long long N, int D, int Dp,
                                                       float* __restrict__ out) {
  const int nq = Dp >> 2, dq = D >> 2;                     // D % 4 == 0, Dp % 4 == 0
  const long long total = N * nq;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / nq;
    const int c = (int)(i - r * nq);
    const float4 v = c < dq ? __ldcs(reinterpret_cast<const float4*>(x + (size_t)r * D) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    __stcs(reinterpret_cast<float4*>(out + (size_t)r * Dp) + c, v);
  }
}

// ------------------------------------------------------------------------------------------
// deterministic statistics: no atomics, fixed summation order, fp64 accumulation
// ------------------------------------------------------------------------------------------
// The fast row pass accumulates the per-code residual sums with fp32 reductions whose order depends on
// scheduling, so two runs of the same step agree to a few ulp, not bitwise.  This pair of kernels computes the
// same statistics reproducibly: the rows come sorted by code (`order`: stable sort, ties by row index; `seg`:
// first position of every code), each code's segment is cut into chunks of DET_CHUNK consecutive positions,
// one block sums a chunk's rows one after the other in fp64 (threads = columns; the last column slot is the
// chunk's sum of squared errors), and a second kernel adds every code's chunk partials in ascending order and
// rounds once to fp32.
constexpr int DET_CHUNK = 128;

__global__ void __launch_bounds__(256) det_partial_kernel(const float* __restrict__ x, const float* __restrict__ zs,
                                                          const float* __restrict__ E, const int* __restrict__ order,
                                                          const long long* __restrict__ seg,
                                                          const long long* __restrict__ chunk_off, int K, int D,
                                                          double* __restrict__ partial) {
  __shared__ double red[8];
  const long long total = chunk_off[K];
  for (long long b = blockIdx.x; b < total; b += gridDim.x) {
    int lo = 0, hi = K;                                    // the code whose chunk range contains b
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (chunk_off[mid] <= b) lo = mid; else hi = mid;
    }
    const int k = lo;
    const long long p0 = seg[k] + (b - chunk_off[k]) * DET_CHUNK, p1 = min(seg[k + 1], p0 + DET_CHUNK);
    const float* er = E + (size_t)k * D;
    double sq = 0.0;
    for (int j = threadIdx.x; j < D; j += blockDim.x) {
      const float ev = er[j];
      double acc = 0.0;
      for (long long p = p0; p < p1; ++p) {
        const size_t r = (size_t)order[p] * D + j;
        const float xv = x[r];
        acc += (double)((zs ? zs[r] : xv) - ev);
        const double dd = (double)(ev - xv);
        sq += dd * dd;
      }
      partial[(size_t)b * (D + 1) + j] = acc;
    }
    // the chunk's squared error: per-thread sums combined in a fixed order
    sq = warp_sum(sq);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += red[w];
      partial[(size_t)b * (D + 1) + D] = t;
    }
  }
}

__global__ void __launch_bounds__(256) det_reduce_kernel(const double* __restrict__ partial,
                                                         const long long* __restrict__ chunk_off, int K, int D,
                                                         float* __restrict__ dwr, double* __restrict__ sse_code) {
  for (int k = blockIdx.x; k < K; k += gridDim.x) {
    const long long c0 = chunk_off[k], c1 = chunk_off[k + 1];
    for (int j = threadIdx.x; j <= D; j += blockDim.x) {
      double acc = 0.0;
      for (long long c = c0; c < c1; ++c) acc += partial[(size_t)c * (D + 1) + j];
      if (j < D) dwr[(size_t)k * D + j] = (float)acc;
      else sse_code[k] = acc;
    }
  }
}

__global__ void det_sse_kernel(const double* __restrict__ sse_code, int K, double* sse) {
  double t = 0.0;
  for (int k = 0; k < K; ++k) t += sse_code[k];            // ascending code order, one thread
  *sse = t;
}

template <bool VEC>
__global__ void __launch_bounds__(256) backward_kernel(const float* __restrict__ x, const float* __restrict__ E,
                                                       const int* __restrict__ idx,
                                                       const float* __restrict__ g_out,
                                                       const float* __restrict__ g_loss, float coef_x,
                                                       long long N, int K, int D, float* __restrict__ g_x) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float c = __ldg(g_loss) * coef_x;
  const long long stride = (long long)gridDim.x * 8;
  for (long long row = (long long)blockIdx.x * 8 + warp; row < N; row += stride) {
    int k = __ldg(idx + row);
    k = min(max(k, 0), K - 1);
    const float* xr = x + (size_t)row * D;
    const float* er = E + (size_t)k * D;
    const float* gr = g_out ? g_out + (size_t)row * D : nullptr;
    float* o = g_x + (size_t)row * D;
    if (VEC) {
      for (int j = lane * 4; j < D; j += 128) {
        float4 xv = lds4(xr + j), ev = ldg4(er + j);
        float4 g = gr ? lds4(gr + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 r = make_float4(fmaf(c, xv.x - ev.x, g.x), fmaf(c, xv.y - ev.y, g.y),
                               fmaf(c, xv.z - ev.z, g.z), fmaf(c, xv.w - ev.w, g.w));
        __stcs(reinterpret_cast<float4*>(o + j), r);
      }
    } else {
      for (int j = lane; j < D; j += 32) {
        float g = gr ? __ldg(gr + j) : 0.f;
        o[j] = fmaf(c, __ldg(xr + j) - __ldg(er + j), g);
      }
    }
  }
}

// The same over the matrix as ONE stream of float4 (rows are D / 4 of them): a CTA takes 256 * BW_U consecutive
// float4 per pass, every thread issues all of its loads (x, g_out, the code entry from L1 / L2) before the first
// store, no lane idles on a row tail (D = 400 is 100 float4: a warp-per-row loop leaves 28 lanes idle every
// fourth pass).  Pure HBM stream: 3 x 4 D bytes per row.
constexpr int BW_U = 4;
__global__ void __launch_bounds__(256) backward_flat_kernel(const float4* __restrict__ x, const float4* __restrict__ E,
                                                            const int* __restrict__ idx, const float4* __restrict__ g_out,
                                                            const float* __restrict__ g_loss, float coef_x,
                                                            unsigned total4, int K, unsigned D4, float4* __restrict__ g_x) {
  const float c = __ldg(g_loss) * coef_x;
  const unsigned span = 256u * BW_U;
  for (unsigned long long base = (unsigned long long)blockIdx.x * span; base < total4; base += (unsigned long long)gridDim.x * span) {
    float4 xv[BW_U], gv[BW_U], ev[BW_U];
#pragma unroll
    for (int u = 0; u < BW_U; ++u) {
      const unsigned long long i64 = base + 256u * u + threadIdx.x;
      if (i64 < total4) {
        const unsigned i = (unsigned)i64;
        const unsigned row = i / D4, col = i - row * D4;
        const int k = min(max(__ldg(idx + row), 0), K - 1);
        xv[u] = __ldcs(x + i);
        gv[u] = g_out ? __ldcs(g_out + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        ev[u] = __ldg(E + (size_t)k * D4 + col);
      }
    }
#pragma unroll
    for (int u = 0; u < BW_U; ++u) {
      const unsigned long long i64 = base + 256u * u + threadIdx.x;
      if (i64 < total4) {
        const float4 r = make_float4(fmaf(c, xv[u].x - ev[u].x, gv[u].x), fmaf(c, xv[u].y - ev[u].y, gv[u].y),
                                     fmaf(c, xv[u].z - ev[u].z, gv[u].z), fmaf(c, xv[u].w - ev[u].w, gv[u].w));
        __stcs(g_x + i64, r);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// scalar outputs / EMA / misc
// ------------------------------------------------------------------------------------------
__global__ void stats_pack_kernel(const int* __restrict__ counts, const double* __restrict__ sse,
                                  const float* __restrict__ dwr, int reps, long long N, int K, int D,
                                  float* __restrict__ packed) {
  const size_t KD = (size_t)K * D;
  float* tail = packed + KD;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  for (size_t k = tid; k < (size_t)K + 2; k += nth) {
    if (k < (size_t)K) tail[k] = counts ? (float)counts[k] : 0.f;
    else if (k == (size_t)K) tail[k] = sse ? (float)(*sse) : 0.f;
    else tail[k] = (float)N;
  }
  for (size_t i = tid; i < KD; i += nth) {                 // sum the replicas in a fixed order
    float a = 0.f;
    for (int r = 0; r < reps; ++r) a += dwr[(size_t)r * KD + i];
    packed[i] = a;
  }
}

__global__ void grad_codebook_kernel(const float* __restrict__ dwr, const float* __restrict__ g_loss,
                                     float coef_e, size_t total, float* __restrict__ g_E) {
  const float c = -__ldg(g_loss) * coef_e;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    g_E[i] = c * dwr[i];
}

template <bool VEC>
__global__ void __launch_bounds__(256) onehot_kernel(const int* __restrict__ idx, long long N, int K,
                                                     float* __restrict__ enc) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long stride = (long long)gridDim.x * 8;
  for (long long row = (long long)blockIdx.x * 8 + warp; row < N; row += stride) {
    const int k = __ldg(idx + row);
    float* o = enc + (size_t)row * K;
    if (VEC) {
      for (int c = lane * 4; c < K; c += 128) {
        float4 v = make_float4(c == k ? 1.f : 0.f, c + 1 == k ? 1.f : 0.f, c + 2 == k ? 1.f : 0.f,
                               c + 3 == k ? 1.f : 0.f);
        __stcs(reinterpret_cast<float4*>(o + c), v);
      }
    } else {
      for (int c = lane; c < K; c += 32) o[c] = (c == k) ? 1.f : 0.f;
    }
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
inline int grid_for(long long work_items, int per_block, int cap_mult) {
  long long g = (work_items + per_block - 1) / per_block;
  long long cap = (long long)num_sms() * cap_mult;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
// Exact re-rank of a SMALL number of listed rows: one CTA per row, warps split the codes, lanes split
// the dimensions (coalesced), straight fp64.  Used for the first kFull64Cap listed rows (the usual
// case: a few hundred rows per million); anything beyond goes to the batched fp32+fp64 kernel above.

template <typename ZT>
__global__ void __launch_bounds__(256) full64_kernel(const ZT* __restrict__ z, const float* __restrict__ E, int K, int D, int Dz,
                                                     const int* __restrict__ list, const int* __restrict__ count,
                                                     int* __restrict__ idx_out, unsigned long long* stats) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* zs = reinterpret_cast<float*>(smem_raw);            // the row, zero padded to a multiple of 32
  __shared__ double bv[8];
  __shared__ int bi[8];
  const int n = min(*count, kFull64Cap);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Dp = (D + 31) & ~31;
  constexpr int KC4 = 4;                                     // codes in flight per warp (latency hiding)
  for (int e = blockIdx.x; e < n; e += gridDim.x) {
    const int row = list[e];
    __syncthreads();
    for (int j = threadIdx.x; j < Dp; j += blockDim.x) zs[j] = j < Dz ? to_f32(z[(size_t)row * Dz + j]) : 0.f;
    __syncthreads();
    double best = INFINITY;
    int besti = 0x7fffffff;
    for (int kb = warp * KC4; kb < K; kb += 8 * KC4) {
      double s[KC4];
#pragma unroll
      for (int c = 0; c < KC4; ++c) s[c] = 0.0;
      for (int j0 = lane; j0 < Dp; j0 += 32 * 4) {
        float ev[KC4][4];
#pragma unroll
        for (int c = 0; c < KC4; ++c) {
          const float* er = E + (size_t)min(kb + c, K - 1) * D;
#pragma unroll
          for (int u = 0; u < 4; ++u) ev[c][u] = (j0 + 32 * u < D) ? __ldg(er + j0 + 32 * u) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = j0 + 32 * u;
          if (j < Dp) {
            const double zv = (double)zs[j];
#pragma unroll
            for (int c = 0; c < KC4; ++c) {
              const double df = zv - (double)ev[c][u];
              s[c] = fma(df, df, s[c]);
            }
          }
        }
      }
#pragma unroll
      for (int c = 0; c < KC4; ++c) {
        const double t = warp_sum(s[c]);
        if (kb + c < K && t < best) { best = t; besti = kb + c; }     // ascending k inside a warp: first wins
      }
    }
    if (lane == 0) { bv[warp] = best; bi[warp] = besti; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double v = bv[0];
      int id = bi[0];
      for (int w = 1; w < 8; ++w)
        if (bv[w] < v || (bv[w] == v && bi[w] < id)) { v = bv[w]; id = bi[w]; }
      idx_out[row] = id;
    }
  }
  if (stats && blockIdx.x == 0 && threadIdx.x == 0 && n > 0) atomicAdd(stats + G2V_STAT_FULL_RECHECK, (unsigned long long)n);
}

template <typename ZT, bool VEC, int R, int NP>
static int launch_full_recheck_v(const ZT* z, const float* E, const void* cb, int K, int D, int Dz, const int32_t* list,
                                 const int32_t* count, int64_t max_rows, int32_t* idx,
                                 unsigned long long* stats, cudaStream_t st, int64_t handled) {
  const size_t smem = (size_t)R * ((D + 3) / 4 * 4) * sizeof(float);
  long long batches = (max_rows - handled + R - 1) / R;
  long long cap = (long long)num_sms() * 4;
  const int grid = (int)(batches < 1 ? 1 : (batches < cap ? batches : cap));
  const float* ntab = reinterpret_cast<const float*>(reinterpret_cast<const char*>(cb) + cb_tab_offset());
  const float* e2 = reinterpret_cast<const float*>(reinterpret_cast<const char*>(cb) + cb_e2_offset());
  if (smem > 40 * 1024)
    G2V_CUDA_CHECK(cudaFuncSetAttribute(full_recheck_kernel<ZT, VEC, R, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  full_recheck_kernel<ZT, VEC, R, NP><<<grid, FR_THREADS, smem, st>>>(z, E, e2, ntab, K, D, list, count, (int)handled, idx, stats, Dz);
  G2V_LAUNCH_CHECK("full_recheck_kernel");
  return G2V_OK;
}

template <typename ZT>
static int launch_full_recheck_t(const ZT* z, const float* E, const void* cb, int K, int D, int Dz, const int32_t* list,
                                 const int32_t* count, int64_t max_rows, int32_t* idx,
                                 unsigned long long* stats, bool overflow_only, cudaStream_t st, int64_t handled) {
  if (overflow_only) {
    if (max_rows <= handled) return G2V_OK;
  } else {
    handled = kFull64Cap;
    const long long cap = max_rows < kFull64Cap ? max_rows : kFull64Cap;
    const long long lim = (long long)num_sms() * 8;
    const int g = (int)(cap < 1 ? 1 : (cap < lim ? cap : lim));
    full64_kernel<ZT><<<g, 256, (size_t)((D + 31) / 32 * 32) * sizeof(float), st>>>(z, E, K, D, Dz, list, count, idx, stats);
    G2V_LAUNCH_CHECK("full64_kernel");
    if (max_rows <= kFull64Cap) return G2V_OK;               // nothing can be left for the batched kernel
  }
  const bool vec = (D % 4 == 0) && aligned16(E);
  // small codebooks: 8 rows per CTA (more CTAs in flight, 4 partial sums -> tight fp32 bound);
  // large codebooks: 16 rows per CTA so each codebook row fetched from L2 serves more latents
  if (!vec) return launch_full_recheck_v<ZT, false, 8, 1>(z, E, cb, K, D, Dz, list, count, max_rows, idx, stats, st, handled);
  if (K <= 2048) return launch_full_recheck_v<ZT, true, 8, 4>(z, E, cb, K, D, Dz, list, count, max_rows, idx, stats, st, handled);
  return launch_full_recheck_v<ZT, true, 16, 1>(z, E, cb, K, D, Dz, list, count, max_rows, idx, stats, st, handled);
}

int launch_full_recheck(const void* z, int z_dtype, const float* E, const void* cb, int K, int D, int Dz, const int32_t* list,
                        const int32_t* count, int64_t max_rows, int32_t* idx, unsigned long long* stats,
                        bool overflow_only, cudaStream_t st, int64_t handled) {
  switch (z_dtype) {
    case G2V_F32: return launch_full_recheck_t(reinterpret_cast<const float*>(z), E, cb, K, D, Dz, list, count, max_rows, idx, stats, overflow_only, st, handled);
    case G2V_F16: return launch_full_recheck_t(reinterpret_cast<const __half*>(z), E, cb, K, D, Dz, list, count, max_rows, idx, stats, overflow_only, st, handled);
    case G2V_BF16: return launch_full_recheck_t(reinterpret_cast<const __nv_bfloat16*>(z), E, cb, K, D, Dz, list, count, max_rows, idx, stats, overflow_only, st, handled);
    default: return G2V_ERR_DTYPE;
  }
}

template <typename ZT>
static int launch_search_simt_t(const ZT* z, int z_dtype, const float* E, const void* cb, int64_t N, int K, int D,
                                int32_t* full_list, int32_t* full_count, int32_t* idx,
                                unsigned long long* stats, cudaStream_t st) {
  const float* ntab = reinterpret_cast<const float*>(reinterpret_cast<const char*>(cb) + cb_tab_offset());
  const float* e2 = reinterpret_cast<const float*>(reinterpret_cast<const char*>(cb) + cb_e2_offset());
  // vector path: 8 elements per load must stay 16-byte aligned in every row
  const bool vec = ((size_t)D * sizeof(ZT) % 16 == 0) && (D % 4 == 0) && aligned16(z) && aligned16(E);
  const size_t smem = sizeof(SearchSmem);
  G2V_CUDA_CHECK(cudaMemsetAsync(full_count, 0, sizeof(int32_t), st));
  cudaEvent_t pev0, pev1;
  profile_take(&pev0, &pev1);
  if (pev0) G2V_CUDA_CHECK(cudaEventRecord(pev0, st));
  // persistent grid: two CTAs per SM, each walks row tiles
  long long tiles = (N + BM - 1) / BM;
  int grid = (int)((tiles < (long long)num_sms() * 2) ? (tiles > 0 ? tiles : 1) : (long long)num_sms() * 2);
  if (vec) {
    G2V_CUDA_CHECK(cudaFuncSetAttribute(search_simt_kernel<true, ZT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    search_simt_kernel<true, ZT><<<grid, NT, smem, st>>>(z, E, e2, ntab, N, K, D, full_list, full_count, idx);
  } else {
    G2V_CUDA_CHECK(cudaFuncSetAttribute(search_simt_kernel<false, ZT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    search_simt_kernel<false, ZT><<<grid, NT, smem, st>>>(z, E, e2, ntab, N, K, D, full_list, full_count, idx);
  }
  G2V_LAUNCH_CHECK("search_simt_kernel");
  if (pev1) G2V_CUDA_CHECK(cudaEventRecord(pev1, st));
  return launch_full_recheck_t(z, E, cb, K, D, D, full_list, full_count, N, idx, stats, false, st, kFull64Cap);
}

// fp32 search of all rows; `full_list` (N ints) and `full_count` (1 int) are scratch
int launch_search_simt(const void* z, int z_dtype, const float* E, const void* cb, int64_t N, int K, int D,
                       int32_t* full_list, int32_t* full_count, int32_t* idx,
                       unsigned long long* stats, cudaStream_t st) {
  switch (z_dtype) {
    case G2V_F32: return launch_search_simt_t(reinterpret_cast<const float*>(z), z_dtype, E, cb, N, K, D, full_list, full_count, idx, stats, st);
    case G2V_F16: return launch_search_simt_t(reinterpret_cast<const __half*>(z), z_dtype, E, cb, N, K, D, full_list, full_count, idx, stats, st);
    case G2V_BF16: return launch_search_simt_t(reinterpret_cast<const __nv_bfloat16*>(z), z_dtype, E, cb, N, K, D, full_list, full_count, idx, stats, st);
    default: return G2V_ERR_DTYPE;
  }
}

size_t apply_ws_bytes(int64_t N, int K) { return sorted_ws_bytes(N, K); }

int launch_apply(const float* x, const float* zs, const float* E, const int32_t* idx, int64_t N, int K, int D,
                 float* out, double* sse, int32_t* counts, float* dwr, int dwr_replicas, cudaStream_t st, void* ws,
                 size_t ws_bytes) {
  const bool vec = (D % 4 == 0) && aligned16(x) && aligned16(E) && (!zs || aligned16(zs)) &&
                   (!out || aligned16(out)) && (!dwr || aligned16(dwr));
  const int use_hist = (counts && K <= 8192) ? 1 : 0;
  const size_t smem = use_hist ? (size_t)K * sizeof(int) : 0;
  const int grid = grid_for(N, APPLY_WARPS, 8);
  // EMA / codebook-gradient sums wanted: aggregate runs of equal codes in registers first (G2V_APPLY_RUNS=0:
  // one reduction per row, the older kernel)
  static const bool runs_on = [] { const char* e = getenv("G2V_APPLY_RUNS"); return !(e && atoi(e) == 0); }();
  static const bool sorted_on = [] { const char* e = getenv("G2V_APPLY_SORTED"); return !(e && atoi(e) == 0); }();
  // (small batches: a warp per row keeps every SM busy; the run walk serialises 32 rows per warp)
  if (vec && dwr && D <= RUNS_MAX_D && K < (1 << 26) && runs_on && N >= 16384) {      // sort key = code << 5 | lane
    const size_t need = sorted_ws_bytes(N, K);
    const bool sorted = sorted_on && ws && need && ws_bytes >= need && aligned16(ws);
    const int g = grid_for((N + 31) / 32, APPLY_WARPS, 2);          // two resident CTAs per SM (registers, ring)
    const int* order = nullptr;
    if (sorted) {
      int* total = reinterpret_cast<int*>(ws);
      int* cursor = total + K;
      int* ord = reinterpret_cast<int*>(reinterpret_cast<char*>(ws) + ((size_t)2 * K * sizeof(int) + 255) / 256 * 256);
      G2V_CUDA_CHECK(cudaMemsetAsync(ws, 0, (size_t)2 * K * sizeof(int), st));
      const size_t ssm = (size_t)2 * K * sizeof(int);
      int per_sm = 0;
      G2V_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, code_sort_kernel, SORT_THREADS, ssm));
      if (per_sm < 1) per_sm = 1;
      const int sgrid = (int)std::max<long long>(1, std::min<long long>((long long)per_sm * num_sms(), (N + 2047) / 2048));
      long long n64 = N;
      int k32 = K;
      int32_t* cnt = counts;
      void* args[] = {(void*)&idx, (void*)&n64, (void*)&k32, (void*)&total, (void*)&cursor, (void*)&ord, (void*)&cnt};
      G2V_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)code_sort_kernel, dim3(sgrid), dim3(SORT_THREADS), args, ssm, st));
      G2V_LAUNCH_CHECK("code_sort_kernel");
      order = ord;
    }
    const size_t rsmem = apply_runs_smem(K, sorted ? 0 : use_hist);
    if (rsmem <= 200 * 1024) {
      // (the attribute is per device and a call sets it: always the same value, so concurrent callers agree)
#define G2V_RUNS_LAUNCH(ZS, SO)                                                                                          \
  do {                                                                                                                  \
    G2V_CUDA_CHECK(cudaFuncSetAttribute(apply_runs_kernel<ZS, SO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
    apply_runs_kernel<ZS, SO><<<g, APPLY_WARPS * 32, rsmem, st>>>(x, zs, E, idx, order, N, K, D, out, sse, counts, dwr,   \
                                                                  dwr_replicas, SO ? 0 : use_hist);                      \
  } while (0)
      if (zs && sorted) G2V_RUNS_LAUNCH(true, true);
      else if (zs) G2V_RUNS_LAUNCH(true, false);
      else if (sorted) G2V_RUNS_LAUNCH(false, true);
      else G2V_RUNS_LAUNCH(false, false);
#undef G2V_RUNS_LAUNCH
      G2V_LAUNCH_CHECK("apply_runs_kernel");
      return G2V_OK;
    }
  }
  if (vec)
    apply_kernel<true><<<grid, APPLY_WARPS * 32, smem, st>>>(x, zs, E, idx, N, K, D, out, sse, counts, dwr, dwr_replicas, use_hist);
  else
    apply_kernel<false><<<grid, APPLY_WARPS * 32, smem, st>>>(x, zs, E, idx, N, K, D, out, sse, counts, dwr, dwr_replicas, use_hist);
  G2V_LAUNCH_CHECK("apply_kernel");
  return G2V_OK;
}

int launch_pad_rows(const float* x, int64_t N, int D, int Dp, float* out, cudaStream_t st) {
  pad_rows_kernel<<<grid_for(N * (Dp / 4), 256 * 4, 16), 256, 0, st>>>(x, N, D, Dp, out);
  G2V_LAUNCH_CHECK("pad_rows_kernel");
  return G2V_OK;
}

int launch_stats_deterministic(const float* x, const float* zs, const float* E, const int32_t* order, const long long* seg,
                               const long long* chunk_off, int64_t max_chunks, int K, int D, double* partial, float* dwr,
                               double* sse_code, double* sse, cudaStream_t st) {
  const long long cap = (long long)num_sms() * 16;
  const int g1 = (int)std::max<long long>(1, std::min<long long>(max_chunks, cap));
  det_partial_kernel<<<g1, 256, 0, st>>>(x, zs, E, order, seg, chunk_off, K, D, partial);
  G2V_LAUNCH_CHECK("det_partial_kernel");
  det_reduce_kernel<<<std::min(K, num_sms() * 8), 256, 0, st>>>(partial, chunk_off, K, D, dwr, sse_code);
  G2V_LAUNCH_CHECK("det_reduce_kernel");
  det_sse_kernel<<<1, 1, 0, st>>>(sse_code, K, sse);
  G2V_LAUNCH_CHECK("det_sse_kernel");
  return G2V_OK;
}

int launch_stats_pack(const int32_t* counts, const double* sse, const float* dwr, int dwr_replicas, int64_t N,
                      int K, int D, float* packed, cudaStream_t st) {
  stats_pack_kernel<<<grid_for((long long)K * D + K + 2, 256, 8), 256, 0, st>>>(counts, sse, dwr, dwr ? dwr_replicas : 0, N,
                                                                                 K, D, packed);
  G2V_LAUNCH_CHECK("stats_pack_kernel");
  return G2V_OK;
}

int launch_backward(const float* x, const float* E, const int32_t* idx, const float* g_out, const float* g_loss,
                    float coef_x, int64_t N, int K, int D, float* g_x, cudaStream_t st) {
  const bool vec = (D % 4 == 0) && aligned16(x) && aligned16(E) && aligned16(g_x) && (!g_out || aligned16(g_out));
  const int grid = grid_for(N, 8, 8);
  const unsigned long long total4 = (unsigned long long)N * (unsigned long long)(D / 4);
  if (vec && total4 < 0xffffffffull) {
    const int fgrid = (int)std::max<unsigned long long>(1, std::min<unsigned long long>((total4 + 256 * BW_U - 1) / (256 * BW_U),
                                                                                      (unsigned long long)num_sms() * 8));
    backward_flat_kernel<<<fgrid, 256, 0, st>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(E), idx,
                                               reinterpret_cast<const float4*>(g_out), g_loss, coef_x, (unsigned)total4, K,
                                               (unsigned)(D / 4), reinterpret_cast<float4*>(g_x));
    G2V_LAUNCH_CHECK("backward_flat_kernel");
    return G2V_OK;
  }
  if (vec) backward_kernel<true><<<grid, 256, 0, st>>>(x, E, idx, g_out, g_loss, coef_x, N, K, D, g_x);
  else backward_kernel<false><<<grid, 256, 0, st>>>(x, E, idx, g_out, g_loss, coef_x, N, K, D, g_x);
  G2V_LAUNCH_CHECK("backward_kernel");
  return G2V_OK;
}

int launch_grad_codebook(const float* dwr, const float* g_loss, float coef_e, int K, int D, float* g_E,
                         cudaStream_t st) {
  const size_t total = (size_t)K * D;
  grad_codebook_kernel<<<grid_for((long long)total, 256, 8), 256, 0, st>>>(dwr, g_loss, coef_e, total, g_E);
  G2V_LAUNCH_CHECK("grad_codebook_kernel");
  return G2V_OK;
}

int launch_onehot(const int32_t* idx, int64_t N, int K, float* enc, cudaStream_t st) {
  const bool vec = (K % 4 == 0) && aligned16(enc);
  const int grid = grid_for(N, 8, 16);
  if (vec) onehot_kernel<true><<<grid, 256, 0, st>>>(idx, N, K, enc);
  else onehot_kernel<false><<<grid, 256, 0, st>>>(idx, N, K, enc);
  G2V_LAUNCH_CHECK("onehot_kernel");
  return G2V_OK;
}

}  // namespace g2v
