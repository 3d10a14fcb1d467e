// Exact nearest-code search in fp64 over ALL rows: the device-side checker of the fast paths.
//
//   idx[n] = argmin_k ( ||e_k||^2 - 2 z_n . e_k ), every product and sum in fp64, first index on exact ties
//
// fp32 (or 16-bit) inputs are exact in fp64, their products are exact, and a sum of D <= 512 such terms carries
// a relative error of ~D * 2^-53: this IS exact arithmetic as far as any fp32 distance gap is concerned.  It
// shares no code and no numerical shortcut with the tensor-core / fp32 search kernels (no error bounds, no
// candidate lists, no rounding of operands), so a full-row comparison against it audits those kernels
// independently at sizes the numpy oracle cannot reach (tests/test_gpu_audit.py, tools/audit_exact.py).
// It is FP64-pipe bound (2 K D flop per row at ~40 TFLOP/s: ~10 s per million rows at K = D = 400) and is
// not part of the product path.
//
// Layout: a CTA takes 64 rows and sweeps the codebook in tiles of 128 codes; the D loop runs in chunks of 8
// through double-buffered shared memory holding fp64 copies; a thread owns 4 rows x 8 codes
// (four pairs of adjacent codes, 32 apart, so that the shared-memory reads of a half-warp are contiguous).
#include "g2v_common.cuh"

#include <math.h>

namespace g2v {
namespace {

constexpr int XM = 64, XN = 128, XK = 8, XT = 256;

__device__ __forceinline__ float ldx(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ldx(const __half* p) { return __half2float(*p); }
__device__ __forceinline__ float ldx(const __nv_bfloat16* p) { return __bfloat162float(*p); }

__global__ void __launch_bounds__(256) e2_f64_kernel(const float* __restrict__ E, int K, int D, double* __restrict__ e2) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= K) return;
  double s = 0.0;
  for (int j = lane; j < D; j += 32) {
    const double v = (double)E[(size_t)warp * D + j];
    s = fma(v, v, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) e2[warp] = s;
}

template <typename ZT>
__global__ void __launch_bounds__(XT, 2) exact64_kernel(const ZT* __restrict__ z, const float* __restrict__ E,
                                                        const double* __restrict__ e2, long long N, int K, int D,
                                                        int* __restrict__ idx) {
  __shared__ double As[2][XK][XM];
  __shared__ double Bs[2][XK][XN];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  // loaders: thread -> (row of the z tile, 2 of its 8 columns); (code of the E tile, 4 of its 8 columns)
  const int a_row = t >> 2, a_col = (t & 3) * 2;
  const int b_row = t >> 1, b_col = (t & 1) * 4;
  const int nk = (D + XK - 1) / XK;
  const int n_ctile = (K + XN - 1) / XN;
  for (long long tile = blockIdx.x; tile * XM < N; tile += gridDim.x) {
    const long long grow = tile * XM + a_row;
    const ZT* zr = grow < N ? z + (size_t)grow * D : nullptr;
    double best[4];
    int besti[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { best[i] = INFINITY; besti[i] = 0x7fffffff; }
    for (int ct = 0; ct < n_ctile; ++ct) {
      const int code = ct * XN + b_row;
      const float* er = code < K ? E + (size_t)code * D : nullptr;
      double acc[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
      float ra[2], rb[4];
      auto fetch = [&](int kc) {
        const int c0 = kc * XK;
#pragma unroll
        for (int u = 0; u < 2; ++u) ra[u] = (zr && c0 + a_col + u < D) ? ldx(zr + c0 + a_col + u) : 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) rb[u] = (er && c0 + b_col + u < D) ? __ldg(er + c0 + b_col + u) : 0.f;
      };
      auto stash = [&](int buf) {
#pragma unroll
        for (int u = 0; u < 2; ++u) As[buf][a_col + u][a_row] = (double)ra[u];
#pragma unroll
        for (int u = 0; u < 4; ++u) Bs[buf][b_col + u][b_row] = (double)rb[u];
      };
      __syncthreads();                      // previous code tile / row tile done with shared memory
      fetch(0);
      stash(0);
      __syncthreads();
      for (int kc = 0; kc < nk; ++kc) {
        const int buf = kc & 1;
        if (kc + 1 < nk) fetch(kc + 1);
#pragma unroll
        for (int k = 0; k < XK; ++k) {
          double a[4], b[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) a[i] = As[buf][k][ty * 4 + i];
#pragma unroll
          for (int j = 0; j < 8; ++j) b[j] = Bs[buf][k][32 * (j >> 1) + tx * 2 + (j & 1)];   // 16-byte reads at a 16-byte lane stride
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        if (kc + 1 < nk) stash(buf ^ 1);
        __syncthreads();
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {         // ascending code order inside a thread: strict '<' keeps the first
        const int k = ct * XN + 32 * (j >> 1) + tx * 2 + (j & 1);
        if (k < K) {
          const double ek = e2[k];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const double d = fma(-2.0, acc[i][j], ek);
            if (d < best[i]) { best[i] = d; besti[i] = k; }
          }
        }
      }
    }
    // merge the 16 column threads of a row (a half-warp): value, then lower index
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double ov = __shfl_xor_sync(0xffffffffu, best[i], o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti[i], o);
        if (ov < best[i] || (ov == best[i] && oi < besti[i])) { best[i] = ov; besti[i] = oi; }
      }
    }
    if (tx == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long r = tile * XM + ty * 4 + i;
        if (r < N) idx[r] = besti[i];
      }
    }
  }
}

template <typename ZT>
int run_exact(const ZT* z, const float* E, int64_t N, int K, int D, int32_t* idx, double* e2, cudaStream_t st) {
  e2_f64_kernel<<<(K * 32 + 255) / 256, 256, 0, st>>>(E, K, D, e2);
  G2V_LAUNCH_CHECK("e2_f64_kernel");
  const long long tiles = (N + XM - 1) / XM;
  const long long cap = (long long)num_sms() * 2;
  const int grid = (int)(tiles < cap ? (tiles < 1 ? 1 : tiles) : cap);
  exact64_kernel<ZT><<<grid, XT, 0, st>>>(z, E, e2, N, K, D, idx);
  G2V_LAUNCH_CHECK("exact64_kernel");
  return G2V_OK;
}

}  // namespace

int launch_search_exact64(const void* z, int z_dtype, const float* E, int64_t N, int K, int D, int32_t* idx, void* ws,
                          cudaStream_t st) {
  double* e2 = reinterpret_cast<double*>(ws);
  switch (z_dtype) {
    case G2V_F32: return run_exact(reinterpret_cast<const float*>(z), E, N, K, D, idx, e2, st);
    case G2V_F16: return run_exact(reinterpret_cast<const __half*>(z), E, N, K, D, idx, e2, st);
    case G2V_BF16: return run_exact(reinterpret_cast<const __nv_bfloat16*>(z), E, N, K, D, idx, e2, st);
    default: return G2V_ERR_DTYPE;
  }
}

}  // namespace g2v
