// Shared declarations for the sm_100a vector-quantizer kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/g2v_vq.h"

namespace g2v {

// ---------------------------------------------------------------------------------------------
// Codebook aux buffer ("cb"): header + norm table + ||e||^2 + fp16 operand copy.  Layout in bytes:
//   [0, 256)                       CbHeader
//   [256, 768)                     ntab[128] fp32: ntab[b] = largest ||e_k|| among codes whose norm falls in
//                                  quarter-octave bucket <= b (see norm_bucket) -- "reachable norm" lookup
//   [768, 768 + 4*Kp)              e2[Kp] fp32 (fp64-accumulated, rounded once; +inf for k >= K)
//   [off16, off16 + 2*Kp*Dp)       E16[Kp][Dp] fp16, zero padded, scaled by header.scale_e
// Kp = K rounded up to 256, Dp = D rounded up to 16.
//
// Why the table: every error bound is linear in the norm of the codes that could still win a row.
// A code can only win if ||e_k|| <= ||z|| + sqrt(d_min), so bounds use the largest norm below that
// reach instead of the global maximum -- EMA codebooks carry dead codes with norms ~1e5 that would
// otherwise make every row look uncertain.
// ---------------------------------------------------------------------------------------------
struct CbHeader {
  float e2max;      // max_k ||e_k||^2
  float e2min;      // min_k ||e_k||^2
  float amax;       // max |e_kj|
  float scale_e;    // power of two the fp16 copy was multiplied by (brings amax into [16384,32768))
  int K, D, Kp, Dp;
  int magic;
  // measured rounding residual of the fp16 copy, r_k = e_k - fp16(e_k*scale_e)/scale_e, split by where the
  // scaled element landed:  ||r_k|| <= sfrac ||e_k|| + rsub
  float sfrac;      // max_k ||r_k restricted to fp16-normal elements|| / ||e_k||   (<= 2^-12)
  float rsub;       // max_k ||r_k restricted to fp16-subnormal elements||          (absolute: <= sqrt(D) 2^-25 / scale_e)
  int pad[53];
};
static_assert(sizeof(CbHeader) == 256, "CbHeader must be 256 bytes");
constexpr int kCbMagic = 0x67327632;  // "g2v2"
constexpr int kNormBuckets = 128;

__host__ __device__ inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
inline size_t cb_tab_offset() { return 256; }
inline size_t cb_e2_offset() { return 768; }
inline size_t cb_e16_offset(int K) { return 768 + (size_t)round_up(K, 256) * 4; }  // 256-byte aligned
inline size_t cb_total_bytes(int K, int D) {
  return cb_e16_offset(K) + (size_t)round_up(K, 256) * round_up(D, 16) * 2;
}

#ifdef __CUDACC__
// quarter-octave bucket of a norm: 2^-16 .. 2^16 -> 0 .. 127 (clamped)
__device__ __forceinline__ int norm_bucket(float c) {
  int b = (__float_as_int(c) >> 21) - ((127 - 16) << 2);
  return b < 0 ? 0 : (b > kNormBuckets - 1 ? kNormBuckets - 1 : b);
}
// largest code norm that a row with norm `znorm` and (approximate) best squared distance `d1` can
// still be won by: ||e_k|| <= ||z|| + sqrt(d_k) and d_k <= d_min.  The factor 2 on the sqrt and the 1 %
// on ||z|| absorb the error of d1 itself.
__device__ __forceinline__ float reachable_norm(const float* __restrict__ ntab, float znorm, float d1) {
  const float cstar = 1.01f * znorm + 2.f * sqrtf(fmaxf(d1, 0.f)) + 1e-30f;
  return ntab[norm_bucket(cstar)];
}
#endif

// error recording (thread local), defined in g2v_api.cu
void set_error_detail(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define G2V_CUDA_CHECK(expr)                                   \
  do {                                                         \
    cudaError_t _e = (expr);                                   \
    if (_e != cudaSuccess) return ::g2v::cuda_fail(_e, #expr); \
  } while (0)

// every kernel launch of the library passes through here: g_launches is the count g2v_launch_count() reports
extern unsigned long long g_launches;
#define G2V_LAUNCH_CHECK(name)                                  \
  do {                                                          \
    cudaError_t _e = cudaGetLastError();                        \
    if (_e != cudaSuccess) return ::g2v::cuda_fail(_e, name);   \
    __atomic_add_fetch(&::g2v::g_launches, 1ULL, __ATOMIC_RELAXED); \
  } while (0)

int num_sms();
// one-shot profiling events (thread local), see g2v_profile_next_search
void profile_take(cudaEvent_t* start, cudaEvent_t* stop);

// ---- launchers implemented in g2v_simt.cu ----------------------------------------------------
int launch_codebook_prepare(const float* E, int K, int D, void* cb, cudaStream_t st);
// fp32 CUDA-core search over all rows + fp64 re-rank of the rows it cannot certify;
// full_list (N ints) / full_count (1 int) are device scratch.
int launch_search_simt(const void* z, int z_dtype, const float* E, const void* cb, int64_t N, int K, int D,
                       int32_t* full_list, int32_t* full_count, int32_t* idx,
                       unsigned long long* stats, cudaStream_t st);
// exact fp64 argmin of the rows in list[0 .. *count) (count <= max_rows).  The first kFull64Cap listed rows
// take a per-row fp64 kernel, the rest a batched fp32+fp64 kernel; overflow_only = the caller has already
// handled the first `handled` rows itself.
constexpr int kFull64Cap = 4096;
int launch_full_recheck(const void* z, int z_dtype, const float* E, const void* cb, int K, int D, int Dz, const int32_t* list,
                        const int32_t* count, int64_t max_rows, int32_t* idx, unsigned long long* stats,
                        bool overflow_only, cudaStream_t st, int64_t handled = kFull64Cap);
int launch_apply(const float* x, const float* zs, const float* E, const int32_t* idx, int64_t N, int K,
                 int D, float* out, double* sse, int32_t* counts, float* dwr, int dwr_replicas, cudaStream_t st,
                 void* ws = nullptr, size_t ws_bytes = 0);
// scratch of the sorted row pass (0: the shape does not take it)
size_t apply_ws_bytes(int64_t N, int K);
int launch_pad_rows(const float* x, int64_t N, int D, int Dp, float* out, cudaStream_t st);
// deterministic (atomics-free, fixed-order, fp64-accumulated) residual sums and squared error from rows sorted by code
int launch_stats_deterministic(const float* x, const float* zs, const float* E, const int32_t* order, const long long* seg,
                               const long long* chunk_off, int64_t max_chunks, int K, int D, double* partial, float* dwr,
                               double* sse_code, double* sse, cudaStream_t st);
int launch_stats_pack(const int32_t* counts, const double* sse, const float* dwr, int dwr_replicas, int64_t N,
                      int K, int D, float* packed, cudaStream_t st);
int launch_stats_finalize(const float* packed, int K, int D, float coef_codebook, float coef_commit,
                          float* loss, float* ppl, cudaStream_t st);
// g2v_finalize.cu: pack + scalars + codebook update (mode 0 none, 1 EMA, 2 Lloyd) + aux buffer, one cooperative launch
int launch_step_finalize(int32_t* counts, double* sse, float* dwr, int dwr_replicas, int do_pack, int64_t rows_local,
                         float* packed, int K, int D, float coef_codebook, float coef_commit, float* loss, float* ppl,
                         int mode, const float* cs_in, float* cs_out, const float* w_in, float* w_out,
                         const float* E_old, float* E_new, float* E_prev, float decay, float eps, double* shift2,
                         void* cb, const float* E_cb, cudaStream_t st);
int launch_fold_rows(const float* E, const float* W, const float* b, int K, int D, int ld_out, float* out, double* g,
                     cudaStream_t st);
// g2v_audit.cu: exact fp64 argmin of every row (verification aid); ws holds K doubles
int launch_search_exact64(const void* z, int z_dtype, const float* E, int64_t N, int K, int D, int32_t* idx, void* ws,
                          cudaStream_t st);
int launch_backward(const float* x, const float* E, const int32_t* idx, const float* g_out,
                    const float* g_loss, float coef_x, int64_t N, int K, int D, float* g_x,
                    cudaStream_t st);
int launch_grad_codebook(const float* dwr, const float* g_loss, float coef_e, int K, int D, float* g_E,
                         cudaStream_t st);
int launch_onehot(const int32_t* idx, int64_t N, int K, float* enc, cudaStream_t st);

// ---- g2v_gemm.cu: C[M,N] (+)= alpha * op(A)[M,K] * op(B)[N,K]^T + bias, fp32 in / out, split-fp16 tcgen05 GEMM
size_t gemm_workspace_bytes(int64_t M, int N, int64_t K, int single_term);
int launch_gemm_f32(const float* A, int64_t lda, int transA, const float* B, int64_t ldb, int transB, int64_t M, int N,
                    int64_t K, const float* bias, float* C, int64_t ldc, float alpha, int accumulate, int single_term,
                    void* ws, cudaStream_t st);

// operands already in the fp16 layout (row-major [rows, kp16], kp16 % 64 == 0), row count on the device
int launch_gemm_prepared(const __half* A16, const int* m_dev, long long m_cap, const __half* B16, int N, long long kp16,
                         float* C, long long ldc, cudaStream_t st);

// ---- g2v_soft.cu: row kernels of the soft quantizer VQ_Payam_GSSoft
int launch_soft_assign(const float* m, float* dot, const float* lv, const float* e2, int64_t N, int K, int D, float* p,
                       float* colsum, cudaStream_t st);
int launch_soft_tail(const float* x, const float* q, int64_t N, int K, int D, float beta, float* out, double* sse,
                     const float* colsum, float* loss, float* ppl, cudaStream_t st);
int launch_soft_bwd(const float* p, const float* dp, const float* d, const float* lv, int64_t N, int K, float* gd, float* glv,
                    float* rowsum_gd, float* col_gd, float* col_glv, cudaStream_t st);
int launch_soft_gx(const float* gmw, const float* x, const float* q, const float* g_out, const float* c, const float* a,
                   int64_t n, float* gx, cudaStream_t st);

// ---- tensor-core path, implemented in g2v_tc.cu ---------------------------------------------
bool tc_supported(int K, int D);
size_t tc_workspace_bytes(int64_t N, int K, int D, int z_dtype);
int launch_search_tc(const void* z, int z_dtype, const float* E, const void* cb, int64_t N, int K, int D,
                     int32_t* idx, unsigned long long* stats, void* ws, size_t ws_bytes, unsigned flags,
                     cudaStream_t st, int Dz = 0);   // Dz: columns of the rows in memory (0 = D)

}  // namespace g2v
