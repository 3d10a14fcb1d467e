// C ABI of the vector-quantizer path (see include/g2v_vq.h).  Argument validation, dispatch
// between the tensor-core and CUDA-core search, workspace carving, and the host-buffer
// end-to-end tokeniser.  No kernels here.
#include "g2v_common.cuh"

#include <stdarg.h>
#include <string.h>

namespace g2v {

static thread_local char g_detail[512] = "";
unsigned long long g_launches = 0;

void set_error_detail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_detail, sizeof(g_detail), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  (void)cudaGetLastError();      // a failed launch / call must not stay pending for the caller's next CUDA call
  set_error_detail("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return G2V_ERR_CUDA;
}

int num_sms() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

static thread_local cudaEvent_t g_prof_start = nullptr, g_prof_stop = nullptr;
void profile_take(cudaEvent_t* start, cudaEvent_t* stop) {
  *start = g_prof_start;
  *stop = g_prof_stop;
  g_prof_start = g_prof_stop = nullptr;
}

static int check_arch() {
  static thread_local int cached_dev = -1, ok = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return cuda_fail(cudaGetLastError(), "cudaGetDevice");
  if (dev != cached_dev) {
    int major = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    ok = (major == 10);
    cached_dev = dev;
  }
  if (!ok) {
    set_error_detail("device is not compute capability 10.x; this library is built for sm_100a only");
    return G2V_ERR_ARCH;
  }
  return G2V_OK;
}

static inline bool bad_shape(int64_t N, int K, int D) { return N < 0 || K <= 0 || D <= 0 || N > 0x7fffffffLL; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int elem_size(int dtype) { return dtype == G2V_F32 ? 4 : 2; }

}  // namespace g2v

using namespace g2v;

extern "C" {

int g2v_version(void) { return G2V_VERSION; }

const char* g2v_strerror(int code) {
  switch (code) {
    case G2V_OK: return "ok";
    case G2V_ERR_INVALID: return "invalid argument";
    case G2V_ERR_ALIGN: return "pointer alignment";
    case G2V_ERR_DTYPE: return "unsupported dtype";
    case G2V_ERR_WORKSPACE: return "workspace too small";
    case G2V_ERR_CUDA: return "CUDA error";
    case G2V_ERR_ARCH: return "unsupported GPU architecture (need sm_100)";
    case G2V_ERR_UNSUPPORTED: return "shape not supported by the selected path";
    default: return "unknown error";
  }
}

const char* g2v_last_error_detail(void) { return g_detail; }

unsigned long long g2v_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

size_t g2v_codebook_bytes(int K, int D) {
  if (K <= 0 || D <= 0) return 0;
  return cb_total_bytes(K, D);
}

int g2v_codebook_prepare(const float* E, int K, int D, void* cb, size_t cb_bytes, void* stream) {
  if (!E || !cb || K <= 0 || D <= 0) return G2V_ERR_INVALID;
  if (cb_bytes < cb_total_bytes(K, D)) return G2V_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(cb) & 255) return G2V_ERR_ALIGN;
  int rc = check_arch();
  if (rc) return rc;
  return launch_codebook_prepare(E, K, D, cb, (cudaStream_t)stream);
}

int g2v_profile_next_search(void* ev_start, void* ev_stop) {
  g_prof_start = (cudaEvent_t)ev_start;
  g_prof_stop = (cudaEvent_t)ev_stop;
  return G2V_OK;
}

int g2v_search_path(int K, int D, unsigned flags) {
  if (K <= 0 || D <= 0) return G2V_ERR_INVALID;
  unsigned algo = flags & G2V_ALGO_MASK;
  if (algo == G2V_ALGO_SIMT) return (int)G2V_ALGO_SIMT;
  if (tc_supported(K, D)) return (int)G2V_ALGO_TC;
  return algo == G2V_ALGO_TC ? G2V_ERR_UNSUPPORTED : (int)G2V_ALGO_SIMT;
}

size_t g2v_workspace_bytes(int64_t N, int K, int D, int z_dtype, unsigned flags) {
  if (bad_shape(N, K, D)) return 0;
  size_t b = 256;  // stats
  unsigned algo = flags & G2V_ALGO_MASK;
  bool tc = (algo == G2V_ALGO_TC) || (algo == G2V_ALGO_AUTO && tc_supported(K, D));
  if (tc) b += align_up(tc_workspace_bytes(N, K, D, z_dtype), 256);
  else b += align_up((size_t)N * 4, 256);   // list of rows for the fp64 re-rank
  return b;
}

int g2v_vq_search(const void* z, int z_dtype, const float* E, const void* cb, int64_t N, int K, int D,
                  int32_t* idx, int64_t* search_stats, void* ws, size_t ws_bytes, unsigned flags,
                  void* stream) {
  if (bad_shape(N, K, D) || !E || !cb || (N > 0 && (!z || !idx))) return G2V_ERR_INVALID;
  if (z_dtype != G2V_F32 && z_dtype != G2V_BF16 && z_dtype != G2V_F16) return G2V_ERR_DTYPE;
  if (N == 0) return G2V_OK;
  int rc = check_arch();
  if (rc) return rc;
  if (ws_bytes < g2v_workspace_bytes(N, K, D, z_dtype, flags) || !ws) return G2V_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(ws) & 255) return G2V_ERR_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  auto* stats = reinterpret_cast<unsigned long long*>(search_stats);
  char* p = reinterpret_cast<char*>(ws) + 256;

  unsigned algo = flags & G2V_ALGO_MASK;
  if (algo == G2V_ALGO_TC && !tc_supported(K, D)) {
    set_error_detail("tensor-core search does not cover K=%d D=%d", K, D);
    return G2V_ERR_UNSUPPORTED;
  }
  bool tc = (algo == G2V_ALGO_TC) || (algo == G2V_ALGO_AUTO && tc_supported(K, D));
  if (tc) return launch_search_tc(z, z_dtype, E, cb, N, K, D, idx, stats, p, ws_bytes - 256, flags, st);
  return launch_search_simt(z, z_dtype, E, cb, N, K, D, reinterpret_cast<int32_t*>(p),
                            reinterpret_cast<int32_t*>(ws), idx, stats, st);
}

int g2v_vq_search_wide(const void* z, int z_dtype, int Dz, const float* E, const void* cb, int64_t N, int K, int D,
                       int32_t* idx, int64_t* search_stats, void* ws, size_t ws_bytes, unsigned flags, void* stream) {
  if (bad_shape(N, K, D) || Dz <= 0 || Dz > D || !E || !cb || (N > 0 && (!z || !idx))) return G2V_ERR_INVALID;
  if (Dz == D) return g2v_vq_search(z, z_dtype, E, cb, N, K, D, idx, search_stats, ws, ws_bytes, flags, stream);
  if (z_dtype != G2V_F32 && z_dtype != G2V_BF16 && z_dtype != G2V_F16) return G2V_ERR_DTYPE;
  if (N == 0) return G2V_OK;
  int rc = check_arch();
  if (rc) return rc;
  if (!tc_supported(K, D)) return G2V_ERR_UNSUPPORTED;
  if (ws_bytes < g2v_workspace_bytes(N, K, D, z_dtype, G2V_ALGO_TC) || !ws) return G2V_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(ws) & 255) return G2V_ERR_ALIGN;
  return launch_search_tc(z, z_dtype, E, cb, N, K, D, idx, reinterpret_cast<unsigned long long*>(search_stats),
                          reinterpret_cast<char*>(ws) + 256, ws_bytes - 256, flags, (cudaStream_t)stream, Dz);
}

int g2v_vq_apply(const float* x, const float* zs, const float* E, const int32_t* idx, int64_t N, int K, int D,
                 float* out, double* sse, int32_t* counts, float* dwr, int dwr_replicas, void* stream) {
  if (bad_shape(N, K, D) || !E || (N > 0 && (!x || !idx)) || (dwr && dwr_replicas < 1)) return G2V_ERR_INVALID;
  if (N == 0) return G2V_OK;
  int rc = check_arch();
  if (rc) return rc;
  return launch_apply(x, zs, E, idx, N, K, D, out, sse, counts, dwr, dwr_replicas, (cudaStream_t)stream);
}

size_t g2v_apply_workspace_bytes(int64_t N, int K) {
  if (N < 0 || K <= 0) return 0;
  return apply_ws_bytes(N, K);
}

int g2v_vq_apply_ws(const float* x, const float* zs, const float* E, const int32_t* idx, int64_t N, int K, int D,
                    float* out, double* sse, int32_t* counts, float* dwr, int dwr_replicas, void* ws, size_t ws_bytes,
                    void* stream) {
  if (bad_shape(N, K, D) || !E || (N > 0 && (!x || !idx)) || (dwr && dwr_replicas < 1)) return G2V_ERR_INVALID;
  if (N == 0) return G2V_OK;
  int rc = check_arch();
  if (rc) return rc;
  return launch_apply(x, zs, E, idx, N, K, D, out, sse, counts, dwr, dwr_replicas, (cudaStream_t)stream, ws, ws_bytes);
}

int g2v_fold_projection(const float* E, const float* W, const float* b, int K, int D, int ld_out, float* out, double* g,
                        void* stream) {
  if (K <= 0 || D <= 0 || ld_out < D || !E || !W || !b || !out || !g || (size_t)D * 4 > 48 * 1024) return G2V_ERR_INVALID;
  return launch_fold_rows(E, W, b, K, D, ld_out, out, g, (cudaStream_t)stream);
}

int g2v_pad_rows(const float* x, int64_t N, int D, int Dp, float* out, void* stream) {
  if (N < 0 || D <= 0 || Dp < D || (D % 4) || (Dp % 4) || (N > 0 && (!x || !out))) return G2V_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) return G2V_ERR_ALIGN;
  if (N == 0) return G2V_OK;
  return launch_pad_rows(x, N, D, Dp, out, (cudaStream_t)stream);
}

int g2v_vq_stats_deterministic(const float* x, const float* zs, const float* E, const int32_t* order,
                                const int64_t* seg, const int64_t* chunk_off, int64_t max_chunks, int K, int D,
                                double* partial, float* dwr, double* sse_code, double* sse, void* stream) {
  if (K <= 0 || D <= 0 || max_chunks < 0 || !x || !E || !order || !seg || !chunk_off || !partial || !dwr || !sse_code || !sse)
    return G2V_ERR_INVALID;
  return launch_stats_deterministic(x, zs, E, order, reinterpret_cast<const long long*>(seg),
                                    reinterpret_cast<const long long*>(chunk_off), max_chunks, K, D, partial, dwr, sse_code,
                                    sse, (cudaStream_t)stream);
}

int g2v_vq_stats_pack(const int32_t* counts, const double* sse, const float* dwr, int dwr_replicas, int64_t N,
                      int K, int D, float* packed, void* stream) {
  if (K <= 0 || D <= 0 || N < 0 || !packed || (dwr && dwr_replicas < 1)) return G2V_ERR_INVALID;
  return launch_stats_pack(counts, sse, dwr, dwr_replicas, N, K, D, packed, (cudaStream_t)stream);
}

int g2v_vq_stats_finalize(const float* packed, int K, int D, float coef_codebook, float coef_commit,
                          float* loss, float* perplexity, void* stream) {
  if (K <= 0 || D <= 0 || !packed) return G2V_ERR_INVALID;
  return launch_stats_finalize(packed, K, D, coef_codebook, coef_commit, loss, perplexity,
                               (cudaStream_t)stream);
}

int g2v_vq_step_finalize(int32_t* counts, double* sse, float* dwr, int dwr_replicas, int64_t rows_local,
                         float* packed, int K, int D, float coef_codebook, float coef_commit, float* loss,
                         float* perplexity, int update, const float* cs_in, float* cs_out, const float* ema_w_in,
                         float* ema_w_out, const float* E_old, float* E_new, float* E_prev, float decay, float eps,
                         double* shift2, void* cb, size_t cb_bytes, void* stream) {
  if (K <= 0 || D <= 0 || rows_local < 0 || (dwr && dwr_replicas < 1)) return G2V_ERR_INVALID;
  const int do_pack = (counts || sse || dwr) ? 1 : 0;
  if ((do_pack || loss || perplexity || update != G2V_UPDATE_NONE) && !packed) return G2V_ERR_INVALID;
  if (update == G2V_UPDATE_EMA) {
    if (!cs_in || !cs_out || !ema_w_in || !ema_w_out || !E_old || !E_new) return G2V_ERR_INVALID;
    if (E_prev && (E_prev == E_old || E_prev == E_new)) return G2V_ERR_INVALID;
  } else if (update == G2V_UPDATE_KMEANS) {
    if (!E_old || !E_new) return G2V_ERR_INVALID;
  } else if (update != G2V_UPDATE_NONE) {
    return G2V_ERR_INVALID;
  }
  if (cb) {
    if (cb_bytes < cb_total_bytes(K, D)) return G2V_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(cb) & 255) return G2V_ERR_ALIGN;
    if (update == G2V_UPDATE_NONE && !E_old) return G2V_ERR_INVALID;
  }
  int rc = check_arch();
  if (rc) return rc;
  return launch_step_finalize(counts, sse, dwr, dwr_replicas, do_pack, rows_local, packed, K, D, coef_codebook,
                              coef_commit, loss, perplexity, update, cs_in, cs_out, ema_w_in, ema_w_out, E_old, E_new,
                              update == G2V_UPDATE_EMA ? E_prev : nullptr, decay, eps, shift2, cb, E_old,
                              (cudaStream_t)stream);
}

int g2v_vq_ema_update(const float* cs_in, float* cs_out, const float* ema_w_in, float* ema_w_out,
                      const float* E_old, float* E_new, const float* packed, float decay, float eps, int K, int D,
                      void* cb, size_t cb_bytes, void* stream) {
  return g2v_vq_step_finalize(nullptr, nullptr, nullptr, 0, 0, const_cast<float*>(packed), K, D, 0.f, 0.f, nullptr,
                              nullptr, G2V_UPDATE_EMA, cs_in, cs_out, ema_w_in, ema_w_out, E_old, E_new, nullptr, decay,
                              eps, nullptr, cb, cb_bytes, stream);
}

int g2v_vq_backward(const float* x, const float* E, const int32_t* idx, const float* g_out,
                    const float* g_loss, float coef_x, int64_t N, int K, int D, float* g_x, void* stream) {
  if (bad_shape(N, K, D) || !E || !g_loss || (N > 0 && (!x || !idx || !g_x))) return G2V_ERR_INVALID;
  if (N == 0) return G2V_OK;
  return launch_backward(x, E, idx, g_out, g_loss, coef_x, N, K, D, g_x, (cudaStream_t)stream);
}

int g2v_vq_grad_codebook(const float* packed_dwr, const float* g_loss, float coef_e, int K, int D,
                         float* g_E, void* stream) {
  if (K <= 0 || D <= 0 || !packed_dwr || !g_loss || !g_E) return G2V_ERR_INVALID;
  return launch_grad_codebook(packed_dwr, g_loss, coef_e, K, D, g_E, (cudaStream_t)stream);
}

int g2v_kmeans_update(const float* E_old, const float* packed, int K, int D, float* E_new, double* shift2,
                      void* cb, size_t cb_bytes, void* stream) {
  return g2v_vq_step_finalize(nullptr, nullptr, nullptr, 0, 0, const_cast<float*>(packed), K, D, 0.f, 0.f, nullptr,
                              nullptr, G2V_UPDATE_KMEANS, nullptr, nullptr, nullptr, nullptr, E_old, E_new, nullptr, 0.f,
                              0.f, shift2, cb, cb_bytes, stream);
}

size_t g2v_gemm_workspace_bytes(int64_t M, int N, int64_t K, unsigned flags) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  return gemm_workspace_bytes(M, N, K, (flags & G2V_GEMM_FP16) ? 1 : 0);
}

int g2v_gemm_f32(const float* A, int64_t lda, int transA, const float* B, int64_t ldb, int transB, int64_t M, int N,
                 int64_t K, const float* bias, float* C, int64_t ldc, float alpha, unsigned flags, void* ws,
                 size_t ws_bytes, void* stream) {
  const int single = (flags & G2V_GEMM_FP16) ? 1 : 0, accumulate = (flags & G2V_GEMM_ACCUMULATE) ? 1 : 0;
  if (M <= 0 || N <= 0 || K <= 0 || !A || !B || !C || M > 0x7fffffffLL || K > 0x7fffffffLL) return G2V_ERR_INVALID;
  if (lda < (transA ? M : K) || ldb < (transB ? (int64_t)N : K) || ldc < N) return G2V_ERR_INVALID;
  if (!ws || ws_bytes < gemm_workspace_bytes(M, N, K, single)) return G2V_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(ws) & 255) return G2V_ERR_ALIGN;
  int rc = check_arch();
  if (rc) return rc;
  return launch_gemm_f32(A, lda, transA, B, ldb, transB, M, N, K, bias, C, ldc, alpha, accumulate, single, ws,
                         (cudaStream_t)stream);
}

int g2v_soft_assign(const float* m, float* dot_d, const float* lv, const float* e2, int64_t N, int K, int D, float* p,
                    float* colsum, void* stream) {
  if (bad_shape(N, K, D) || !e2 || !colsum || (N > 0 && (!m || !dot_d || !lv || !p))) return G2V_ERR_INVALID;
  if ((size_t)K * 8 > 48 * 1024) return G2V_ERR_UNSUPPORTED;        // column sums live in (default-size) shared memory
  if (N == 0) return G2V_OK;
  return launch_soft_assign(m, dot_d, lv, e2, N, K, D, p, colsum, (cudaStream_t)stream);
}

int g2v_soft_tail(const float* x, const float* q, int64_t N, int K, int D, float beta, float* out, double* sse,
                  const float* colsum, float* loss, float* perplexity, void* stream) {
  if (bad_shape(N, K, D) || !sse || !colsum || !loss || !perplexity || (N > 0 && (!x || !q || !out))) return G2V_ERR_INVALID;
  return launch_soft_tail(x, q, N, K, D, beta, out, sse, colsum, loss, perplexity, (cudaStream_t)stream);
}

int g2v_soft_backward(const float* p, const float* dp, const float* d, const float* lv, int64_t N, int K, float* gd,
                      float* glv, float* rowsum_gd, float* col_gd, float* col_glv, void* stream) {
  if (N < 0 || K <= 0 || !col_gd || !col_glv || (N > 0 && (!p || !dp || !d || !lv || !gd || !glv || !rowsum_gd)))
    return G2V_ERR_INVALID;
  if ((size_t)K * 8 > 48 * 1024) return G2V_ERR_UNSUPPORTED;
  if (N == 0) return G2V_OK;
  return launch_soft_bwd(p, dp, d, lv, N, K, gd, glv, rowsum_gd, col_gd, col_glv, (cudaStream_t)stream);
}

int g2v_soft_gx(const float* gmw, const float* x, const float* q, const float* g_out, const float* c, const float* a,
                int64_t n, float* gx, void* stream) {
  if (n < 0 || !c || !a || (n > 0 && (!gmw || !x || !q || !gx))) return G2V_ERR_INVALID;
  if (n == 0) return G2V_OK;
  return launch_soft_gx(gmw, x, q, g_out, c, a, n, gx, (cudaStream_t)stream);
}

size_t g2v_exact_workspace_bytes(int K) { return K > 0 ? align_up((size_t)K * sizeof(double), 256) : 0; }

int g2v_vq_search_exact(const void* z, int z_dtype, const float* E, int64_t N, int K, int D, int32_t* idx,
                        void* ws, size_t ws_bytes, void* stream) {
  if (bad_shape(N, K, D) || !E || (N > 0 && (!z || !idx))) return G2V_ERR_INVALID;
  if (z_dtype != G2V_F32 && z_dtype != G2V_BF16 && z_dtype != G2V_F16) return G2V_ERR_DTYPE;
  if (N == 0) return G2V_OK;
  if (!ws || ws_bytes < g2v_exact_workspace_bytes(K)) return G2V_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(ws) & 7) return G2V_ERR_ALIGN;
  int rc = check_arch();
  if (rc) return rc;
  return launch_search_exact64(z, z_dtype, E, N, K, D, idx, ws, (cudaStream_t)stream);
}

int g2v_onehot(const int32_t* idx, int64_t N, int K, float* enc, void* stream) {
  if (N < 0 || K <= 0 || (N > 0 && (!idx || !enc))) return G2V_ERR_INVALID;
  if (N == 0) return G2V_OK;
  return launch_onehot(idx, N, K, enc, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// host-buffer tokeniser: 3-slot ring, H2D / search / D2H overlapped on three streams
// ---------------------------------------------------------------------------------------------
static constexpr int kSlots = 3;

size_t g2v_tokenize_host_bytes(int64_t chunk_rows, int K, int D, int z_dtype, unsigned flags) {
  if (bad_shape(chunk_rows, K, D) || chunk_rows == 0) return 0;
  size_t per = align_up((size_t)chunk_rows * D * elem_size(z_dtype), 256) + align_up((size_t)chunk_rows * 4, 256) +
               align_up(g2v_workspace_bytes(chunk_rows, K, D, z_dtype, flags), 256) + 256;
  return per * kSlots;
}

int g2v_tokenize_host(const void* z_host, int z_dtype, int64_t N, const float* E, const void* cb, int K, int D,
                      int32_t* idx_host, int64_t chunk_rows, int64_t* search_stats_host, void* ws,
                      size_t ws_bytes, unsigned flags) {
  if (bad_shape(N, K, D) || chunk_rows <= 0 || !E || !cb || (N > 0 && (!z_host || !idx_host))) return G2V_ERR_INVALID;
  if (z_dtype != G2V_F32 && z_dtype != G2V_BF16 && z_dtype != G2V_F16) return G2V_ERR_DTYPE;
  if (!ws || ws_bytes < g2v_tokenize_host_bytes(chunk_rows, K, D, z_dtype, flags)) return G2V_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(ws) & 255) return G2V_ERR_ALIGN;
  int rc = check_arch();
  if (rc) return rc;
  const size_t esz = elem_size(z_dtype);
  const size_t zb = align_up((size_t)chunk_rows * D * esz, 256), ib = align_up((size_t)chunk_rows * 4, 256);
  const size_t wb = align_up(g2v_workspace_bytes(chunk_rows, K, D, z_dtype, flags), 256);
  const size_t per = zb + ib + wb + 256;

  cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[kSlots] = {}, ev_run[kSlots] = {}, ev_out[kSlots] = {};
  rc = G2V_OK;
  auto cleanup = [&]() {
    for (int i = 0; i < kSlots; ++i) {
      if (ev_in[i]) cudaEventDestroy(ev_in[i]);
      if (ev_run[i]) cudaEventDestroy(ev_run[i]);
      if (ev_out[i]) cudaEventDestroy(ev_out[i]);
    }
    if (s_in) cudaStreamDestroy(s_in);
    if (s_run) cudaStreamDestroy(s_run);
    if (s_out) cudaStreamDestroy(s_out);
  };
#define TOK_CHECK(expr)                            \
  do {                                             \
    cudaError_t _e = (expr);                       \
    if (_e != cudaSuccess) {                       \
      rc = cuda_fail(_e, #expr);                   \
      cudaDeviceSynchronize();                     \
      cleanup();                                   \
      return rc;                                   \
    }                                              \
  } while (0)
  TOK_CHECK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
  TOK_CHECK(cudaStreamCreateWithFlags(&s_run, cudaStreamNonBlocking));
  TOK_CHECK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
  for (int i = 0; i < kSlots; ++i) {
    TOK_CHECK(cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming));
    TOK_CHECK(cudaEventCreateWithFlags(&ev_run[i], cudaEventDisableTiming));
    TOK_CHECK(cudaEventCreateWithFlags(&ev_out[i], cudaEventDisableTiming));
  }
  char* base = reinterpret_cast<char*>(ws);
  int64_t* dstats = reinterpret_cast<int64_t*>(base + per * kSlots - 256);  // last slot's tail: shared stats
  TOK_CHECK(cudaMemsetAsync(dstats, 0, 64, s_run));
  const char* zh = reinterpret_cast<const char*>(z_host);
  int64_t chunk = 0;
  for (int64_t r0 = 0; r0 < N; r0 += chunk_rows, ++chunk) {
    const int slot = (int)(chunk % kSlots);
    const int64_t n = (N - r0 < chunk_rows) ? (N - r0) : chunk_rows;
    char* zd = base + per * slot;
    int32_t* id = reinterpret_cast<int32_t*>(zd + zb);
    void* wsd = zd + zb + ib;
    if (chunk >= kSlots) {
      // slot reuse: input buffer free once its search ran; idx buffer free once copied out
      TOK_CHECK(cudaStreamWaitEvent(s_in, ev_run[slot], 0));
      TOK_CHECK(cudaStreamWaitEvent(s_run, ev_out[slot], 0));
    }
    TOK_CHECK(cudaMemcpyAsync(zd, zh + (size_t)r0 * D * esz, (size_t)n * D * esz, cudaMemcpyHostToDevice, s_in));
    TOK_CHECK(cudaEventRecord(ev_in[slot], s_in));
    TOK_CHECK(cudaStreamWaitEvent(s_run, ev_in[slot], 0));
    int r = g2v_vq_search(zd, z_dtype, E, cb, n, K, D, id, dstats, wsd, wb, flags, s_run);
    if (r) { cudaDeviceSynchronize(); cleanup(); return r; }
    TOK_CHECK(cudaEventRecord(ev_run[slot], s_run));
    TOK_CHECK(cudaStreamWaitEvent(s_out, ev_run[slot], 0));
    TOK_CHECK(cudaMemcpyAsync(idx_host + r0, id, (size_t)n * 4, cudaMemcpyDeviceToHost, s_out));
    TOK_CHECK(cudaEventRecord(ev_out[slot], s_out));
  }
  if (search_stats_host) {
    TOK_CHECK(cudaStreamWaitEvent(s_out, ev_run[(chunk + kSlots - 1) % kSlots], 0));
    TOK_CHECK(cudaMemcpyAsync(search_stats_host, dstats, 64, cudaMemcpyDeviceToHost, s_out));
  }
  TOK_CHECK(cudaStreamSynchronize(s_out));
  TOK_CHECK(cudaStreamSynchronize(s_run));
  TOK_CHECK(cudaStreamSynchronize(s_in));
#undef TOK_CHECK
  cleanup();
  return G2V_OK;
}

}  // extern "C"
