/* g2v_vq.h -- C ABI of the B200 (sm_100a) vector-quantizer hot path.
 *
 * This is the drop-in boundary for the quantizer layer of pjyazdian/Gesture2Vec.
 * The reference has no FFI: its quantizers are eager PyTorch modules.  Each entry
 * point below names the reference statements (file:line under
 * /root/reference/scripts/model/) whose work it replaces; the Python modules in
 * gesture2vec_b200/quantizers.py bind these through ctypes and keep the
 * reference's nn.Module interface.  See INTEGRATION.md for the binding.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the
 *     parameter name ends in _host;
 *   - all work is enqueued on the caller's stream (cudaStream_t passed as void*);
 *     nothing synchronises the device except the *_host entry point;
 *   - the library allocates no persistent device memory: scratch comes from the
 *     caller through (ws, ws_bytes), sized by g2v_workspace_bytes();
 *   - return value: 0 = ok, negative = error (g2v_strerror); never throws/exits;
 *   - row-major fp32 everywhere unless a dtype code says otherwise; rows of z/x/out
 *     are contiguous with stride D; the codebook E is [K, D].
 */
#ifndef G2V_VQ_H_
#define G2V_VQ_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define G2V_VERSION 200

/* error codes */
#define G2V_OK 0
#define G2V_ERR_INVALID (-1)    /* null pointer / non-positive size / bad flag */
#define G2V_ERR_ALIGN (-2)      /* pointer not aligned as required             */
#define G2V_ERR_DTYPE (-3)      /* unsupported dtype code                      */
#define G2V_ERR_WORKSPACE (-4)  /* workspace / codebook-aux buffer too small   */
#define G2V_ERR_CUDA (-5)       /* a CUDA call failed (g2v_last_error_detail)  */
#define G2V_ERR_ARCH (-6)       /* device is not sm_100                        */
#define G2V_ERR_UNSUPPORTED (-7)/* shape outside what the selected path covers */

/* dtype codes for the search rows */
#define G2V_F32 0
#define G2V_BF16 1
#define G2V_F16 2

/* flags for g2v_vq_search */
#define G2V_ALGO_AUTO 0u        /* tcgen05 path when the shape allows, else SIMT */
#define G2V_ALGO_SIMT 1u        /* fp32 CUDA-core path (any K, D)                */
#define G2V_ALGO_TC 2u          /* tcgen05 tensor-core path, error if unsupported */
#define G2V_ALGO_MASK 3u
#define G2V_NO_RECHECK 4u       /* keep the fast-pass winner (bf16/fp16 "fast" variant) */
#define G2V_NO_REFINE 8u        /* test / benchmark aid: whole-row re-ranks go straight to fp64 instead of through the
                                 * tensor-core refine pass (results are identical) */
#define G2V_LIST_ALL_ROWS 16u    /* test aid: the tensor-core pass lists EVERY row as a whole-row re-rank, so the refine pass
                                 * (or, with G2V_NO_REFINE, the fp64 path) decides all of them -- the audit of its error bound */
/* test / benchmark aid: pin the tensor-core sweep kernel (results are identical; a variant that does not cover
 * the shape falls back to the automatic choice) */
#define G2V_TC_VARIANT_MASK (7u << 8)
#define G2V_TC_VARIANT_AUTO (0u << 8)
#define G2V_TC_VARIANT_TMEM (1u << 8)   /* tc_tmem_kernel: rows -> fp16 A operand in tensor memory, CTA pairs */
#define G2V_TC_VARIANT_FUSED (2u << 8)  /* tc_search_kernel<1,true>: fp32 rows converted in-kernel, single CTA */
#define G2V_TC_VARIANT_PREP (3u << 8)   /* row_prep_kernel + tc_search_kernel<.,false>: fp16 rows in shared memory */

/* slots of the optional int64 search_stats[8] output */
#define G2V_STAT_ROWS 0         /* rows searched                                      */
#define G2V_STAT_PAIR_RECHECK 1 /* rows whose top-2 were re-ranked exactly (fp64)     */
#define G2V_STAT_FULL_RECHECK 2 /* rows whose whole distance row was recomputed (fp64) */
#define G2V_STAT_FALLBACK_ROWS 3/* rows the tensor-core pass handed to the fp32 path   */
#define G2V_STAT_REFINE_ROWS 4  /* listed whole rows that took the fp32-accurate tensor-core refine pass */
#define G2V_STAT_REFINE_EXACT 5 /* fp64 code evaluations the refine pass needed for them */

int g2v_version(void);
const char* g2v_strerror(int code);
/* text of the last failure on the calling thread (CUDA error string etc.) */
const char* g2v_last_error_detail(void);
/* kernels this library has launched so far in this process (all threads); a benchmark reports the difference
 * around its timed region as its launch count */
unsigned long long g2v_launch_count(void);

/* Codebook auxiliary data (||e||^2 in fp32 rounded from fp64, norm maxima for the
 * error bounds, and the fp16 operand copy the tensor-core path streams by TMA).
 * Must be re-prepared whenever E changes.  Replaces torch.sum(weight**2, dim=1)
 * (DAE_model.py:322,425; Autoencoder_VQVAE_model.py:1134,1236). */
size_t g2v_codebook_bytes(int K, int D);
int g2v_codebook_prepare(const float* E, int K, int D, void* cb, size_t cb_bytes, void* stream);

/* Which search path the flags select for this shape: G2V_ALGO_SIMT or G2V_ALGO_TC
 * (negative error if G2V_ALGO_TC is forced on a shape it does not cover). */
int g2v_search_path(int K, int D, unsigned flags);

/* Measurement aid: the NEXT g2v_vq_search on this host thread records `ev_start` / `ev_stop`
 * (cudaEvent_t) on its stream immediately around its dominant kernel (the tcgen05 sweep, or the
 * fp32 sweep on the SIMT path), so a benchmark can time that kernel alone inside a timed region. */
int g2v_profile_next_search(void* ev_start, void* ev_stop);

/* Scratch needed by g2v_vq_search for N rows. */
size_t g2v_workspace_bytes(int64_t N, int K, int D, int z_dtype, unsigned flags);

/* Nearest-code search: idx[n] = argmin_k ||z_n - e_k||^2, first index on exact ties.
 * Replaces the distance matrix + argmin (DAE_model.py:320-327, 423-434;
 * Autoencoder_VQVAE_model.py:1132-1142, 1234-1244, 1760-1767) without materialising
 * the [N,K] distances.  Every path is a fast pass with a known error bound followed
 * by an exact (fp64) re-rank of the rows whose top candidates are closer than that
 * bound, so the result is the exact-arithmetic argmin.
 *   z          [N,D] rows in z_dtype
 *   E, cb      fp32 codebook [K,D] and its prepared aux buffer
 *   idx        out, int32 [N]
 *   search_stats  optional int64[8], ACCUMULATED (caller zeroes)                      */
int g2v_vq_search(const void* z, int z_dtype, const float* E, const void* cb,
                  int64_t N, int K, int D, int32_t* idx, int64_t* search_stats,
                  void* ws, size_t ws_bytes, unsigned flags, void* stream);

/* g2v_vq_search for rows that are NARROWER than the codebook: z is [N, Dz] with Dz <= D and counts as [z | 0].
 * This is how a codebook with a linear projection folded in (quantizers.py VQVAE_VQ_Payam_EMA._fold: pre_linear of
 * Autoencoder_VQVAE_model.py:1230 becomes D+4 code columns) is searched with the RAW rows: the TMA loads of the
 * tensor-core sweep zero-fill the missing operand columns, so no widened copy of the rows is made.  Covered by the
 * tcgen05 sweep that reads the rows through TMA (Dz % 4 == 0 for fp32, % 8 for 16-bit rows, 16-byte aligned, N > 128);
 * G2V_ERR_UNSUPPORTED otherwise (the caller widens the rows with g2v_pad_rows and calls g2v_vq_search). */
int g2v_vq_search_wide(const void* z, int z_dtype, int Dz, const float* E, const void* cb, int64_t N, int K, int D,
                       int32_t* idx, int64_t* search_stats, void* ws, size_t ws_bytes, unsigned flags, void* stream);

/* Gather + straight-through value + loss/EMA statistics in one pass over the rows.
 * Replaces the one-hot GEMM gather, both mse_loss reductions, the STE add, the
 * column sums of the one-hot and encodings^T @ flat_input
 * (DAE_model.py:334-345, 445-464; Autoencoder_VQVAE_model.py:1151-1170, 1256-1275).
 *   x       [N,D] fp32 rows the loss/STE are taken against (raw inputs)
 *   zs      [N,D] fp32 rows the EMA sums are taken over (NULL -> x; differs only for
 *           Autoencoder_VQVAE_model.VQ_Payam_EMA which searches on pre_linear(x), :1230)
 *   out     optional [N,D]: x + (E[idx] - x)
 *   sse     optional double[1], accumulated: sum (E[idx]-x)^2
 *   counts  optional int32[K], accumulated: rows per code
 *   dwr     optional fp32 [dwr_replicas, K, D], accumulated: sum_{n: idx=k} (zs_n - e_k)  (the
 *           EMA sum encodings^T @ flat_input minus counts*E; residual form keeps grad_E accurate).
 *           dwr_replicas >= 1 independent copies spread the atomic traffic of hot codes (thread
 *           blocks round-robin over them); g2v_vq_stats_pack sums the copies. */
int g2v_vq_apply(const float* x, const float* zs, const float* E, const int32_t* idx,
                 int64_t N, int K, int D, float* out, double* sse, int32_t* counts,
                 float* dwr, int dwr_replicas, void* stream);

/* The same with scratch for the bulk path (N >= 16384 and residual sums wanted): the rows are first grouped by code
 * on the device (one cooperative launch), so that the row pass sums whole runs of equal codes in registers and issues
 * a handful of vector reductions per code instead of one per run of a 32-row batch -- the reductions, not DRAM, bound
 * the unsorted pass.  Results are the same up to the order of the fp32 additions.  ws: g2v_apply_workspace_bytes(N, K)
 * bytes (0 = this shape takes the unsorted pass; ws may then be NULL), 16-byte aligned. */
size_t g2v_apply_workspace_bytes(int64_t N, int K);
int g2v_vq_apply_ws(const float* x, const float* zs, const float* E, const int32_t* idx,
                    int64_t N, int K, int D, float* out, double* sse, int32_t* counts,
                    float* dwr, int dwr_replicas, void* ws, size_t ws_bytes, void* stream);

/* Fold the linear layer in front of a search into the codebook (pre_linear, Autoencoder_VQVAE_model.py:1230):
 * argmin_k |W z + b - e_k|^2 = argmin_k (c_k - 2 z.(W^T e_k)), c_k = |e_k|^2 - 2 b.e_k.  Writes out[k, 0..D) =
 * W^T e_k (fp64 accumulation, one rounding; W is nn.Linear's [D_out, D_in] weight read as rows i, columns j) and
 * g[k] = c_k - |out[k, 0..D)|^2 in fp64; the caller completes the folded code with the coordinate sqrt(g[k] + C). */
int g2v_fold_projection(const float* E, const float* W, const float* b, int K, int D, int ld_out, float* out, double* g,
                        void* stream);

/* out [N, Dp] = [x | 0]: the raw rows widened to the width of a folded codebook (quantizers.py
 * VQVAE_VQ_Payam_EMA._fold: pre_linear, Autoencoder_VQVAE_model.py:1230, folded into [K, D+4] codes).  D, Dp
 * multiples of 4, 16-byte aligned buffers. */
int g2v_pad_rows(const float* x, int64_t N, int D, int Dp, float* out, void* stream);

/* Reproducible statistics (SURVEY.md 7 "Atomics determinism"): the residual sums dwr [K,D] and the squared error
 * of g2v_vq_apply WITHOUT atomics -- rows sorted by code (order: stable, ties by row index; seg [K+1]: first
 * position of each code; chunk_off [K+1]: prefix sum of ceil(count_k / G2V_DET_CHUNK)), every chunk of
 * consecutive positions summed row after row in fp64, every code's chunks added in ascending order, one rounding to
 * fp32.  Bit-identical from run to run; ~1e-7 relative to the exact sums.  partial: double [max_chunks, D+1]
 * scratch, max_chunks >= chunk_off[K] (N / G2V_DET_CHUNK + K always suffices); sse_code: double [K] scratch. */
#define G2V_DET_CHUNK 128
int g2v_vq_stats_deterministic(const float* x, const float* zs, const float* E, const int32_t* order,
                                const int64_t* seg, const int64_t* chunk_off, int64_t max_chunks, int K, int D,
                                double* partial, float* dwr, double* sse_code, double* sse, void* stream);

/* Pack the per-step statistics into ONE fp32 buffer for the data-parallel all-reduce:
 * packed = [ dwr (K*D) | counts (K) | sse (1) | rows (1) ], K*D+K+2 floats.  The dwr_replicas
 * copies written by g2v_vq_apply are summed into packed[0 .. K*D) (dwr may be NULL: zeros). */
int g2v_vq_stats_pack(const int32_t* counts, const double* sse, const float* dwr, int dwr_replicas,
                      int64_t N, int K, int D, float* packed, void* stream);

/* mse = sse/(rows*D); loss = coef_codebook*mse + coef_commit*mse; perplexity =
 * exp(-sum p log(p+1e-10)), p = counts/rows -- from a (possibly all-reduced) packed buffer.
 * Replaces DAE_model.py:340-347, 474-481.  (coef_codebook, coef_commit) = (1, beta) for
 * VQ_Payam and (0, beta) for VQ_Payam_EMA.  loss / perplexity are device scalars. */
int g2v_vq_stats_finalize(const float* packed, int K, int D, float coef_codebook, float coef_commit,
                          float* loss, float* perplexity, void* stream);

/* EMA codebook update from a packed statistics buffer.  Replaces DAE_model.py:451-471 /
 * Autoencoder_VQVAE_model.py:1262-1282, 1777-1797:
 *   cs <- cs*decay + (1-decay)*counts ; n = sum cs ; cs <- (cs+eps)/(n+K*eps)*n
 *   ema_w <- ema_w*decay + (1-decay)*dw ; E <- ema_w / cs[:,None]      (dw = dwr + counts*E)
 * The new state goes to cs_out / ema_w_out / E_new, like the fresh tensors the reference produces each step
 * (:1276-1282); each may alias its input (in-place state, e.g. for CUDA-graph replay).
 * If cb != NULL the aux buffer is re-prepared for E_new in the same launch. */
int g2v_vq_ema_update(const float* cs_in, float* cs_out, const float* ema_w_in, float* ema_w_out,
                      const float* E_old, float* E_new, const float* packed, float decay, float eps, int K, int D,
                      void* cb, size_t cb_bytes, void* stream);

/* Everything a step does after the row pass, in ONE (cooperative) launch: pack the accumulators of
 * g2v_vq_apply (= g2v_vq_stats_pack, and the accumulators are handed back ZEROED, ready for the next step),
 * loss / perplexity (= g2v_vq_stats_finalize), the codebook update selected by `update` (= g2v_vq_ema_update
 * or g2v_kmeans_update) and the aux buffer of the new codebook (= g2v_codebook_prepare).
 *   counts / sse / dwr   accumulators; all three NULL = `packed` already holds the statistics (e.g. it was packed
 *                        by g2v_vq_stats_pack and all-reduced across ranks)
 *   rows_local           rows of this call (written to packed[K*D+K+1] when packing)
 *   loss, perplexity     optional device scalars
 *   update               G2V_UPDATE_NONE: no codebook change (cs / ema_w / E_new unused); cb, if given, is prepared
 *                        for E_old.  G2V_UPDATE_EMA: as g2v_vq_ema_update; E_prev (optional, [K,D]) receives a
 *                        copy of E_old, which an in-place update (E_new == E_old) needs for the backward pass.
 *                        G2V_UPDATE_KMEANS: as g2v_kmeans_update (shift2 optional).  cb, if given, is prepared
 *                        for E_new in both. */
#define G2V_UPDATE_NONE 0
#define G2V_UPDATE_EMA 1
#define G2V_UPDATE_KMEANS 2
int g2v_vq_step_finalize(int32_t* counts, double* sse, float* dwr, int dwr_replicas, int64_t rows_local,
                         float* packed, int K, int D, float coef_codebook, float coef_commit, float* loss,
                         float* perplexity, int update, const float* cs_in, float* cs_out, const float* ema_w_in,
                         float* ema_w_out, const float* E_old, float* E_new, float* E_prev, float decay, float eps,
                         double* shift2, void* cb, size_t cb_bytes, void* stream);

/* Backward of the layer wrt the inputs (closed form of the autograd graph the reference
 * builds):  g_x = g_out + g_loss[0]*coef_x*(x - E[idx]),
 * coef_x = 2*beta/(N*D).  g_out may be NULL (no gradient through `quantized`). */
int g2v_vq_backward(const float* x, const float* E, const int32_t* idx, const float* g_out,
                    const float* g_loss, float coef_x, int64_t N, int K, int D, float* g_x,
                    void* stream);

/* VQ_Payam only: g_E[k] = -g_loss[0]*coef_e*dwr[k], coef_e = 2/(N*D)
 * (the gradient the reference gets through the one-hot GEMM, DAE_model.py:334-341). */
int g2v_vq_grad_codebook(const float* packed_dwr, const float* g_loss, float coef_e,
                         int K, int D, float* g_E, void* stream);

/* Lloyd M-step of k-means on the statistics of an assignment pass (g2v_vq_search + g2v_vq_apply with dwr +
 * g2v_vq_stats_pack, all-reduced across ranks if the rows are sharded): E_new[k] = E_old[k] + dwr[k]/counts[k],
 * a cluster without rows keeps its centre; *shift2 (optional, double, accumulated) += ||E_new - E_old||_F^2.
 * Replaces the centre update inside sklearn.cluster.KMeans(...).fit as the reference calls it on gesture
 * latents (Clustering.py:718-720, train_DAE.py:257-263; sklearn is a third-party dependency of the
 * reference, requirements.txt).  E_new == E_old is allowed; if cb != NULL the aux buffer is re-prepared. */
int g2v_kmeans_update(const float* E_old, const float* packed, int K, int D, float* E_new, double* shift2,
                      void* cb, size_t cb_bytes, void* stream);

/* encodings = one_hot(idx) as a dense fp32 [N,K] (DAE_model.py:328-331); the reference
 * returns it and callers argmax it (Clustering.py:156, lmdb_data_loader.py:1281). */
int g2v_onehot(const int32_t* idx, int64_t N, int K, float* enc, void* stream);

/* Dense projection on the tensor cores at fp32 accuracy:  C[M,N] (+)= alpha * op(A) * op(B)^T + bias[N].
 * Replaces the fp32 nn.Linear / torch.matmul calls next to the search: pre_linear of
 * Autoencoder_VQVAE_model.VQ_Payam_EMA (:1230) and VectorQuantizerEMA (:1755), and the products of the soft
 * quantizer VQ_Payam_GSSoft (:1390-1412) and of its backward pass.  Each fp32 operand is split into two fp16
 * terms and the three significant partial products are folded into ONE tcgen05 GEMM with a 3x longer reduction
 * (csrc/g2v_gemm.cu): ~2^-21 relative, like an fp32 SGEMM.
 *   A    fp32, [M, K] row-major with leading dimension lda; transA != 0: stored [K, M] (the reduction runs over rows)
 *   B    fp32, [N, K] row-major (nn.Linear's weight layout) with ldb; transB != 0: stored [K, N]
 *   C    fp32 [M, N] with ldc; bias optional ([N])
 *   flags  G2V_GEMM_ACCUMULATE: C += ... instead of C = ...;  G2V_GEMM_FP16: one fp16 term per operand (2^-11
 *          per product, unbiased): for reductions over ~10^6 rows (weight gradients), a third of the traffic
 * A reduction much longer than the output (weight gradients) is split over thread-block pairs and accumulated
 * with fp32 atomics.  ws: g2v_gemm_workspace_bytes(M, N, K, flags) bytes (the fp16 operand copies). */
#define G2V_GEMM_ACCUMULATE 1u
#define G2V_GEMM_FP16 2u
size_t g2v_gemm_workspace_bytes(int64_t M, int N, int64_t K, unsigned flags);
int g2v_gemm_f32(const float* A, int64_t lda, int transA, const float* B, int64_t ldb, int transB, int64_t M, int N,
                 int64_t K, const float* bias, float* C, int64_t ldc, float alpha, unsigned flags, void* ws,
                 size_t ws_bytes, void* stream);

/* Row kernels of the soft quantizer VQ_Payam_GSSoft (Autoencoder_VQVAE_model.py:1304-1433), the layer
 * Autoencoder_VQVAE.__init__ instantiates (:816-820).  Its dense products go through g2v_gemm_f32; these do the rest.
 *
 * g2v_soft_assign   soft_prob (:1349-1372) fused with the distance assembly (:1393-1397) and smooth (:1399):
 *                   dot_d [N,K] holds m E^T on entry and d = (|m|^2 + |e|^2) - 2 m.e on exit; p [N,K] receives
 *                   p~ / sum_k p~, p~ = exp(-(d / 400)(0.5 s)) / sqrt(s), s = 1 / exp(lv)^2; colsum [K] (caller
 *                   zeroes) accumulates sum_n p for the perplexity.
 * g2v_soft_tail     out = x + (q - x) (:1424); sse (double, caller zeroes) += sum (q - x)^2; then loss = mse +
 *                   beta mse and perplexity = exp(-sum avg log(avg + 1e-10)), avg = colsum / N (:1418-1426).
 * g2v_soft_backward closed-form backward of the soft assignment given dp = dL/dp [N,K]: gd = dL/dd, glv = dL/dlv
 *                   [N,K], rowsum_gd [N], and the column sums col_gd / col_glv [K] (caller zeroes).
 * g2v_soft_gx       gx = c[0] gmw + a[0] (x - q) + g_out over n elements (c, a device scalars; g_out optional). */
int g2v_soft_assign(const float* m, float* dot_d, const float* lv, const float* e2, int64_t N, int K, int D, float* p,
                    float* colsum, void* stream);
int g2v_soft_tail(const float* x, const float* q, int64_t N, int K, int D, float beta, float* out, double* sse,
                  const float* colsum, float* loss, float* perplexity, void* stream);
int g2v_soft_backward(const float* p, const float* dp, const float* d, const float* lv, int64_t N, int K, float* gd,
                      float* glv, float* rowsum_gd, float* col_gd, float* col_glv, void* stream);
int g2v_soft_gx(const float* gmw, const float* x, const float* q, const float* g_out, const float* c, const float* a,
                int64_t n, float* gx, void* stream);

/* Verification aid, not a product path: the exact-arithmetic nearest code of EVERY row, all products and
 * sums in fp64 (no error bounds, candidate lists or operand rounding shared with g2v_vq_search), first index
 * on exact ties -- what torch.argmin over fp64 distances of DAE_model.py:320-327 would return.  FP64-bound
 * (~10 s per million rows at K = D = 400).  ws: g2v_exact_workspace_bytes(K) bytes of scratch. */
size_t g2v_exact_workspace_bytes(int K);
int g2v_vq_search_exact(const void* z, int z_dtype, const float* E, int64_t N, int K, int D, int32_t* idx,
                        void* ws, size_t ws_bytes, void* stream);

/* End-to-end tokenisation with HOST buffers (the batched form of Clustering.py:102-166):
 * copies z_host -> device in chunks, searches, copies idx back; copies and kernels overlap
 * on internal streams; returns after everything completed.  z_host/idx_host should be
 * pinned for full PCIe rate.  ws must hold g2v_tokenize_host_bytes(). */
size_t g2v_tokenize_host_bytes(int64_t chunk_rows, int K, int D, int z_dtype, unsigned flags);
int g2v_tokenize_host(const void* z_host, int z_dtype, int64_t N, const float* E, const void* cb,
                      int K, int D, int32_t* idx_host, int64_t chunk_rows, int64_t* search_stats_host,
                      void* ws, size_t ws_bytes, unsigned flags);

#ifdef __cplusplus
}
#endif
#endif /* G2V_VQ_H_ */
