// fp32-accurate GEMM on the tcgen05 tensor cores (sm_100a):  C[M,N] (+)= alpha * A[M,K] * B[N,K]^T + bias[N]
//
// The quantizer path has dense projections next to the nearest-code search: pre_linear of
// Autoencoder_VQVAE_model.VQ_Payam_EMA (:1230) / VectorQuantizerEMA (:1755), and every product of the soft
// quantizer VQ_Payam_GSSoft (:1374-1433: mean_layer, logvar_layer, the distance contraction, p @ E, and their
// backward).  The reference runs them as fp32 SGEMMs; here they run on the tensor cores at fp32 accuracy by
// splitting each fp32 operand into two fp16 terms, x * s = hi + lo (s a power of two that brings the tensor's
// largest magnitude to [256, 512); hi = fp16(x s), lo = fp16(x s - hi): 22 significant bits), and
// concatenating the terms ALONG THE REDUCTION DIMENSION:
//
//     A' = [ hi_a | lo_a | hi_a ]      B' = [ hi_b | hi_b | lo_b ]      A' B'^T = hi hi + lo hi + hi lo
//
// so ONE plain fp16 GEMM with a three times longer K loop produces the product to ~2^-21 relative (the dropped
// lo lo term and the fp32 accumulation are below that).  Two kernels:
//   split_prep_kernel   fp32 [R, C] (optionally read transposed) -> the fp16 operand [rows, 3 Kp]
//   tc_gemm_kernel      CTA pairs (cta_group::2, UMMA M = 256, N = BN <= 256), both operands streamed by TMA
//                       through a 6-stage ring, two accumulator stages in tensor memory, four epilogue warps per
//                       CTA (scale, bias, fp32 stores; red.add for split-K), persistent over (m, n, k-split) tiles.
#include "g2v_tcgen05.cuh"

#include <math.h>

namespace g2v {
namespace {

constexpr int GM_TM = 128;            // rows per CTA (UMMA M = 256 per pair)
constexpr int GM_KC = 64;             // fp16 per K panel (128 bytes, SWIZZLE_128B)
constexpr int GM_STAGES = 6;
constexpr int GM_THREADS = 192;       // TMA warp, MMA warp, 4 epilogue warps
constexpr int GM_A_STAGE = GM_TM * GM_KC * 2;   // 16384

// ------------------------------------------------------------------------------------------
// operand preparation
// ------------------------------------------------------------------------------------------
// largest magnitude of a [rows, cols] matrix with row pitch ld; contiguous 16-byte aligned matrices (the usual case)
// are read as one float4 stream, 4 loads in flight per thread
__global__ void __launch_bounds__(256) amax_kernel(const float* __restrict__ x, long long rows, long long cols, long long ld,
                                                   float* amax) {
  float m = 0.f;
  const long long n = rows * cols;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
  if (ld == cols && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const long long n4 = n >> 2;
    long long i = tid;
    for (; i + 3 * nthr < n4; i += 4 * nthr) {
      const float4 a = __ldg(x4 + i), b = __ldg(x4 + i + nthr), c = __ldg(x4 + i + 2 * nthr), d = __ldg(x4 + i + 3 * nthr);
      m = fmaxf(m, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
      m = fmaxf(m, fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w))));
      m = fmaxf(m, fmaxf(fmaxf(fabsf(c.x), fabsf(c.y)), fmaxf(fabsf(c.z), fabsf(c.w))));
      m = fmaxf(m, fmaxf(fmaxf(fabsf(d.x), fabsf(d.y)), fmaxf(fabsf(d.z), fabsf(d.w))));
    }
    for (; i < n4; i += nthr) {
      const float4 a = __ldg(x4 + i);
      m = fmaxf(m, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
    }
    for (long long j = (n4 << 2) + tid; j < n; j += nthr) m = fmaxf(m, fabsf(__ldg(x + j)));
  } else {
    for (long long i = tid; i < n; i += nthr) {
      const long long r = i / cols, c = i - r * cols;
      m = fmaxf(m, fabsf(x[r * ld + c]));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<int*>(amax), __float_as_int(m));
}

__device__ __forceinline__ float pow2_scale(float amax) {
  if (!(amax > 0.f) || !isfinite(amax)) return 1.f;
  int e;
  frexpf(amax, &e);                   // amax = m 2^e, m in [0.5, 1)  ->  amax 2^(9-e) in [256, 512)
  return ldexpf(1.f, 9 - e);
}

// hi / lo terms of 4 values and their placement: segment order [hi, lo, hi] (which == 0, the A side) or
// [hi, hi, lo] (which == 1, the B side)
__device__ __forceinline__ void split4(const float (&v)[4], float sc, uint2& hi, uint2& lo) {
  const __half2 h01 = __floats2half2_rn(v[0] * sc, v[1] * sc), h23 = __floats2half2_rn(v[2] * sc, v[3] * sc);
  const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
  const __half2 l01 = __floats2half2_rn(v[0] * sc - f01.x, v[1] * sc - f01.y);
  const __half2 l23 = __floats2half2_rn(v[2] * sc - f23.x, v[3] * sc - f23.y);
  hi.x = *reinterpret_cast<const uint32_t*>(&h01); hi.y = *reinterpret_cast<const uint32_t*>(&h23);
  lo.x = *reinterpret_cast<const uint32_t*>(&l01); lo.y = *reinterpret_cast<const uint32_t*>(&l23);
}

// operand rows = src rows, reduction = src columns
__global__ void __launch_bounds__(256) split_prep_kernel(const float* __restrict__ src, long long R, int C, long long ld,
                                                         int Kp, int which, int terms, const float* __restrict__ amax,
                                                         __half* __restrict__ dst) {
  const float sc = pow2_scale(*amax);
  const int nq = Kp >> 2;
  const long long total = R * nq;
  const bool vec = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / nq;
    const int c = (int)(i - r * nq) * 4;
    float v[4];
    if (vec && c + 4 <= C) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(src + r * ld + c));
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = (c + u < C) ? __ldg(src + r * ld + c + u) : 0.f;
    }
    uint2 hi, lo;
    split4(v, sc, hi, lo);
    __half* o = dst + (size_t)r * ((size_t)terms * Kp) + c;
    *reinterpret_cast<uint2*>(o) = hi;
    if (terms == 3) {
      *reinterpret_cast<uint2*>(o + Kp) = which == 0 ? lo : hi;
      *reinterpret_cast<uint2*>(o + 2 * (size_t)Kp) = which == 0 ? hi : lo;
    }
  }
}

// operand rows = src COLUMNS, reduction = src rows (a 32 x 32 tile goes through shared memory)
__global__ void __launch_bounds__(256) split_prep_t_kernel(const float* __restrict__ src, long long R, int C, long long ld,
                                                           long long Kp, int which, int terms,
                                                           const float* __restrict__ amax, __half* __restrict__ dst) {
  __shared__ float tile[32][33];
  const float sc = pow2_scale(*amax);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;        // 8 warps
  const long long tiles_r = Kp / 32;                             // reduction (src rows), padded
  const int tiles_c = (C + 31) / 32;
  for (long long t = blockIdx.x; t < tiles_r * tiles_c; t += gridDim.x) {
    const long long tr = t / tiles_c;
    const int tc = (int)(t - tr * tiles_c);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long r = tr * 32 + ty * 4 + j;
      const int c = tc * 32 + tx;
      tile[ty * 4 + j][tx] = (r < R && c < C) ? __ldg(src + r * ld + c) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = tc * 32 + ty * 4 + j;                        // operand row
      if (c < C) {
        const float v = tile[tx][ty * 4 + j] * sc;
        const __half h = __float2half_rn(v);
        const __half l = __float2half_rn(v - __half2float(h));
        __half* o = dst + (size_t)c * ((size_t)terms * Kp) + tr * 32 + tx;
        o[0] = h;
        if (terms == 3) {
          o[Kp] = which == 0 ? l : h;
          o[2 * Kp] = which == 0 ? h : l;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// the GEMM
// ------------------------------------------------------------------------------------------
struct GemmParams {
  long long M;
  int N, BN;
  long long n_panels;               // terms * Kp / 64
  int tiles_n, split_k;
  long long n_tiles;                // tiles_m * tiles_n * split_k
  const float* amax_a;
  const float* amax_b;
  const float* bias;                // [N] or null
  float* C;
  long long ldc;
  float alpha;
  int atomic;                       // split-K: accumulate with red.add into a zeroed / pre-filled C
  const int* m_dev;                 // optional: the row count lives on the device (min(*m_dev, M) rows are multiplied)
};

template <int DUMMY>
__global__ void __launch_bounds__(GM_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams P) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* gbase = smem_dyn + (base - raw);
  const uint32_t b_stage = (uint32_t)(P.BN / 2) * GM_KC * 2;
  const uint32_t stage = GM_A_STAGE + ((b_stage + 1023u) & ~1023u);
  const uint32_t bars = base + GM_STAGES * stage;
  auto bar_full = [&](int s) { return bars + 8u * s; };
  auto bar_empty = [&](int s) { return bars + 8u * (GM_STAGES + s); };
  auto bar_accfull = [&](int a) { return bars + 8u * (2 * GM_STAGES + a); };
  auto bar_accempty = [&](int a) { return bars + 8u * (2 * GM_STAGES + 2 + a); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + GM_STAGES * stage + 8 * (2 * GM_STAGES + 4));
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long n_pairs = gridDim.x / 2, pair = blockIdx.x / 2;

  if (threadIdx.x == 0) {
    for (int s = 0; s < GM_STAGES; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(bar_accfull(a), 1); mbar_init(bar_accempty(a), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long per_split = (P.n_panels + P.split_k - 1) / P.split_k;
  // a device-side row count shrinks the tile range (the operand buffers are sized for P.M rows)
  const long long M = P.m_dev ? min((long long)max(__ldg(P.m_dev), 0), P.M) : P.M;
  const long long n_tiles = P.m_dev ? ((M + 2 * GM_TM - 1) / (2 * GM_TM)) * P.tiles_n * P.split_k : P.n_tiles;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    uint32_t s = 0, ph = 0;
    for (long long t = pair; t < n_tiles; t += n_pairs) {
      const int ks = (int)(t % P.split_k);
      const long long tn = (t / P.split_k) % P.tiles_n, tm = t / ((long long)P.split_k * P.tiles_n);
      const long long p0 = ks * per_split, p1 = min(P.n_panels, p0 + per_split);
      const int arow = (int)(tm * 2 * GM_TM + cta_rank * GM_TM);
      const int brow = (int)(tn * P.BN + cta_rank * (P.BN / 2));
      for (long long p = p0; p < p1; ++p) {
        mbar_wait(bar_empty(s), ph ^ 1u);
        if (elect_one()) {
          if (leader) mbar_expect_tx(bar_full(s), 2u * GM_A_STAGE + 2u * b_stage);        // bytes of both CTAs
          tma_load_2d<2>(base + s * stage, &tmA, (int)(p * GM_KC), arow, bar_full(s));
          tma_load_2d<2>(base + s * stage + GM_A_STAGE, &tmB, (int)(p * GM_KC), brow, bar_full(s));
        }
        __syncwarp();
        if (++s == GM_STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (leader CTA) ===========================
    if (leader) {
      uint32_t s = 0, ph = 0, it = 0;
      const uint32_t idesc = umma_idesc(2 * GM_TM, P.BN);
      for (long long t = pair; t < n_tiles; t += n_pairs, ++it) {
        const int ks = (int)(t % P.split_k);
        const long long p0 = ks * per_split, p1 = min(P.n_panels, p0 + per_split);
        const uint32_t as = it & 1u, around = it >> 1;
        mbar_wait(bar_accempty(as), (around & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 256u;
        for (long long p = p0; p < p1; ++p) {
          mbar_wait(bar_full(s), ph);
          tc_fence_after();
          const uint64_t ad0 = umma_desc(base + s * stage, 1024, 2);
          const uint64_t bd0 = umma_desc(base + s * stage + GM_A_STAGE, 1024, 2);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < GM_KC / 16; ++kk)
              tc_mma_f16<2>(d_tmem, ad0 + 2u * kk, bd0 + 2u * kk, idesc, (p != p0 || kk != 0) ? 1u : 0u);
            tc_commit<2>(bar_empty(s));
            if (p == p1 - 1) tc_commit<2>(bar_accfull(as));
          }
          __syncwarp();
          if (++s == GM_STAGES) { s = 0; ph ^= 1u; }
        }
        if (p1 <= p0 && elect_one()) tc_commit<2>(bar_accfull(as));      // empty K range (cannot happen: split_k <= n_panels)
      }
    }
  } else {
    // =========================== epilogue: 4 warps, thread = row ===========================
    const int q = warp & 3;                       // TMEM lane quarter this warp may access (warps 2..5 -> 2,3,0,1)
    const float inv = P.amax_a ? P.alpha / (pow2_scale(*P.amax_a) * pow2_scale(*P.amax_b)) : P.alpha;
    uint32_t it = 0;
    for (long long t = pair; t < n_tiles; t += n_pairs, ++it) {
      const int ks = (int)(t % P.split_k);
      const long long tn = (t / P.split_k) % P.tiles_n, tm = t / ((long long)P.split_k * P.tiles_n);
      const uint32_t as = it & 1u, around = it >> 1;
      const long long row = tm * 2 * GM_TM + cta_rank * GM_TM + q * 32 + lane;
      const int col0 = (int)(tn * P.BN);
      mbar_wait(bar_accfull(as), around & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * 256u;
      float* crow = P.C + row * P.ldc;
      const bool vec = (P.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(P.C) & 15) == 0) && !P.atomic;
      for (int c = 0; c < P.BN; c += 16) {
        uint32_t v[16];
        tc_ld16(taddr + (uint32_t)c, v);
        tc_wait_ld();
        if (row < M) {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const int col = col0 + c + 4 * j4;
            float o[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              o[u] = __uint_as_float(v[4 * j4 + u]) * inv;
              if (P.bias && ks == 0 && col + u < P.N) o[u] += __ldg(P.bias + col + u);
            }
            if (vec && col + 4 <= P.N) {
              *reinterpret_cast<float4*>(crow + col) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if (col + u < P.N) {
                  if (P.atomic) atomicAdd(crow + col + u, o[u]);
                  else crow[col + u] = o[u];
                }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(bar_accempty(as));
        else mbar_arrive_cluster(bar_accempty(as), 0);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

inline size_t al256g(size_t v) { return (v + 255) / 256 * 256; }
inline long long kp_of(long long k) { return (k + 63) / 64 * 64; }

}  // namespace

// workspace: [ amax_a (256 B) | amax_b (256 B) | A' | B' ]
size_t gemm_workspace_bytes(int64_t M, int N, int64_t K, int single_term) {
  const size_t kp = (size_t)kp_of(K), terms = single_term ? 1 : 3;
  return 512 + al256g((size_t)M * terms * kp * 2) + al256g((size_t)N * terms * kp * 2) + 256;
}

int launch_gemm_f32(const float* A, int64_t lda, int transA, const float* B, int64_t ldb, int transB, int64_t M, int N,
                    int64_t K, const float* bias, float* C, int64_t ldc, float alpha, int accumulate, int single_term,
                    void* ws, cudaStream_t st) {
  const long long Kp = kp_of(K);
  const int terms = single_term ? 1 : 3;
  char* w = reinterpret_cast<char*>(ws);
  float* amax_a = reinterpret_cast<float*>(w);
  float* amax_b = reinterpret_cast<float*>(w + 256);
  __half* Ap = reinterpret_cast<__half*>(w + 512);
  __half* Bp = reinterpret_cast<__half*>(w + 512 + al256g((size_t)M * terms * (size_t)Kp * 2));
  G2V_CUDA_CHECK(cudaMemsetAsync(w, 0, 512, st));
  const int sms = num_sms();
  auto blocks = [&](long long items, int per) { long long g = (items + per - 1) / per; return (int)std::max<long long>(1, std::min<long long>(g, (long long)sms * 8)); };
  // stored shapes: A is [M, K] (or [K, M] if transA), B is [N, K] (or [K, N] if transB)
  {
    const long long ra = transA ? K : M, ca = transA ? M : K;
    amax_kernel<<<blocks(ra * ca, 1024), 256, 0, st>>>(A, ra, ca, lda, amax_a);
    G2V_LAUNCH_CHECK("amax_kernel");
    const long long rb = transB ? K : N, cbn = transB ? N : K;
    amax_kernel<<<blocks(rb * cbn, 1024), 256, 0, st>>>(B, rb, cbn, ldb, amax_b);
    G2V_LAUNCH_CHECK("amax_kernel");
  }
  if (!transA) {
    split_prep_kernel<<<blocks(M * (Kp / 4), 256), 256, 0, st>>>(A, M, (int)K, lda, (int)Kp, 0, terms, amax_a, Ap);
    G2V_LAUNCH_CHECK("split_prep_kernel");
  } else {
    split_prep_t_kernel<<<blocks((Kp / 32) * ((M + 31) / 32), 1), 256, 0, st>>>(A, K, (int)M, lda, Kp, 0, terms, amax_a, Ap);
    G2V_LAUNCH_CHECK("split_prep_t_kernel");
  }
  if (!transB) {
    split_prep_kernel<<<blocks((long long)N * (Kp / 4), 256), 256, 0, st>>>(B, N, (int)K, ldb, (int)Kp, 1, terms, amax_b, Bp);
    G2V_LAUNCH_CHECK("split_prep_kernel");
  } else {
    split_prep_t_kernel<<<blocks((Kp / 32) * ((N + 31) / 32), 1), 256, 0, st>>>(B, K, N, ldb, Kp, 1, terms, amax_b, Bp);
    G2V_LAUNCH_CHECK("split_prep_t_kernel");
  }
  // tile geometry: BN = the multiple of 16 (<= 256) that covers N in the fewest tiles with the least padding
  int tiles_n = (N + 255) / 256;
  int BN = (int)(((long long)(N + tiles_n - 1) / tiles_n + 15) / 16 * 16);
  if (BN < 16) BN = 16;
  const long long tiles_m = (M + 2 * GM_TM - 1) / (2 * GM_TM);
  const long long n_panels = terms * Kp / GM_KC;
  const long long out_tiles = tiles_m * tiles_n;
  const int n_pairs_max = sms / 2;
  int split_k = 1;
  if (out_tiles < n_pairs_max && n_panels >= 64)
    split_k = (int)std::min<long long>(n_panels / 16, (n_pairs_max + out_tiles - 1) / out_tiles);
  if (split_k < 1) split_k = 1;
  {   // every split must own at least one panel (an empty split would add an unwritten accumulator)
    const long long per = (n_panels + split_k - 1) / split_k;
    split_k = (int)((n_panels + per - 1) / per);
  }
  const int atomic = (split_k > 1 || accumulate) ? 1 : 0;
  if (split_k > 1 && !accumulate) {
    // zero the output block row by row (ldc may exceed N)
    if (ldc == N) G2V_CUDA_CHECK(cudaMemsetAsync(C, 0, (size_t)M * N * sizeof(float), st));
    else G2V_CUDA_CHECK(cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)M, st));
  }
  alignas(64) CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_map(&tmA, Ap, (uint64_t)M, (uint64_t)(terms * Kp), GM_KC, GM_TM, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_map(&tmB, Bp, (uint64_t)N, (uint64_t)(terms * Kp), GM_KC, (uint32_t)(BN / 2), CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  GemmParams P;
  P.M = M; P.N = N; P.BN = BN; P.n_panels = n_panels; P.tiles_n = tiles_n; P.split_k = split_k;
  P.n_tiles = out_tiles * split_k;
  P.amax_a = amax_a; P.amax_b = amax_b; P.bias = bias; P.C = C; P.ldc = ldc; P.alpha = alpha; P.atomic = atomic;
  P.m_dev = nullptr;
  const uint32_t b_stage = (uint32_t)(BN / 2) * GM_KC * 2;
  const size_t smem = (size_t)GM_STAGES * (GM_A_STAGE + ((b_stage + 1023u) & ~1023u)) + 8 * (2 * GM_STAGES + 4) + 16 + 1024;
  if ((rc = set_dyn_smem(reinterpret_cast<const void*>(&tc_gemm_kernel<0>), smem))) return rc;
  const long long pairs = std::min<long long>(n_pairs_max, P.n_tiles);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(pairs * 2));
  cfg.blockDim = dim3(GM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  G2V_CUDA_CHECK(cudaLaunchKernelEx(&cfg, tc_gemm_kernel<0>, tmA, tmB, P));
  G2V_LAUNCH_CHECK("tc_gemm_kernel");
  return G2V_OK;
}

// The product of two operands that are ALREADY in the kernel's fp16 layout (row-major [rows, kp16], kp16 a multiple
// of 64, 256-byte aligned bases): C[m, n] = sum_j A16[m, j] B16[n, j], raw fp32 accumulators, for the first
// min(*m_dev, m_cap) rows of A16.  The exact re-rank of the search uses it (g2v_tc.cu, refine pass): the row count
// is produced on the device by the kernel before it, so the tile range is resolved inside the kernel.
int launch_gemm_prepared(const __half* A16, const int* m_dev, long long m_cap, const __half* B16, int N, long long kp16,
                         float* C, long long ldc, cudaStream_t st) {
  const int sms = num_sms();
  int tiles_n = (N + 255) / 256;
  int BN = (int)(((long long)(N + tiles_n - 1) / tiles_n + 15) / 16 * 16);
  if (BN < 16) BN = 16;
  const long long tiles_m = (m_cap + 2 * GM_TM - 1) / (2 * GM_TM);
  alignas(64) CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_map(&tmA, A16, (uint64_t)m_cap, (uint64_t)kp16, GM_KC, GM_TM, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_map(&tmB, B16, (uint64_t)N, (uint64_t)kp16, GM_KC, (uint32_t)(BN / 2), CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  GemmParams P;
  P.M = m_cap; P.N = N; P.BN = BN; P.n_panels = kp16 / GM_KC; P.tiles_n = tiles_n; P.split_k = 1;
  P.n_tiles = tiles_m * tiles_n;
  P.amax_a = nullptr; P.amax_b = nullptr; P.bias = nullptr; P.C = C; P.ldc = ldc; P.alpha = 1.f; P.atomic = 0;
  P.m_dev = m_dev;
  const uint32_t b_stage = (uint32_t)(BN / 2) * GM_KC * 2;
  const size_t smem = (size_t)GM_STAGES * (GM_A_STAGE + ((b_stage + 1023u) & ~1023u)) + 8 * (2 * GM_STAGES + 4) + 16 + 1024;
  if ((rc = set_dyn_smem(reinterpret_cast<const void*>(&tc_gemm_kernel<0>), smem))) return rc;
  const long long pairs = std::max<long long>(1, std::min<long long>(sms / 2, P.n_tiles));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(pairs * 2));
  cfg.blockDim = dim3(GM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  G2V_CUDA_CHECK(cudaLaunchKernelEx(&cfg, tc_gemm_kernel<0>, tmA, tmB, P));
  G2V_LAUNCH_CHECK("tc_gemm_kernel");
  return G2V_OK;
}

}  // namespace g2v
