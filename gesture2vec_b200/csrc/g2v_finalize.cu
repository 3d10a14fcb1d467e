// Per-step K x D work of the vector-quantizer path in ONE cooperative launch (sm_100a):
//
//   phase A   pack the row-pass accumulators into the fp32 statistics buffer
//             [dwr (K*D) | counts (K) | sse | rows] (summing the dwr replicas) and hand the accumulators
//             back zeroed; reset the codebook aux header
//   phase B   loss / perplexity from the (possibly all-reduced) statistics; EMA cluster sizes (Laplace
//             smoothing needs the sum over all codes: every block reduces it redundantly, in the same order);
//             per code row: EMA weights and the new codebook (or the Lloyd centre update), and that row's
//             ||e||^2 (fp64 accumulate), |e|_max and norm bucket for the aux buffer
//   phase C   norm-table prefix maximum, fp16 operand copy of the new codebook at the power-of-two scale
//             its |e|_max asks for, relative rounding residual `sfrac`
//
// with grid-wide barriers between the phases (cooperative launch: the runtime refuses the launch rather than
// deadlock if the grid cannot be co-resident).  The same kernel, with phases switched off, serves
// g2v_codebook_prepare, g2v_vq_stats_finalize, g2v_vq_ema_update and g2v_kmeans_update, so a training step
// ends in one launch instead of nine (stats_pack, stats_finalize, ema_cs, ema_w, 2 memsets + 3 codebook kernels).
//
// Reference statements: DAE_model.py:340-347, 451-481; Autoencoder_VQVAE_model.py:1163-1172, 1262-1294.
#include "g2v_common.cuh"

#include <cooperative_groups.h>
#include <math.h>

namespace cg = cooperative_groups;

namespace g2v {
namespace {

constexpr int FIN_THREADS = 256;
constexpr int FIN_WARPS = FIN_THREADS / 32;

struct FinParams {
  int K, D, Kp, Dp;
  long long rows_local;
  // phase A
  int* counts;            // int32[K] accumulator (zeroed on exit) or null
  double* sse;            // double[1] accumulator (zeroed on exit) or null
  float* dwr;             // fp32 [reps][K*D] accumulators (zeroed on exit) or null
  int reps, do_pack;
  float* packed;          // [K*D + K + 2]; null only when nothing below needs it
  // scalar outputs
  float coef_codebook, coef_commit;
  float* loss;
  float* ppl;
  // codebook update: 0 none, 1 EMA, 2 Lloyd centre update
  int mode;
  const float* cs_in;
  float* cs_out;
  const float* w_in;
  float* w_out;
  const float* E_old;
  float* E_new;
  float* E_prev;          // optional copy of E_old (EMA mode), for in-place updates
  float decay, one_m, eps, keps;
  double* shift2;
  // aux buffer to prepare (null = skip) for the codebook E_cb (== E_new when mode != 0)
  void* cb;
  const float* E_cb;
};

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// sum over the block in a fixed order (thread-strided partials -> xor tree -> warps in ascending order): every
// block that runs it over the same data gets the same bits
__device__ double block_sum_fixed(double v, double* sh) {
  v = warp_sum_d(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < FIN_WARPS; ++w) s += sh[w];
  return s;
}

__global__ void __launch_bounds__(FIN_THREADS) finalize_kernel(const FinParams P) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sh[FIN_WARPS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t gtid = (size_t)blockIdx.x * FIN_THREADS + tid, gn = (size_t)gridDim.x * FIN_THREADS;
  const int K = P.K, D = P.D;
  const size_t KD = (size_t)K * D;
  CbHeader* hdr = reinterpret_cast<CbHeader*>(P.cb);
  float* ntab = P.cb ? reinterpret_cast<float*>(reinterpret_cast<char*>(P.cb) + 256) : nullptr;
  float* e2 = P.cb ? reinterpret_cast<float*>(reinterpret_cast<char*>(P.cb) + 768) : nullptr;
  __half* e16 = P.cb ? reinterpret_cast<__half*>(reinterpret_cast<char*>(P.cb) + 768 + (size_t)P.Kp * 4) : nullptr;

  // ------------------------------------------------------------------ phase A
  if (P.do_pack) {
    float* tail = P.packed + KD;
    for (size_t k = gtid; k < (size_t)K + 2; k += gn) {
      if (k < (size_t)K) {
        tail[k] = P.counts ? (float)P.counts[k] : 0.f;
        if (P.counts) P.counts[k] = 0;
      } else if (k == (size_t)K) {
        tail[k] = P.sse ? (float)(*P.sse) : 0.f;
        if (P.sse) *P.sse = 0.0;
      } else {
        tail[k] = (float)P.rows_local;
      }
    }
    const bool v4 = (KD % 4 == 0) && ((reinterpret_cast<uintptr_t>(P.packed) & 15) == 0) &&
                    (!P.dwr || (reinterpret_cast<uintptr_t>(P.dwr) & 15) == 0);
    if (v4) {
      const size_t n4 = KD / 4;
      const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
      for (size_t i = gtid; i < n4; i += gn) {
        float4 a = zero;
        for (int r = 0; r < P.reps; ++r) {              // replicas in a fixed order
          float4* src = reinterpret_cast<float4*>(P.dwr + (size_t)r * KD) + i;
          const float4 v = __ldcg(src);
          a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
          *src = zero;
        }
        reinterpret_cast<float4*>(P.packed)[i] = a;
      }
    } else {
      for (size_t i = gtid; i < KD; i += gn) {
        float a = 0.f;
        for (int r = 0; r < P.reps; ++r) {
          a += __ldcg(P.dwr + (size_t)r * KD + i);
          P.dwr[(size_t)r * KD + i] = 0.f;
        }
        P.packed[i] = a;
      }
    }
  }
  if (P.cb && blockIdx.x == 0) {                          // header + norm table reset (768 bytes)
    uint32_t* w = reinterpret_cast<uint32_t*>(P.cb);
    for (int i = tid; i < 192; i += FIN_THREADS) w[i] = 0u;
    __syncthreads();
    if (tid == 0) hdr->e2min = __int_as_float(0x7f7f7f7f);   // large positive for atomicMin on the bit pattern
  }
  grid.sync();

  // ------------------------------------------------------------------ phase B
  const float* tail = P.packed ? P.packed + KD : nullptr;
  if (blockIdx.x == 0 && (P.loss || P.ppl)) {
    const float rows = __ldcg(tail + K + 1);
    double h = 0.0;
    for (int k = tid; k < K; k += FIN_THREADS) {
      const float p = __ldcg(tail + k) / rows;            // avg_probs = mean(encodings, 0)
      h += (double)(p * logf(p + 1e-10f));
    }
    h = block_sum_fixed(h, sh);
    if (tid == 0) {
      if (P.ppl) *P.ppl = expf(-(float)h);
      if (P.loss) {
        const float mse = (float)((double)__ldcg(tail + K) / ((double)rows * (double)D));
        *P.loss = __fadd_rn(__fmul_rn(P.coef_codebook, mse), __fmul_rn(P.coef_commit, mse));
      }
    }
  }
  float n_cs = 0.f, den_cs = 1.f;
  if (P.mode == 1) {
    // cs <- cs*decay + (1-decay)*counts ; n = sum cs ; cs <- (cs + eps) / (n + K*eps) * n
    double part = 0.0;
    for (int k = tid; k < K; k += FIN_THREADS)
      part += (double)__fadd_rn(__fmul_rn(P.cs_in[k], P.decay), __fmul_rn(P.one_m, __ldcg(tail + k)));
    n_cs = (float)block_sum_fixed(part, sh);
    den_cs = __fadd_rn(n_cs, P.keps);
  }
  // cs_out may alias cs_in (in-place state for CUDA-graph replay): it is written by block 0 only after a
  // grid-wide barrier, once every block has read cs_in for the sum above and for its own code rows
  auto write_cs = [&]() {
    if (blockIdx.x == 0)
      for (int k = tid; k < K; k += FIN_THREADS) {
        const float v = __fadd_rn(__fmul_rn(P.cs_in[k], P.decay), __fmul_rn(P.one_m, __ldcg(tail + k)));
        P.cs_out[k] = __fmul_rn(__fdiv_rn(__fadd_rn(v, P.eps), den_cs), n_cs);
      }
  };
  double shift_part = 0.0;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vec_rows = (D % 4 == 0) && D <= 512 && al16(P.E_old) && al16(P.E_new) && al16(P.E_cb) && al16(P.packed) &&
                        al16(P.w_in) && al16(P.w_out) && al16(P.E_prev);
  if (P.mode != 0 || P.cb) {
    for (int k = blockIdx.x * FIN_WARPS + warp; k < P.Kp; k += gridDim.x * FIN_WARPS) {
      if (k >= K) {
        if (P.cb && lane == 0) e2[k] = INFINITY;
        continue;
      }
      const size_t r0 = (size_t)k * D;
      const float cnt = P.mode ? __ldcg(tail + k) : 0.f;
      float csn = 1.f;
      if (P.mode == 1) {
        const float v = __fadd_rn(__fmul_rn(P.cs_in[k], P.decay), __fmul_rn(P.one_m, cnt));
        csn = __fmul_rn(__fdiv_rn(__fadd_rn(v, P.eps), den_cs), n_cs);
      }
      double s2 = 0.0;
      float am = 0.f;
      // one code row.  Vector path (D % 4 == 0, 16-byte aligned rows, D <= 512): a lane owns the float4 columns
      // lane, lane + 32, lane + 64, lane + 96 and issues all of its loads before the first use, so the row costs
      // one memory round trip instead of D / 32 dependent ones (this launch is the latency of a small-batch step)
      auto elem = [&](float eo, float pk, float wi, float& wo, float& en) {
        float e;
        if (P.mode == 1) {
          const float dw = fmaf(cnt, eo, pk);              // sum of rows = residual sum + count * code
          wo = __fadd_rn(__fmul_rn(wi, P.decay), __fmul_rn(P.one_m, dw));
          e = __fdiv_rn(wo, csn);
        } else if (P.mode == 2) {
          const float step = cnt > 0.f ? __fdiv_rn(pk, cnt) : 0.f;   // empty cluster keeps its centre
          e = __fadd_rn(eo, step);
          shift_part += (double)step * (double)step;
        } else {
          e = eo;
        }
        en = e;
        s2 += (double)e * (double)e;
        am = fmaxf(am, fabsf(e));
      };
      const float* src_e = P.mode ? P.E_old : P.E_cb;
      if (vec_rows) {
        const int nq = D >> 2;
        float4 eo[4], pk[4], wi[4];
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = lane + 32 * u;
          const bool on = c < nq;
          eo[u] = on ? *reinterpret_cast<const float4*>(src_e + r0 + 4 * c) : z4;
          pk[u] = (on && P.mode) ? __ldcg(reinterpret_cast<const float4*>(P.packed + r0 + 4 * c)) : z4;
          wi[u] = (on && P.mode == 1) ? *reinterpret_cast<const float4*>(P.w_in + r0 + 4 * c) : z4;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = lane + 32 * u;
          if (c < nq) {
            float4 wo = z4, en;
            elem(eo[u].x, pk[u].x, wi[u].x, wo.x, en.x);
            elem(eo[u].y, pk[u].y, wi[u].y, wo.y, en.y);
            elem(eo[u].z, pk[u].z, wi[u].z, wo.z, en.z);
            elem(eo[u].w, pk[u].w, wi[u].w, wo.w, en.w);
            if (P.mode == 1) {
              if (P.E_prev) *reinterpret_cast<float4*>(P.E_prev + r0 + 4 * c) = eo[u];   // the backward pass still needs the old codes
              *reinterpret_cast<float4*>(P.w_out + r0 + 4 * c) = wo;
            }
            if (P.mode) *reinterpret_cast<float4*>(P.E_new + r0 + 4 * c) = en;
          }
        }
      } else {
        for (int j = lane; j < D; j += 32) {
          const float eo = src_e[r0 + j];
          const float pk = P.mode ? __ldcg(P.packed + r0 + j) : 0.f;
          const float wi = P.mode == 1 ? P.w_in[r0 + j] : 0.f;
          float wo = 0.f, en;
          elem(eo, pk, wi, wo, en);
          if (P.mode == 1) {
            if (P.E_prev) P.E_prev[r0 + j] = eo;
            P.w_out[r0 + j] = wo;
          }
          if (P.mode) P.E_new[r0 + j] = en;
        }
      }
      if (P.cb) {
        s2 = warp_sum_d(s2);
        am = warp_max_f(am);
        if (lane == 0) {
          const float f2 = (float)s2;
          e2[k] = f2;
          // non-negative floats order like their bit patterns
          atomicMax(reinterpret_cast<int*>(&hdr->e2max), __float_as_int(f2));
          atomicMin(reinterpret_cast<int*>(&hdr->e2min), __float_as_int(f2));
          atomicMax(reinterpret_cast<int*>(&hdr->amax), __float_as_int(am));
          const float nrm = sqrtf(f2) * 1.0001f;
          atomicMax(reinterpret_cast<int*>(&ntab[norm_bucket(nrm)]), __float_as_int(nrm));
        }
      }
    }
  }
  if (P.mode == 2 && P.shift2) {
    const double t = block_sum_fixed(shift_part, sh);
    if (tid == 0 && t != 0.0) atomicAdd(P.shift2, t);
  }
  if (!P.cb) {                                             // uniform over the grid
    if (P.mode == 1) {
      grid.sync();
      write_cs();
    }
    return;
  }
  grid.sync();
  if (P.mode == 1) write_cs();

  // ------------------------------------------------------------------ phase C
  const float amax = __ldcg(&hdr->amax);
  float sc = 1.f;
  if (amax > 0.f && isfinite(amax)) {
    int e;
    // amax = m * 2^e, m in [0.5,1)  ->  amax * 2^(15-e) in [16384,32768): the top of the fp16 range, so that codes
    // 2^20 and more below the largest entry (live codes next to dead EMA codes) still land on normal numbers
    frexpf(amax, &e);
    sc = ldexpf(1.f, 15 - e);
  }
  const float inv = 1.f / sc;      // power of two: exact
  if (blockIdx.x == 0 && warp == 0) {
    // prefix maximum over the 128 norm buckets: 4 per lane, then an inclusive scan over the lanes
    float v[4];
    float run = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) { run = fmaxf(run, __ldcg(ntab + 4 * lane + i)); v[i] = run; }
    float incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl = fmaxf(incl, up);
    }
    const float before = __shfl_up_sync(0xffffffffu, incl, 1);
#pragma unroll
    for (int i = 0; i < 4; ++i) ntab[4 * lane + i] = (lane > 0) ? fmaxf(v[i], before) : v[i];
    if (lane == 0) {
      hdr->scale_e = sc;
      hdr->K = K; hdr->D = D; hdr->Kp = P.Kp; hdr->Dp = P.Dp;
      hdr->magic = kCbMagic;
    }
  }
  const bool vec_c = (D % 4 == 0) && P.Dp <= 512 && al16(P.E_cb);
  for (int k = blockIdx.x * FIN_WARPS + warp; k < P.Kp; k += gridDim.x * FIN_WARPS) {
    __half* o = e16 + (size_t)k * P.Dp;
    float s2 = 0.f, n2 = 0.f, u2 = 0.f;                    // residual^2 on normal / norm^2 / residual^2 on subnormal elements
    if (vec_c) {                                           // float4 in, 4 x fp16 (8 bytes) out; Dp % 16 == 0
      const int nq = D >> 2, nqp = P.Dp >> 2;
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = lane + 32 * u;
        v[u] = (k < K && c < nq) ? __ldcg(reinterpret_cast<const float4*>(P.E_cb + (size_t)k * D + 4 * c))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = lane + 32 * u;
        if (c < nqp) {
          const float vv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
          const __half2 h01 = __floats2half2_rn(vv[0] * sc, vv[1] * sc), h23 = __floats2half2_rn(vv[2] * sc, vv[3] * sc);
          const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
          const float back[4] = {f01.x, f01.y, f23.x, f23.y};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float r = vv[t] - back[t] * inv;
            if (fabsf(vv[t] * sc) >= 6.1035156e-5f) s2 = fmaf(r, r, s2);
            else u2 = fmaf(r, r, u2);
            n2 = fmaf(vv[t], vv[t], n2);
          }
          uint2 pk2;
          pk2.x = *reinterpret_cast<const uint32_t*>(&h01);
          pk2.y = *reinterpret_cast<const uint32_t*>(&h23);
          *reinterpret_cast<uint2*>(o + 4 * c) = pk2;
        }
      }
    } else {
      for (int j = lane; j < P.Dp; j += 32) {
        const float v = (k < K && j < D) ? __ldcg(P.E_cb + (size_t)k * D + j) : 0.f;
        const __half h = __float2half_rn(v * sc);
        const float r = v - __half2float(h) * inv;
        if (fabsf(v * sc) >= 6.1035156e-5f) s2 = fmaf(r, r, s2);
        else u2 = fmaf(r, r, u2);
        n2 = fmaf(v, v, n2);
        o[j] = h;
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      s2 += __shfl_xor_sync(0xffffffffu, s2, off);
      n2 += __shfl_xor_sync(0xffffffffu, n2, off);
      u2 += __shfl_xor_sync(0xffffffffu, u2, off);
    }
    // normal elements round relative to themselves, so their residual scales with the code's own norm; what
    // lands on fp16 subnormals rounds on a fixed grid and is bounded absolutely:  ||r_e,k|| <= sfrac ||e_k|| + rsub
    if (lane == 0 && k < K && n2 > 0.f) {
      atomicMax(reinterpret_cast<int*>(&hdr->sfrac), __float_as_int(sqrtf(s2 / n2) * 1.001f));
      if (u2 > 0.f) atomicMax(reinterpret_cast<int*>(&hdr->rsub), __float_as_int(sqrtf(u2) * 1.001f));
    }
  }
}

// Fold a linear projection into a codebook (quantizers.py VQVAE_VQ_Payam_EMA._fold): for code k
//   out[k, 0..D)  = fp32( sum_i E[k,i] W[i,j] )                      -- W^T e_k, accumulated in fp64, rounded once
//   g[k]          = |e_k|^2 - 2 b.e_k - |out[k, 0..D)|^2   (fp64)    -- what the extra coordinate has to carry
// FOLD_CB codes per block, their rows widened to fp64 once in shared memory; threads stride the output columns (W rows
// are read coalesced and converted once per FOLD_CB products -- the fp32 -> fp64 conversions of a one-code-per-block
// version ran on the 16-lane XU pipe and made this 0.27 ms at K=512, D=400).
constexpr int FOLD_CB = 4;
constexpr int FOLD_THREADS = 512;               // one output column per thread (D <= 512), 16 warps to hide the W loads
__global__ void __launch_bounds__(FOLD_THREADS) fold_rows_kernel(const float* __restrict__ E, const float* __restrict__ W,
                                                                 const float* __restrict__ b, int K, int D, int ld_out,
                                                                 float* __restrict__ out, double* __restrict__ g) {
  __shared__ double sh[FOLD_CB][FOLD_THREADS / 32];
  extern __shared__ double er[];                // [FOLD_CB][D] code rows in fp64
  const int k0 = blockIdx.x * FOLD_CB;
  for (int t = threadIdx.x; t < FOLD_CB * D; t += blockDim.x) {
    const int c = t / D, i = t - c * D;
    er[t] = (k0 + c < K) ? (double)E[(size_t)(k0 + c) * D + i] : 0.0;
  }
  __syncthreads();
  double n2[FOLD_CB], cc[FOLD_CB];
#pragma unroll
  for (int c = 0; c < FOLD_CB; ++c) n2[c] = cc[c] = 0.0;
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    double acc[FOLD_CB];
#pragma unroll
    for (int c = 0; c < FOLD_CB; ++c) acc[c] = 0.0;
    int i = 0;
    for (; i + 8 <= D; i += 8) {                // eight W loads in flight per thread
      float w[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) w[u] = __ldg(W + (size_t)(i + u) * D + j);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const double wd = (double)w[u];
#pragma unroll
        for (int c = 0; c < FOLD_CB; ++c) acc[c] = fma(er[c * D + i + u], wd, acc[c]);
      }
    }
    for (; i < D; ++i) {
      const double wd = (double)__ldg(W + (size_t)i * D + j);
#pragma unroll
      for (int c = 0; c < FOLD_CB; ++c) acc[c] = fma(er[c * D + i], wd, acc[c]);
    }
    const double bj = (double)b[j];
#pragma unroll
    for (int c = 0; c < FOLD_CB; ++c) {
      const float f = (float)acc[c];
      if (k0 + c < K) out[(size_t)(k0 + c) * ld_out + j] = f;
      n2[c] += (double)f * (double)f;
      cc[c] += er[c * D + j] * (er[c * D + j] - 2.0 * bj);
    }
  }
#pragma unroll
  for (int c = 0; c < FOLD_CB; ++c) {
    const double v = warp_sum_d(cc[c] - n2[c]);
    if ((threadIdx.x & 31) == 0) sh[c][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < FOLD_CB && k0 + threadIdx.x < K) {
    const int c = threadIdx.x;
    double t = 0.0;
    for (int w = 0; w < FOLD_THREADS / 32; ++w) t += sh[c][w];
    g[k0 + c] = t;
  }
}

// largest cooperative grid of finalize_kernel on the current device (cached per thread and device)
int max_coop_grid() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 1;
  if (dev != cached_dev) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, finalize_kernel, FIN_THREADS, 0) != cudaSuccess || per_sm < 1)
      per_sm = 1;
    cached = per_sm * num_sms();
    cached_dev = dev;
  }
  return cached;
}

int launch_fin(FinParams& P, cudaStream_t st) {
  P.Kp = round_up(P.K, 256);
  P.Dp = round_up(P.D, 16);
  // enough warps for one code row each, enough threads for the pack pass; never more than can be co-resident
  long long want = 1;
  if (P.mode != 0 || P.cb) want = (P.Kp + FIN_WARPS - 1) / FIN_WARPS;
  if (P.do_pack) {
    // one float4 column of the replicas per thread: the pack pass reads and re-zeroes reps x K x D floats, and with
    // too few threads that is the whole cost of the launch (58 us at 64 blocks, K=512, 8 replicas)
    const long long pack_blocks = ((long long)P.K * P.D / 4 + FIN_THREADS - 1) / FIN_THREADS;
    if (pack_blocks > want) want = pack_blocks;
  }
  const long long cap = std::min<long long>(max_coop_grid(), 2LL * num_sms());
  const int grid = (int)std::max<long long>(1, std::min(want, cap));
  void* args[] = {&P};
  G2V_CUDA_CHECK(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(&finalize_kernel), dim3(grid), dim3(FIN_THREADS),
                                             args, 0, st));
  G2V_LAUNCH_CHECK("finalize_kernel");
  return G2V_OK;
}

FinParams fin_blank(int K, int D) {
  FinParams P = {};
  P.K = K;
  P.D = D;
  return P;
}

}  // namespace

int launch_fold_rows(const float* E, const float* W, const float* b, int K, int D, int ld_out, float* out, double* g,
                     cudaStream_t st) {
  fold_rows_kernel<<<(K + FOLD_CB - 1) / FOLD_CB, FOLD_THREADS, (size_t)FOLD_CB * D * sizeof(double), st>>>(E, W, b, K, D, ld_out, out, g);
  G2V_LAUNCH_CHECK("fold_rows_kernel");
  return G2V_OK;
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
int launch_codebook_prepare(const float* E, int K, int D, void* cb, cudaStream_t st) {
  FinParams P = fin_blank(K, D);
  P.cb = cb;
  P.E_cb = E;
  return launch_fin(P, st);
}

int launch_stats_finalize(const float* packed, int K, int D, float coef_codebook, float coef_commit, float* loss,
                          float* ppl, cudaStream_t st) {
  FinParams P = fin_blank(K, D);
  P.packed = const_cast<float*>(packed);
  P.coef_codebook = coef_codebook;
  P.coef_commit = coef_commit;
  P.loss = loss;
  P.ppl = ppl;
  return launch_fin(P, st);
}

int launch_step_finalize(int32_t* counts, double* sse, float* dwr, int dwr_replicas, int do_pack, int64_t rows_local,
                         float* packed, int K, int D, float coef_codebook, float coef_commit, float* loss, float* ppl,
                         int mode, const float* cs_in, float* cs_out, const float* w_in, float* w_out,
                         const float* E_old, float* E_new, float* E_prev, float decay, float eps, double* shift2,
                         void* cb, const float* E_cb, cudaStream_t st) {
  FinParams P = fin_blank(K, D);
  P.E_prev = E_prev;
  P.rows_local = rows_local;
  P.counts = counts;
  P.sse = sse;
  P.dwr = dwr;
  P.reps = dwr ? dwr_replicas : 0;
  P.do_pack = do_pack;
  P.packed = packed;
  P.coef_codebook = coef_codebook;
  P.coef_commit = coef_commit;
  P.loss = loss;
  P.ppl = ppl;
  P.mode = mode;
  P.cs_in = cs_in;
  P.cs_out = cs_out;
  P.w_in = w_in;
  P.w_out = w_out;
  P.E_old = E_old;
  P.E_new = E_new;
  P.decay = decay;
  P.one_m = (float)(1.0 - (double)decay);
  P.eps = eps;
  P.keps = (float)((double)K * (double)eps);
  P.shift2 = shift2;
  P.cb = cb;
  P.E_cb = mode != 0 ? E_new : E_cb;
  return launch_fin(P, st);
}

}  // namespace g2v
