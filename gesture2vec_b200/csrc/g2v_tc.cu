// Tensor-core nearest-code search for sm_100a: tcgen05.mma (fp16 operands, fp32 accumulators in
// TMEM) fed by TMA, fused with an argmin epilogue, so the [N,K] distance matrix never exists.
//
// run_tc() picks one of two sweep kernels, all on the caller's stream:
//   tc_tmem_kernel<ZT>        fp32 rows at any K, bf16 / fp16 rows up to 16 code tiles.  CTA pairs
//                             (cta_group::2, UMMA M = 256).  Rows stream by TMA into a ring of staging slots;
//                             8 converter warps round them to fp16 in registers, keep ||z||^2 and the exact
//                             rounding residual per row, and write the A operand straight into tensor memory
//                             (tcgen05.st); the codebook streams from L2, half of every stage per CTA.
//   tc_search_kernel<CG,FUSED> everything else.  The A operand is an fp16 row tile in shared memory: written by
//                             row_prep_kernel beforehand (with per-row constants; CTA pairs, deep codebook
//                             ring -- the large-K path) or converted in-kernel from an fp32 staging ring (FUSED).
// Warp roles in every variant: TMA producers (codebook stages / rows), one MMA-issuing warp (one elected lane;
// two accumulator stages in TMEM), 8 epilogue warps: tcgen05.ld -> d = e2 - 2 z.e as a 23-bit fixed-point
// key (packed FFMA2, FMNMX, IMAD) -> 32 running top-2 chains per row (3 VIMNMX per key).
// A row whose best codes are closer than its error bound tau goes to a candidate list (row + up to three
// codes), a chain list (one chain of K/32 codes may hide the winner) or the whole-row list.  rerank_kernel
// re-ranks the first two exactly in fp64; the whole rows of a bulk search take the refine pass first (split-fp16
// operands, tc_gemm_kernel, refine_rows_kernel: fp64 only on the codes an fp32-accurate bound cannot exclude) on a
// side stream next to it; full_recheck_kernel (g2v_simt.cu) takes a whole-row list that overflows the pass.
//
// Operand layouts: K-major, SWIZZLE_128B panels of 64 fp16 (TMA box 64 x rows) and, for the
// D % 64 remainder, SWIZZLE_32B panels of 16 fp16 (one UMMA_K step each).
#include "g2v_tcgen05.cuh"

#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <type_traits>

namespace g2v {
namespace {

constexpr int TM = 128;                 // rows per tile (UMMA M)
constexpr int TN = 256;                 // codes per accumulator stage (max UMMA N)
constexpr int KC = 64;                  // fp16 per full K panel (128 bytes)
constexpr int KT = 16;                  // fp16 per tail K panel (32 bytes) == UMMA K
constexpr int MAX_STAGES = 8;            // barrier slots; the ring depth itself is chosen per launch
constexpr int A_PANEL = TM * KC * 2;    // 16384
constexpr int A_TAIL = TM * KT * 2;     // 4096
constexpr int B_PANEL = TN * KC * 2;    // 32768
constexpr int B_TAIL = TN * KT * 2;     // 8192
constexpr int MAX_CHUNKS = 12;
constexpr int NTHREADS = 384;           // 4 producer/MMA warps + 8 epilogue warps
constexpr int NTHREADS_FUSED = 512;     // + 4 warps converting fp32 rows to the fp16 operand tile
constexpr int CONV_WARP0 = 12;
constexpr int MAX_ZSLOTS = 8;           // fp32 staging slots: 64 rows x 64 columns (256-byte rows keep the
constexpr int ZROWS = 64;               // TMA request count per byte low); the D % 64 tail uses 128 x 16
constexpr int ZSLOT = ZROWS * KC * 4;   // 16384 bytes
constexpr int kPfTiles = 2;             // L2 prefetch distance of the row tiles
constexpr int RS_RING = 4;              // row-statistics ring (tiles in flight between converter and epilogue)
constexpr int EPI_WARP0 = 4;
constexpr float kMagic = 12582912.0f;   // 1.5 * 2^23: float bits = 0x4B400000 + round(v)
constexpr float kKeyCap = 16777215.0f;  // kMagic + 2^22 - 1: largest key value
constexpr int kMaxDp = 496;
constexpr int kMaxK = 16384;            // 9-bit column-group field of the key
// role-level timeline (tools/tc_trace.py): compiled in only with -DG2V_TRACE=1, the trace state costs registers
#ifndef G2V_TRACE
#define G2V_TRACE 0
#endif
constexpr bool kTraceBuild = G2V_TRACE != 0;
// bring-up switches (ablation timing, results are wrong with them): only in a -DG2V_TRACE=1 build, where the
// G2V_TC_DEBUG environment variable sets them; in the product build the masks are 0 and every test on them folds away
constexpr unsigned kDbgOn = kTraceBuild ? 1u : 0u;
constexpr unsigned kDbgSkipMma = 0x10000u * kDbgOn, kDbgSkipEpi = 0x20000u * kDbgOn, kDbgSkipConv = 0x40000u * kDbgOn,
                   kDbgPrefetch = 0x80000u * kDbgOn, kDbgNoB = 0x100000u * kDbgOn, kDbgNoZ = 0x200000u * kDbgOn;

struct RowInfo {       // 32 bytes per row, written by row_prep_kernel
  float cS;            // -2 / (scale_z * scale_e) * S
  float S;             // fixed-point scale (power of two)
  float a1, a0;        // |dot_hat - dot| <= a1*c + a0 for a code of norm c
  float znorm, z2;     // ||z|| (rounded up), ||z||^2
  float pad0, pad1;
};
static_assert(sizeof(RowInfo) == 32, "RowInfo is read as two float4");

struct TcParams {
  long long N;
  int K, D, Dp;
  int n_full, n_tail, n_chunks;
  int n_ntiles, n_last_mma;     // code tiles; UMMA N of the last one (multiple of 16)
  int n_row_tiles;
  int nstage;                   // depth of the codebook-stage ring
  int nzslot;                   // fused variant: depth of the fp32 staging ring
  int n_ksteps;                 // Dp / 16
  const CbHeader* hdr;
  long long* trace;             // bring-up aid: [role][event] timestamps of CTA 0 (nullptr = off)
  const RowInfo* rowinfo;
  const float* ntab;            // reachable-norm table of the codebook (g2v_common.cuh)
  const float* e2;
  int* idx;
  int* pair_list;               // int4 per entry: row, code a, code b, code c (-1 if only two)
  int* full_list;               // 1 int per entry: row
  int* chain_list;              // int4 per entry: row, chain (code % 32), extra code a, extra code b (-1 = none)
  int* counters;                // [0] candidate entries, [1] full-row fallbacks, [2] chain entries
  unsigned flags;
};

// ------------------------------------------------------------------------------------------
// shared memory plan
// ------------------------------------------------------------------------------------------
struct SmemPlan {
  uint32_t a_off, b_off, b_stage, e2_off, xch_off, zst_off, rs_off, bar_off, tmem_off, total;
};
// cg = CTAs cooperating on one MMA (cta_group): each holds its own 128-row A tile and 1/cg of every
// codebook stage
// nzslot > 0: fused variant (fp32 staging ring + row statistics ring)
__host__ __device__ inline SmemPlan smem_plan(int n_full, int n_tail, int cg, int nstage, int nzslot = 0) {
  SmemPlan p;
  p.a_off = 0;
  uint32_t a_bytes = (uint32_t)n_full * A_PANEL + (uint32_t)n_tail * A_TAIL;
  p.b_off = (a_bytes + 1023u) & ~1023u;
  p.b_stage = B_PANEL / cg;
  p.e2_off = p.b_off + (uint32_t)nstage * p.b_stage;
  p.xch_off = p.e2_off + 2 * TN * 4;          // 128 rows x 12 words: epilogue half 1 -> half 0
  p.zst_off = (p.xch_off + TM * 12 * 4 + 127u) & ~127u;
  p.rs_off = p.zst_off + (uint32_t)nzslot * ZSLOT;
  p.bar_off = p.rs_off + (nzslot > 0 ? RS_RING * TM * 8 : 0);
  p.tmem_off = p.bar_off + 8 * (2 * MAX_STAGES + 2 * MAX_CHUNKS + 4 + 2 * MAX_ZSLOTS + RS_RING);
  p.total = p.tmem_off + 16;
  return p;
}

// ------------------------------------------------------------------------------------------
// epilogue helpers
// ------------------------------------------------------------------------------------------
// two fp32 FMAs in one instruction (FFMA2): halves the issue slots of the fp32 work in the epilogue and the
// converters, which share the four schedulers with everything else (-DG2V_PACKED_F32=0 keeps scalar FFMA)
#ifndef G2V_PACKED_F32
#define G2V_PACKED_F32 1
#endif
__device__ __forceinline__ float2 ffma2(const float2 a, const float2 b, const float2 c) {
#if G2V_PACKED_F32
  uint64_t ra, rb, rc, rd;
  float2 d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
#else
  return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}

// 16 columns of a chunk: key = fixed-point(e2 - 2 z.e) << 9 | column group; one running top-2 chain
// per column position
template <bool PARTIAL>
__device__ __forceinline__ void epi_chunk(const uint32_t (&v)[16], const float* __restrict__ e2c, float cS, float S,
                                          uint32_t group, int nvalid, uint32_t (&m1)[16], uint32_t (&m2)[16]) {
#pragma unroll
  for (int j4 = 0; j4 < 4; ++j4) {
    const float4 e = *reinterpret_cast<const float4*>(e2c + 4 * j4);
    const float2 cS2 = make_float2(cS, cS), S2 = make_float2(S, S), M2 = make_float2(kMagic, kMagic);
    const float2 x01 = ffma2(make_float2(__uint_as_float(v[4 * j4]), __uint_as_float(v[4 * j4 + 1])), cS2,
                             ffma2(make_float2(e.x, e.y), S2, M2));
    const float2 x23 = ffma2(make_float2(__uint_as_float(v[4 * j4 + 2]), __uint_as_float(v[4 * j4 + 3])), cS2,
                             ffma2(make_float2(e.z, e.w), S2, M2));
    const float xx[4] = {x01.x, x01.y, x23.x, x23.y};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = 4 * j4 + u;
      float x = fminf(xx[u], kKeyCap);               // far codes saturate instead of wrapping
      uint32_t key = __float_as_uint(x) * 512u + group;
      if (PARTIAL && j >= nvalid) key = 0xFFFFFFFFu;
      m2[j] = max(m1[j], min(m2[j], key));
      m1[j] = min(m1[j], key);
    }
  }
}

__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
// volatile: the load stays where it is written (half a piece ahead of its use) instead of being sunk to it
__device__ __forceinline__ float4 lds4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
// 8 columns (chains OFF .. OFF+7 of this thread): keys as in epi_chunk.  Written stage by stage over the eight
// independent columns: ptxas keeps that order, and every stage then has eight instructions in flight instead
// of one dependent chain per column (which left the epilogue warps latency-bound at ~0.17 IPC).
template <int OFF, bool PARTIAL>
__device__ __forceinline__ void epi8(const uint32_t (&v)[8], const float4 e0, const float4 e1, float cS, float S, uint32_t group,
                                     int nvalid, uint32_t (&m1)[16], uint32_t (&m2)[16]) {
  const float2 cS2 = make_float2(cS, cS), S2 = make_float2(S, S), M2 = make_float2(kMagic, kMagic);
  const float2 ee[4] = {make_float2(e0.x, e0.y), make_float2(e0.z, e0.w), make_float2(e1.x, e1.y), make_float2(e1.z, e1.w)};
  float2 t[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) t[p] = ffma2(ee[p], S2, M2);
#pragma unroll
  for (int p = 0; p < 4; ++p)
    t[p] = ffma2(make_float2(__uint_as_float(v[2 * p]), __uint_as_float(v[2 * p + 1])), cS2, t[p]);
  float f[8];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    f[2 * p] = fminf(t[p].x, kKeyCap);                  // far codes saturate instead of wrapping
    f[2 * p + 1] = fminf(t[p].y, kKeyCap);
  }
  uint32_t key[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    key[j] = __float_as_uint(f[j]) * 512u + group;
    if (PARTIAL && j >= nvalid) key[j] = 0xFFFFFFFFu;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) lo[j] = min(m2[OFF + j], key[j]);
#pragma unroll
  for (int j = 0; j < 8; ++j) m2[OFF + j] = max(m1[OFF + j], lo[j]);
#pragma unroll
  for (int j = 0; j < 8; ++j) m1[OFF + j] = min(m1[OFF + j], key[j]);
}
// One code tile of one epilogue warp: its 16-column pieces g0 .. g1-1 (codes [32 g + 16 eh, +16)), software-pipelined
// in halves of 8 columns -- while one half is turned into keys, the tensor-memory load and the ||e||^2 loads
// of the next half are in flight.  tcol0 / e2addr0 address code 16 eh (tensor-memory column, shared-memory byte).
// The key's column-group field is group0 + g.
__device__ __forceinline__ void epi_sweep(uint32_t tcol0, uint32_t e2addr0, int g0, int g1, int e_valid, int eh, float cS, float S,
                                          uint32_t group0, bool skip, uint32_t (&m1)[16], uint32_t (&m2)[16]) {
  if (g0 >= g1) return;
  uint32_t va[8], vb[8];
  tc_ld8(tcol0 + 32u * g0, va);
  float4 ea0 = lds4(e2addr0 + 128u * g0), ea1 = lds4(e2addr0 + 128u * g0 + 16u), eb0, eb1;
  for (int g = g0; g < g1; ++g) {
    const int nv = e_valid - (32 * g + 16 * eh);          // valid codes of this piece (>= 1)
    tc_wait_ld();
    tc_ld8(tcol0 + 32u * g + 8u, vb);
    eb0 = lds4(e2addr0 + 128u * g + 32u);
    eb1 = lds4(e2addr0 + 128u * g + 48u);
    if (!skip) {
      if (nv >= 8) epi8<0, false>(va, ea0, ea1, cS, S, group0 + (uint32_t)g, 8, m1, m2);
      else epi8<0, true>(va, ea0, ea1, cS, S, group0 + (uint32_t)g, nv, m1, m2);
    }
    tc_wait_ld();
    if (g + 1 < g1) {
      tc_ld8(tcol0 + 32u * (g + 1), va);
      ea0 = lds4(e2addr0 + 128u * (g + 1));
      ea1 = lds4(e2addr0 + 128u * (g + 1) + 16u);
    }
    if (!skip) {
      if (nv >= 16) epi8<8, false>(vb, eb0, eb1, cS, S, group0 + (uint32_t)g, 8, m1, m2);
      else epi8<8, true>(vb, eb0, eb1, cS, S, group0 + (uint32_t)g, nv - 8, m1, m2);
    }
  }
}

struct Cand {
  uint32_t key;    // best key of a chain
  uint32_t key2;   // second-best key of the same chain
  int j;           // chain id = column % 32
};
// keep the three smallest chain minima (c1 <= c2 <= c3) and the fourth smallest key (k4)
// (branch-free: rows of a warp take different paths, and the selects of successive insertions overlap)
__device__ __forceinline__ void cand_insert(Cand& c1, Cand& c2, Cand& c3, uint32_t& k4, const Cand n) {
  const bool l1 = n.key < c1.key, l2 = n.key < c2.key, l3 = n.key < c3.key;
  k4 = min(k4, max(n.key, c3.key));      // whichever of the new key and the old third does not stay among the three
  c3.key = l3 ? (l2 ? c2.key : n.key) : c3.key;
  c3.key2 = l3 ? (l2 ? c2.key2 : n.key2) : c3.key2;
  c3.j = l3 ? (l2 ? c2.j : n.j) : c3.j;
  c2.key = l2 ? (l1 ? c1.key : n.key) : c2.key;
  c2.key2 = l2 ? (l1 ? c1.key2 : n.key2) : c2.key2;
  c2.j = l2 ? (l1 ? c1.j : n.j) : c2.j;
  c1.key = l1 ? n.key : c1.key;
  c1.key2 = l1 ? n.key2 : c1.key2;
  c1.j = l1 ? n.j : c1.j;
}

// Per-row decision from the three best chain minima (each with the runner-up of its own chain) and the
// fourth chain minimum: certify the winner, or list the row for an exact re-rank of the codes that can
// still win (candidate list / one chain / the whole row).
__device__ __forceinline__ void finish_row(const Cand& c1, const Cand& c2, const Cand& c3, uint32_t k4, const RowInfo& ri,
                                           long long row, bool valid, int K, const float* __restrict__ ntab, unsigned flags,
                                           int* __restrict__ idx, int* pair_list, int* chain_list, int* full_list,
                                           int* counters) {
  const uint32_t t1 = c1.key;
  const int code1 = (int)(t1 & 511u) * 32 + c1.j;
  if (!valid) return;
  const uint32_t v1 = t1 >> 9;
  // certification threshold: per-code error at the largest norm that can still win this row
  // (twice for the distance, twice for a gap of two codes), the fp32 rounding of e2, and the
  // fixed-point rounding of both keys
  const float d1 = ri.z2 + ((float)(int)v1 - 4194304.f) / ri.S;
  const float cu = reachable_norm(ntab, ri.znorm, d1);
  const float tauf = (4.f * (ri.a1 * cu + ri.a0) + 2.3841858e-7f * cu * cu) * ri.S + 4.f;
  const uint32_t tau = (uint32_t)fminf(tauf, 4194304.f);
  auto near = [&](uint32_t key) { return (key >> 9) - v1 <= tau; };   // keys are >= t1
  if (flags & G2V_LIST_ALL_ROWS) {          // test aid: every row goes through the whole-row path (refine pass / fp64)
    full_list[atomicAdd(counters + 1, 1)] = (int)row;
  } else if (!(flags & G2V_NO_RECHECK) && (near(c2.key) || near(c1.key2))) {
    // Exact candidates are known only if no chain hides a code: a chain's third-best is unknown,
    // so a runner-up within tau (of any of the three chains) or a fourth chain within tau means
    // the whole row must be re-ranked.  Otherwise the candidates are the chain minima within tau.
    const bool h1 = near(c1.key2), h2 = near(c2.key2), h3 = near(c3.key2), h4 = near(k4);
    const int code2 = (int)(c2.key & 511u) * 32 + c2.j;
    const int code3 = (int)(c3.key & 511u) * 32 + c3.j;
    const bool n2 = near(c2.key) && code2 < K, n3 = near(c3.key) && code3 < K;
    if (!h1 && !h2 && !h3 && !h4) {
      // every candidate is known: the chain minima within tau
      const int slot = atomicAdd(counters + 0, 1);
      reinterpret_cast<int4*>(pair_list)[slot] = make_int4((int)row, code1, n2 ? code2 : code1, n3 ? code3 : -1);
    } else if (!h4 && ((int)h1 + (int)h2 + (int)h3) == 1) {
      // exactly one chain may hide codes: re-rank that chain's K/32 codes plus the other chains' minima
      const int slot = atomicAdd(counters + 2, 1);
      const int cj = h1 ? c1.j : (h2 ? c2.j : c3.j);
      const int ea = h1 ? (n2 ? code2 : -1) : code1;
      const int eb = h3 ? (n2 ? code2 : -1) : (n3 ? code3 : -1);
      reinterpret_cast<int4*>(chain_list)[slot] = make_int4((int)row, cj, ea, eb);
    } else {
      const int slot = atomicAdd(counters + 1, 1);
      full_list[slot] = (int)row;
    }
  }
  idx[row] = min(code1, K - 1);
}

// Per-row constants from ||z||^2 and the squared norm of the fp16 rounding residual of the row.
//   |dot_hat - dot| for a code of norm c:
//     operand rounding (Cauchy-Schwarz on the measured residuals; ||r_e|| <= sfrac c + rsub):
//          |r_z| c + (|z| + |r_z|) (sfrac c + rsub)
//     tensor-core accumulation: (2^-19 alignment + one fp32 rounding per K step) |z| c
//   = a1 c + a0.  The epilogue evaluates it at the largest code norm that can still win the row.
//   Fixed-point range: keys of codes that can still win lie in [-|z|^2, (|z| + c_min)^2]; anything
//   larger saturates in the epilogue.
__device__ __forceinline__ RowInfo make_rowinfo(float z2, float r2, float inv_scale_z, float sfrac, float rsub, float e2min,
                                                float scale_e, int n_ksteps) {
  const float znorm = sqrtf(z2) * 1.0001f, rnorm = sqrtf(r2) * 1.0001f;
  const float acc = 1.9073486e-6f + (float)(n_ksteps + 2) * 1.1920929e-7f;
  RowInfo ri;
  ri.a1 = rnorm * (1.f + sfrac) + znorm * (sfrac + acc);
  ri.a0 = (znorm + rnorm) * rsub;
  ri.znorm = znorm;
  ri.z2 = z2;
  const float cmin = sqrtf(e2min);
  const float R = 2.f * (znorm + cmin) * (znorm + cmin) + 1e-30f;
  float S = 1.f;
  if (isfinite(R)) {
    int e;
    frexpf(R, &e);                                    // R < 2^e
    S = ldexpf(1.f, max(min(21 - e, 100), -100));     // R * S < 2^21
  }
  ri.S = S;
  ri.cS = (-2.f * inv_scale_z / scale_e) * S;
  ri.pad0 = ri.pad1 = 0.f;
  return ri;
}

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
template <int CG, bool FUSED>
__global__ void __launch_bounds__(FUSED ? NTHREADS_FUSED : NTHREADS, 1)
tc_search_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAt,
                 const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBt,
                 const __grid_constant__ CUtensorMap tmBl, const __grid_constant__ CUtensorMap tmBlt,
                 const __grid_constant__ CUtensorMap tmPf, const TcParams P) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* gbase = smem_dyn + (base - raw);
  const SmemPlan sp = smem_plan(P.n_full, P.n_tail, CG, P.nstage, FUSED ? P.nzslot : 0);
  const uint32_t NST = (uint32_t)P.nstage;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int n_groups = gridDim.x / CG, group = blockIdx.x / CG;   // CTA groups walk the row tiles

  const uint32_t sA = base + sp.a_off, sB = base + sp.b_off;
  float* e2s = reinterpret_cast<float*>(gbase + sp.e2_off);
  uint32_t* xch = reinterpret_cast<uint32_t*>(gbase + sp.xch_off);
  const uint32_t bars = base + sp.bar_off;
  // barrier map
  auto bar_full = [&](int s) { return bars + 8u * s; };
  auto bar_empty = [&](int s) { return bars + 8u * (MAX_STAGES + s); };
  auto bar_afull = [&](int c) { return bars + 8u * (2 * MAX_STAGES + c); };
  auto bar_aempty = [&](int c) { return bars + 8u * (2 * MAX_STAGES + MAX_CHUNKS + c); };
  auto bar_accfull = [&](int a) { return bars + 8u * (2 * MAX_STAGES + 2 * MAX_CHUNKS + a); };
  auto bar_accempty = [&](int a) { return bars + 8u * (2 * MAX_STAGES + 2 * MAX_CHUNKS + 2 + a); };
  auto bar_zfull = [&](int s) { return bars + 8u * (2 * MAX_STAGES + 2 * MAX_CHUNKS + 4 + s); };
  auto bar_zempty = [&](int s) { return bars + 8u * (2 * MAX_STAGES + 2 * MAX_CHUNKS + 4 + MAX_ZSLOTS + s); };
  auto bar_rsfull = [&](int s) { return bars + 8u * (2 * MAX_STAGES + 2 * MAX_CHUNKS + 4 + 2 * MAX_ZSLOTS + s); };
  const uint32_t sZ = base + sp.zst_off;
  float2* rowstat = reinterpret_cast<float2*>(gbase + sp.rs_off);      // [RS_RING][TM]: (||z||^2, ||z - z16||^2)
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + sp.tmem_off);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_chunks = P.n_chunks, n_full = P.n_full;

  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
    // fused: the A panels are written by 4 converter warps in each CTA of the group instead of by TMA
    for (int c = 0; c < MAX_CHUNKS; ++c) { mbar_init(bar_afull(c), FUSED ? CG : 1); mbar_init(bar_aempty(c), 1); }
    for (int s = 0; s < MAX_ZSLOTS; ++s) { mbar_init(bar_zfull(s), 1); mbar_init(bar_zempty(s), 4); }
    for (int s = 0; s < RS_RING; ++s) mbar_init(bar_rsfull(s), 4);
    for (int a = 0; a < 2; ++a) { mbar_init(bar_accfull(a), 1); mbar_init(bar_accempty(a), 8 * CG); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if constexpr (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                   "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                   "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if constexpr (CG == 1) __syncthreads();
  else cluster_sync_all();       // the peer's barriers must be initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // bring-up trace: role r (0 B-prod, 1 MMA, 2 A-prod, 3 epilogue warp 4, 4 converter warp 12), up to 512 events each
  int trace_n = 0;
  auto TRACE = [&](int role, int tag) {
    if (kTraceBuild && P.trace && blockIdx.x == 0 && lane == 0 && trace_n < 512) {
      long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      P.trace[(role * 512 + trace_n) * 2] = tag;
      P.trace[(role * 512 + trace_n) * 2 + 1] = t;
      ++trace_n;
    }
  };
  // chunk geometry: chunk c < n_full is a 64-wide SW128 panel, otherwise a 16-wide SW32 tail
  auto a_chunk_addr = [&](int c) { return c < n_full ? sA + (uint32_t)c * A_PANEL : sA + (uint32_t)n_full * A_PANEL + (uint32_t)(c - n_full) * A_TAIL; };
  auto chunk_col = [&](int c) { return c < n_full ? c * KC : n_full * KC + (c - n_full) * KT; };

  if (warp == 0) {
    // =========================== B producer ===========================
    {   // whole warp walks the schedule (uniform control flow); one elected lane issues
      uint32_t s = 0, sph = 0;  // ring position and its phase (no division by the run-time ring depth in the loop)
      for (int tile = group; tile < P.n_row_tiles; tile += n_groups) {
        for (int nt = 0; nt < P.n_ntiles; ++nt) {
          for (int c = 0; c < n_chunks; ++c) {
            mbar_wait(bar_empty(s), sph ^ 1u);
            const bool full = c < n_full;
            // the last code tile only fetches the rows its MMA reads (n_last_mma <= 256); each CTA
            // of a pair fetches its 1/CG slice of the code rows
            const bool last = (nt == P.n_ntiles - 1);
            const uint32_t rows = last ? (uint32_t)P.n_last_mma : (uint32_t)TN;
            const CUtensorMap* tm = last ? (full ? &tmBl : &tmBlt) : (full ? &tmB : &tmBt);
            if (P.flags & kDbgNoB) {
              if (elect_one() && leader) mbar_arrive(bar_full(s));
            } else if (elect_one()) {
              if (leader) mbar_expect_tx(bar_full(s), rows * (full ? KC : KT) * 2u);   // bytes of all CTAs
              tma_load_2d<CG>(sB + (uint32_t)s * sp.b_stage, tm, chunk_col(c), nt * TN + (int)(cta_rank * (rows / CG)),
                              bar_full(s));
            }
            __syncwarp();
            if (++s == NST) { s = 0; sph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 2) {
    // =========================== A producer ===========================
    if constexpr (FUSED) {
      // fp32 rows, 16 columns at a time, into the staging ring (tmA is the fp32 map here)
      // per row tile: for every full 64-column panel two half-tiles of 64 rows (tmA, 256-byte rows), then
      // the 16-column tails for all 128 rows (tmAt, fp32)
      // The ring holds only ~48 KB, far less than HBM latency x per-SM bandwidth, so every row tile
      // is first pulled into L2 (TMA prefetch, 128 rows x 64 columns per request) kPfTiles tiles ahead.
      uint32_t slot = 0, zphase = 0;
      const bool use_pf = (P.flags & kDbgPrefetch) != 0;   // measured: no gain (the stream is not DRAM-latency-bound)
      auto prefetch_tile = [&](int t) {
        if (use_pf && t < P.n_row_tiles && elect_one())
          for (int c0 = 0; c0 < P.D; c0 += KC) tma_prefetch_2d(&tmPf, c0, (t * CG + (int)cta_rank) * TM);
        __syncwarp();
      };
      for (int a = 0; a < kPfTiles; ++a) prefetch_tile(group + a * n_groups);
      for (int tile = group; tile < P.n_row_tiles; tile += n_groups) {
        prefetch_tile(tile + kPfTiles * n_groups);
        const int row0 = (tile * CG + (int)cta_rank) * TM;
        for (int u = 0; u < 2 * n_full + P.n_tail; ++u) {
          mbar_wait(bar_zempty(slot), zphase ^ 1u);
          TRACE(2, u);
          if (P.flags & kDbgNoZ) {
            if (elect_one()) mbar_arrive(bar_zfull(slot));
          } else if (elect_one()) {
            if (u < 2 * n_full) {
              mbar_expect_tx(bar_zfull(slot), ZSLOT);
              tma_load_2d<1>(sZ + slot * ZSLOT, &tmA, (u >> 1) * KC, row0 + (u & 1) * ZROWS, bar_zfull(slot));
            } else {
              mbar_expect_tx(bar_zfull(slot), TM * KT * 4);
              tma_load_2d<1>(sZ + slot * ZSLOT, &tmAt, n_full * KC + (u - 2 * n_full) * KT, row0, bar_zfull(slot));
            }
          }
          __syncwarp();
          if (++slot == (uint32_t)P.nzslot) { slot = 0; zphase ^= 1u; }
        }
      }
    } else {
      uint32_t ti = 0;
      for (int tile = group; tile < P.n_row_tiles; tile += n_groups, ++ti) {
        for (int c = 0; c < n_chunks; ++c) {
          mbar_wait(bar_aempty(c), (ti & 1u) ^ 1u);
          const bool full = c < n_full;
          if (elect_one()) {
            if (leader) mbar_expect_tx(bar_afull(c), (uint32_t)CG * (full ? A_PANEL : A_TAIL));
            tma_load_2d<CG>(a_chunk_addr(c), full ? &tmA : &tmAt, chunk_col(c), (tile * CG + (int)cta_rank) * TM,
                            bar_afull(c));
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (leader) {                  // the leader CTA issues for the whole group; one elected lane per step
      uint32_t s = 0, sph = 0, it = 0, ti = 0;
      for (int tile = group; tile < P.n_row_tiles; tile += n_groups, ++ti) {
        for (int nt = 0; nt < P.n_ntiles; ++nt, ++it) {
          const uint32_t as = it & 1u, around = it >> 1;
          const bool last_nt = (nt == P.n_ntiles - 1);
          const uint32_t idesc = umma_idesc(TM * CG, last_nt ? P.n_last_mma : TN);
          TRACE(1, 100 + nt);
          mbar_wait(bar_accempty(as), (around & 1u) ^ 1u);
          TRACE(1, 200 + nt);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + as * TN;
          for (int c = 0; c < n_chunks; ++c) {
            if (nt == 0) {
              if (FUSED && CG == 2) mbar_wait_cluster(bar_afull(c), ti & 1u);
              else mbar_wait(bar_afull(c), ti & 1u);
              if (FUSED) fence_proxy_async();
            }
            mbar_wait(bar_full(s), sph);
            tc_fence_after();
            const uint32_t a_addr = a_chunk_addr(c), b_addr = sB + (uint32_t)s * sp.b_stage;
            const bool fullp = c < n_full;
            const uint64_t ad0 = fullp ? umma_desc(a_addr, 1024, 2) : umma_desc(a_addr, 256, 6);
            const uint64_t bd0 = fullp ? umma_desc(b_addr, 1024, 2) : umma_desc(b_addr, 256, 6);
            if (elect_one()) {
              if (P.flags & kDbgSkipMma) {
                // bring-up aid: measure everything but the tensor work
              } else if (fullp) {
#pragma unroll
                for (int k = 0; k < KC / KT; ++k)        // +32 bytes per K step = +2 in the address field
                  tc_mma_f16<CG>(d_tmem, ad0 + 2u * k, bd0 + 2u * k, idesc, (c | k) != 0);
              } else {
                tc_mma_f16<CG>(d_tmem, ad0, bd0, idesc, c != 0);
              }
              tc_commit<CG>(bar_empty(s));                 // B stage free once these MMAs retire
              if (last_nt) tc_commit<CG>(bar_aempty(c));   // A panel free after its last use in this row tile
              if (c == n_chunks - 1) tc_commit<CG>(bar_accfull(as));
            }
            __syncwarp();
            if (++s == NST) { s = 0; sph ^= 1u; }
          }
        }
      }
    }
  } else if (warp >= EPI_WARP0 && warp < CONV_WARP0) {
    // =========================== epilogue ===========================
    // 8 warps: two per TMEM lane quarter, each taking 16 of every 32 columns (so every scheduler
    // has two epilogue warps to interleave).  Thread = one row; 16 running top-2 chains per thread.
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int eh = (warp - EPI_WARP0) >> 2;          // which 16-column half of each 32-column chunk
    const int r = q * 32 + lane;                     // row within the tile == TMEM lane
    const int et = threadIdx.x - EPI_WARP0 * 32;     // 0..255
    uint32_t* xrow = xch + (size_t)r * 12;           // hand-off slot of this row (half 1 -> half 0)
    uint32_t it = 0, ti = 0;
    const float h_sfrac = FUSED ? P.hdr->sfrac : 0.f, h_rsub = FUSED ? P.hdr->rsub : 0.f, h_e2min = FUSED ? P.hdr->e2min : 0.f,
                h_scale_e = FUSED ? P.hdr->scale_e : 1.f;
    for (int tile = group; tile < P.n_row_tiles; tile += n_groups, ++ti) {
      const long long row = ((long long)tile * CG + cta_rank) * TM + r;
      const bool valid = row < P.N;
      RowInfo ri;
      if constexpr (FUSED) {
        // statistics of this tile's rows come from the converter warps (same CTA)
        mbar_wait(bar_rsfull(ti % RS_RING), (ti / RS_RING) & 1u);
        const float2 st = rowstat[(ti % RS_RING) * TM + r];
        ri = make_rowinfo(st.x, st.y, 1.f, h_sfrac, h_rsub, h_e2min, h_scale_e, P.n_ksteps);
      } else {
        ri = valid ? P.rowinfo[row] : RowInfo{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      }
      uint32_t m1[16], m2[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) { m1[j] = 0xFFFFFFFFu; m2[j] = 0xFFFFFFFFu; }

      for (int nt = 0; nt < P.n_ntiles; ++nt, ++it) {
        const uint32_t as = it & 1u, around = it >> 1;
        float* e2c = e2s + as * TN;
        {  // stage this code tile's ||e||^2 (one per thread)
          const int k0 = nt * TN + et;
          e2c[et] = (k0 < P.K) ? __ldg(P.e2 + k0) : 0.f;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (warp == EPI_WARP0) TRACE(3, 100 + nt);
        mbar_wait(bar_accfull(as), around & 1u);
        if (warp == EPI_WARP0) TRACE(3, 200 + nt);
        tc_fence_after();
        const int ncols = min(TN, P.K - nt * TN);
        // pieces of 16 columns at code 32 ch + 16 eh of this tile; ||e||^2 of the tile staged in e2c
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * TN + eh * 16;
        epi_sweep(taddr, smem_u32(e2c) + 64u * (uint32_t)eh, 0, (ncols - 16 * eh + 31) >> 5, ncols, eh, ri.cS, ri.S,
                  (uint32_t)(nt * 8), (P.flags & kDbgSkipEpi) != 0, m1, m2);
        // all TMEM reads of this stage are complete (last wait::ld above): hand it back
        tc_fence_before();
        __syncwarp();
        if (warp == EPI_WARP0) TRACE(3, 300 + nt);
        if (lane == 0) {
          if (CG == 1 || leader) mbar_arrive(bar_accempty(as));
          else mbar_arrive_cluster(bar_accempty(as), 0);      // the leader's MMA thread waits for both CTAs
        }
      }

      // ---- per-row decision ----
      // the three best chain minima (each with the runner-up of its own chain) and the fourth chain
      // minimum: first over this thread's 16 chains, then merged with the other half's
      const Cand none{0xFFFFFFFFu, 0xFFFFFFFFu, 0};
      Cand c1 = none, c2 = none, c3 = none;
      uint32_t k4 = 0xFFFFFFFFu;
#pragma unroll
      for (int j = 0; j < 16; ++j) cand_insert(c1, c2, c3, k4, Cand{m1[j], m2[j], eh * 16 + j});
      if (eh == 1) {
        xrow[0] = c1.key; xrow[1] = c1.key2; xrow[2] = (uint32_t)c1.j;
        xrow[3] = c2.key; xrow[4] = c2.key2; xrow[5] = (uint32_t)c2.j;
        xrow[6] = c3.key; xrow[7] = c3.key2; xrow[8] = (uint32_t)c3.j;
        xrow[9] = k4;
      }
      asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");     // the two warps of this lane quarter
      if (eh == 0) {
        cand_insert(c1, c2, c3, k4, Cand{xrow[0], xrow[1], (int)xrow[2]});
        cand_insert(c1, c2, c3, k4, Cand{xrow[3], xrow[4], (int)xrow[5]});
        cand_insert(c1, c2, c3, k4, Cand{xrow[6], xrow[7], (int)xrow[8]});
        k4 = min(k4, xrow[9]);
        finish_row(c1, c2, c3, k4, ri, row, valid, P.K, P.ntab, P.flags, P.idx, P.pair_list, P.chain_list, P.full_list, P.counters);
      }
      asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");     // slot may be rewritten by the next tile
    }
  }

  if constexpr (FUSED) {
    if (warp >= CONV_WARP0) {
      // =========================== converters ===========================
      // fp32 staging slot (128 rows x 16 columns) -> fp16 into the swizzled A panels; lane = (row % 8,
      // float4 of the row), so a warp reads 512 contiguous bytes and its 8-byte stores hit 8 rows whose
      // 16-byte chunks the swizzle spreads over all banks.
      // lane = (which of 2 rows, which float4 of the 64-column row): a warp reads 2 x 256 contiguous
      // bytes per step and writes two 128-byte panel rows whose 16-byte chunks the swizzle permutes.
      // Each lane owns rows {h*64 + cw*16 + 2i + rsel} and keeps their |z|^2 / |z - z16|^2 in registers.
      const int cw = warp - CONV_WARP0;
      const int q = lane & 15, rsel = lane >> 4;
      // per-lane constants: row of step i inside a half is r_i = cw*16 + 2i + rsel, and (h*64 + r_i) & 7
      // == (2i + rsel) & 7, so the swizzled 16-byte chunk of this lane depends on i only
      const uint32_t src_lane = (uint32_t)((cw * 16 + rsel) * (KC * 4) + q * 16);          // + i * 2 rows
      const uint32_t dst_lane = (uint32_t)((cw * 16 + rsel) * 128 + (q & 1) * 8);          // + i * 2 rows + swizzle
      uint32_t swz[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) swz[i] = ((((uint32_t)q >> 1) ^ (uint32_t)((2 * i + rsel) & 7)) << 4) + (uint32_t)i * 256u;
      uint32_t slot = 0, zphase = 0, ti = 0;
      for (int tile = group; tile < P.n_row_tiles; tile += n_groups, ++ti) {
        float z2[16], r2[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { z2[i] = 0.f; r2[i] = 0.f; }
        // ---- full 64-column panels: two 64-row slots each ----
        for (int c = 0; c < n_full; ++c) {
          if (cw == 0) TRACE(4, 100 + c);
          mbar_wait(bar_aempty(c), (ti & 1u) ^ 1u);                     // MMA finished with the previous tile's panel
          if (cw == 0) TRACE(4, 200 + c);
          const uint32_t abase = sA + (uint32_t)c * A_PANEL + dst_lane;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            mbar_wait(bar_zfull(slot), zphase);
            if (cw == 0) TRACE(4, 300 + c * 2 + h);
            const unsigned char* zsrc = gbase + sp.zst_off + slot * ZSLOT + src_lane;
            float4 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const float4*>(zsrc + i * 2 * (KC * 4));
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_zempty(slot));               // slot is in registers
            if (++slot == (uint32_t)P.nzslot) { slot = 0; zphase ^= 1u; }
            if (!(P.flags & kDbgSkipConv)) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const __half2 h01 = __floats2half2_rn(v[i].x, v[i].y), h23 = __floats2half2_rn(v[i].z, v[i].w);
                const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                const float e0 = v[i].x - f01.x, e1 = v[i].y - f01.y, e2r = v[i].z - f23.x, e3 = v[i].w - f23.y;
                z2[h * 8 + i] = fmaf(v[i].x, v[i].x, fmaf(v[i].y, v[i].y, fmaf(v[i].z, v[i].z, fmaf(v[i].w, v[i].w, z2[h * 8 + i]))));
                r2[h * 8 + i] = fmaf(e0, e0, fmaf(e1, e1, fmaf(e2r, e2r, fmaf(e3, e3, r2[h * 8 + i]))));
                asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(abase + (uint32_t)h * (64u * 128u) + swz[i]),
                             "r"(*reinterpret_cast<const uint32_t*>(&h01)), "r"(*reinterpret_cast<const uint32_t*>(&h23))
                             : "memory");
              }
            }
          }
          if (cw == 0) TRACE(4, 400 + c);
          fence_proxy_async();                                          // my stores -> visible to the MMA's reads
          if (cw == 0) TRACE(4, 500 + c);
          asm volatile("bar.sync 7, 128;" ::: "memory");                // all four converter warps
          if (cw == 0) TRACE(4, 600 + c);
          if (threadIdx.x == CONV_WARP0 * 32) {                         // one arrive per CTA and panel
            if (CG == 1 || leader) mbar_arrive(bar_afull(c));
            else mbar_arrive_cluster(bar_afull(c), 0);
          }
        }
        // ---- 16-column tails: one slot holds all 128 rows; only lanes q < 4 carry data ----
        for (int c = n_full; c < n_chunks; ++c) {
          mbar_wait(bar_aempty(c), (ti & 1u) ^ 1u);
          mbar_wait(bar_zfull(slot), zphase);
          const unsigned char* zsrc = gbase + sp.zst_off + slot * ZSLOT;
          const uint32_t abase = a_chunk_addr(c);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint32_t rr = (uint32_t)(h * 64 + cw * 16 + 2 * i + rsel);
              if (q < 4) {
                const float4 v = *reinterpret_cast<const float4*>(zsrc + rr * (KT * 4) + q * 16);
                const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
                const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                const float e0 = v.x - f01.x, e1 = v.y - f01.y, e2r = v.z - f23.x, e3 = v.w - f23.y;
                z2[h * 8 + i] = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, z2[h * 8 + i]))));
                r2[h * 8 + i] = fmaf(e0, e0, fmaf(e1, e1, fmaf(e2r, e2r, fmaf(e3, e3, r2[h * 8 + i]))));
                const uint32_t dst = abase + rr * 32u + (((((uint32_t)q >> 1) & 1u) ^ ((rr >> 2) & 1u)) << 4) + ((uint32_t)q & 1u) * 8u;
                asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst), "r"(*reinterpret_cast<const uint32_t*>(&h01)),
                             "r"(*reinterpret_cast<const uint32_t*>(&h23))
                             : "memory");
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_zempty(slot));
          if (++slot == (uint32_t)P.nzslot) { slot = 0; zphase ^= 1u; }
          if (c == n_chunks - 1) break;                                 // statistics first (below), then publish
          fence_proxy_async();
          asm volatile("bar.sync 7, 128;" ::: "memory");
          if (threadIdx.x == CONV_WARP0 * 32) {
            if (CG == 1 || leader) mbar_arrive(bar_afull(c));
            else mbar_arrive_cluster(bar_afull(c), 0);
          }
        }
        // row statistics of the finished tile, before its last panel is published
        {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float a = z2[i], b = r2[i];
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) {
              a += __shfl_xor_sync(0xffffffffu, a, o);
              b += __shfl_xor_sync(0xffffffffu, b, o);
            }
            if (q == 0) rowstat[(ti % RS_RING) * TM + (i >> 3) * 64 + cw * 16 + 2 * (i & 7) + rsel] = make_float2(a, b);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_rsfull(ti % RS_RING));
        }
        if (P.n_tail > 0) {                                             // publish the last tail panel
          fence_proxy_async();
          asm volatile("bar.sync 7, 128;" ::: "memory");
          if (threadIdx.x == CONV_WARP0 * 32) {
            if (CG == 1 || leader) mbar_arrive(bar_afull(n_chunks - 1));
            else mbar_arrive_cluster(bar_afull(n_chunks - 1), 0);
          }
        }
      }
    }
  }

  tc_fence_before();
  if constexpr (CG == 1) __syncthreads();
  else cluster_sync_all();       // neither CTA may leave while the other still reads its smem / TMEM
  if (warp == 2) {
    if constexpr (CG == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// tensor-memory helpers of tc_tmem_kernel (A operand written by tcgen05.st, read by tcgen05.mma)
// ------------------------------------------------------------------------------------------
constexpr int TME_CONV_WARP0 = 12;
constexpr int TME_CONV_WARPS = 8;

__device__ __forceinline__ void tc_mma_f16_ts2(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_st_16x256b_x4(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tc_st_16x256b_x1(uint32_t taddr, uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3) {
  asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v0), "r"(v1), "r"(v2), "r"(v3)
               : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// ------------------------------------------------------------------------------------------
// "tmem" variant: fp32 rows -> fp16 A operand in tensor memory, CTA pairs, both operands streamed
// ------------------------------------------------------------------------------------------
// What bounds the fused shared-memory variant at K <= 512 is shared-memory bandwidth and the depth of
// its fp32 staging ring (the fp16 A tile takes 100 KB).  Here the A operand lives in tensor memory
// (tcgen05.st.16x256b, lane = row, 32-bit column = two K elements; tcgen05.mma reads it from there), so
// shared memory only holds rings: >= 128 KB of fp32 row slots in flight per SM (HBM latency under load is
// 1.5 - 3 us) and the codebook stages, of which each CTA of a pair (cta_group::2, M = 256) loads half.
// TMEM columns: [0, Dp/2) A operand, then two accumulator stages of `ntile` codes.
constexpr int TME_THREADS = 640;        // 4 control warps, 8 epilogue warps, 8 converter warps
constexpr int TME_MAX_ZSLOTS = 10;
constexpr int TME_ZCOLS = 32;           // fp32 columns per staging slot (128 bytes: one SWIZZLE_128B row)
constexpr int TME_ZSLOT = TM * TME_ZCOLS * 4;   // 16384
constexpr int TME_MAX_BST = 16;
constexpr int kTmeDefaultABufs = 1;   // see plan_tmem()

struct TmeParams {
  long long N;
  int K, D, Dp;
  int Dz;                         // columns of the rows in memory (<= D; the TMA zero-fills the rest of the operand)
  int n_full, n_tail, n_chunks;   // K panels of the A operand: 64 columns, then 16-column tails
  int ntile, n_ntiles, n_last;    // codes per accumulator stage; code tiles; codes of the last tile (multiple of 16)
  int a_bufs, acc_col0;           // A operand buffers in TMEM (1 or 2); first accumulator column in TMEM
  int ahead;                      // converters turn super-chunk 0 of the next tile into registers before the A buffer is free
  int n_row_tiles;                // tiles of 256 rows (128 per CTA)
  int n_ksteps;
  int Kpad;                       // (n_ntiles - 1) * ntile + n_last
  int nb, nz;                     // codebook stages / fp32 row slots
  uint32_t b_stage;               // bytes of one codebook stage in one CTA
  const CbHeader* hdr;
  const float* ntab;
  const float* e2;
  int* idx;
  int* pair_list;
  int* full_list;
  int* chain_list;
  int* counters;
  long long* trace;
  unsigned flags;
};

struct TmePlan {
  uint32_t b_off, z_off, e2_off, xch_off, rs_off, bar_off, tmem_off, total;
};
constexpr int TME_NBARS = 2 * TME_MAX_BST + 2 * TME_MAX_ZSLOTS + 6 * MAX_CHUNKS + 4 + RS_RING;
__host__ __device__ inline TmePlan tme_plan(int nb, uint32_t b_stage, int nz, int Kpad) {
  TmePlan p;
  p.b_off = 0;
  p.z_off = (uint32_t)nb * b_stage;
  p.e2_off = p.z_off + (uint32_t)nz * TME_ZSLOT;
  p.xch_off = p.e2_off + (((uint32_t)Kpad * 4u + 127u) & ~127u);
  p.rs_off = p.xch_off + TM * 12 * 4;
  p.bar_off = p.rs_off + RS_RING * TM * 8;
  p.tmem_off = p.bar_off + 8 * TME_NBARS;
  p.total = p.tmem_off + 16;
  return p;
}

// four consecutive 16-bit row elements (8 bytes of a staging slot) as fp32
template <typename ZT>
__device__ __forceinline__ float4 unpack16(const uint2 u) {
  if constexpr (sizeof(ZT) == 4) {
    return make_float4(0.f, 0.f, 0.f, 0.f);      // not used for fp32 rows
  } else if constexpr (std::is_same<ZT, __half>::value) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
  } else {
    return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u), __uint_as_float(u.y << 16),
                       __uint_as_float(u.y & 0xFFFF0000u));
  }
}

template <typename ZT>
__global__ void __launch_bounds__(TME_THREADS, 1)
tc_tmem_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmZt,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBt,
               const __grid_constant__ CUtensorMap tmBl, const __grid_constant__ CUtensorMap tmBlt, const TmeParams P) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* gbase = smem_dyn + (base - raw);
  const TmePlan sp = tme_plan(P.nb, P.b_stage, P.nz, P.Kpad);
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int n_groups = gridDim.x / 2, group = blockIdx.x / 2;
  const uint32_t NB = (uint32_t)P.nb, NZ = (uint32_t)P.nz;

  const uint32_t sB = base + sp.b_off, sZ = base + sp.z_off;
  float* e2s = reinterpret_cast<float*>(gbase + sp.e2_off);
  uint32_t* xch = reinterpret_cast<uint32_t*>(gbase + sp.xch_off);
  float2* rowstat = reinterpret_cast<float2*>(gbase + sp.rs_off);
  const uint32_t bars = base + sp.bar_off;
  auto bar_bfull = [&](int s) { return bars + 8u * s; };
  auto bar_bempty = [&](int s) { return bars + 8u * (TME_MAX_BST + s); };
  auto bar_zfull = [&](int s) { return bars + 8u * (2 * TME_MAX_BST + s); };
  auto bar_zempty = [&](int s) { return bars + 8u * (2 * TME_MAX_BST + TME_MAX_ZSLOTS + s); };
  constexpr int A0 = 2 * TME_MAX_BST + 2 * TME_MAX_ZSLOTS;
  // per A buffer b and super-chunk c
  auto bar_aconv = [&](int b, int c) { return bars + 8u * (A0 + b * MAX_CHUNKS + c); };              // leader: the converters of both CTAs wrote it
  auto bar_aempty = [&](int b, int c) { return bars + 8u * (A0 + (4 + b) * MAX_CHUNKS + c); };       // the MMAs finished reading it
  auto bar_accfull = [&](int a) { return bars + 8u * (A0 + 6 * MAX_CHUNKS + a); };
  auto bar_accempty = [&](int a) { return bars + 8u * (A0 + 6 * MAX_CHUNKS + 2 + a); };
  auto bar_rsfull = [&](int s) { return bars + 8u * (A0 + 6 * MAX_CHUNKS + 4 + s); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + sp.tmem_off);

  constexpr bool Z32 = sizeof(ZT) == 4;     // fp32 rows; otherwise bf16 / fp16 rows
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_chunks = P.n_chunks, n_full = P.n_full;
  auto acc_col_of = [&](uint32_t a) { return (uint32_t)P.acc_col0 + a * (uint32_t)P.ntile; };   // accumulator stage a
  const uint32_t a_cols = (uint32_t)P.Dp / 2, abm = (uint32_t)P.a_bufs - 1u, absh = (uint32_t)P.a_bufs >> 1;   // buffer = ti & abm, use = ti >> absh
  const int n_sc = (n_full + 1) >> 1;      // super-chunks (codebook stages / A hand-overs) per code tile

  if (threadIdx.x == 0) {
    for (int s = 0; s < TME_MAX_BST; ++s) { mbar_init(bar_bfull(s), 1); mbar_init(bar_bempty(s), 1); }
    for (int s = 0; s < TME_MAX_ZSLOTS; ++s) { mbar_init(bar_zfull(s), 1); mbar_init(bar_zempty(s), TME_CONV_WARPS); }
    for (int b = 0; b < 2; ++b)
      for (int c = 0; c < MAX_CHUNKS; ++c) { mbar_init(bar_aconv(b, c), 2 * TME_CONV_WARPS); mbar_init(bar_aempty(b, c), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(bar_accfull(a), 1); mbar_init(bar_accempty(a), 16); }
    for (int s = 0; s < RS_RING; ++s) mbar_init(bar_rsfull(s), TME_CONV_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < P.Kpad; i += TME_THREADS) e2s[i] = (i < P.K) ? __ldg(P.e2 + i) : 0.f;
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  int trace_n = 0;
  auto TRACE = [&](int role, int tag) {   // role 0 B producer, 1 MMA, 2 row producer, 3 epilogue warp 4, 4 converter warp 12
    if (kTraceBuild && P.trace && blockIdx.x == 0 && lane == 0 && trace_n < 512) {
      long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      P.trace[(role * 512 + trace_n) * 2] = tag;
      P.trace[(role * 512 + trace_n) * 2 + 1] = t;
      ++trace_n;
    }
  };

  if (warp == 0) {
    // =========================== codebook producer (each CTA fetches its half of every stage) ===========================
    // a stage = one "super-chunk" of a code tile: two 64-column panels, the last one also the 16-column tails
    uint32_t s = 0, sph = 0;     // ring position and its phase
    for (int tile = group; tile < P.n_row_tiles; tile += n_groups) {
      for (int nt = 0; nt < P.n_ntiles; ++nt) {
        const bool last = (nt == P.n_ntiles - 1);
        const uint32_t rows = (uint32_t)(last ? P.n_last : P.ntile);        // codes of the tile; this CTA fetches half
        const int row = nt * P.ntile + (int)(cta_rank * (rows / 2));
        for (int sc = 0; sc < n_sc; ++sc) {
          mbar_wait_idle(bar_bempty(s), sph ^ 1u);
          const int p0 = 2 * sc, np = min(2, n_full - p0);
          const int nt_here = (sc == n_sc - 1) ? P.n_tail : 0;
          if (elect_one()) {
            if (leader) mbar_expect_tx(bar_bfull(s), rows * (uint32_t)(np * KC + nt_here * KT) * 2u);   // bytes of both CTAs
            uint32_t dst = sB + s * P.b_stage;
            for (int k = 0; k < np; ++k, dst += (rows / 2) * 128u)
              tma_load_2d<2>(dst, last ? &tmBl : &tmB, (p0 + k) * KC, row, bar_bfull(s));
            for (int t = 0; t < nt_here; ++t, dst += (rows / 2) * 32u)
              tma_load_2d<2>(dst, last ? &tmBlt : &tmBt, n_full * KC + t * KT, row, bar_bfull(s));
          }
          __syncwarp();
          if (++s == NB) { s = 0; sph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // =========================== MMA issuer ===========================
      uint32_t s = 0, sph = 0, it = 0, ti = 0;
      const bool skip_mma = (P.flags & kDbgSkipMma) != 0;
      for (int tile = group; tile < P.n_row_tiles; tile += n_groups, ++ti) {
        const int ab = (int)(ti & abm);
        const uint32_t apar = (ti >> absh) & 1u;
        const uint32_t a_base = tmem_base + (uint32_t)ab * a_cols;
        for (int nt = 0; nt < P.n_ntiles; ++nt, ++it) {
          const uint32_t as = it & 1u, around = it >> 1;
          const bool last_nt = (nt == P.n_ntiles - 1);
          const uint32_t idesc = umma_idesc(2 * TM, last_nt ? P.n_last : P.ntile);
          TRACE(1, 100 + nt);
          mbar_wait(bar_accempty(as), (around & 1u) ^ 1u);
          TRACE(1, 200 + nt);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc_col_of(as);
          const uint32_t rows_b = (uint32_t)(last_nt ? P.n_last : P.ntile) / 2u;     // code rows per CTA in a stage
          for (int sc = 0; sc < n_sc; ++sc) {
            if (nt == 0) {
              // the converter warps of BOTH CTAs arrive here (the peer's with a remote arrive: one hop less per
              // super-chunk than a relay warp in the peer; the hand-over chain is the critical path of a tile)
              mbar_wait_cluster_idle(bar_aconv(ab, sc), apar);
              TRACE(1, 300 + sc);
            }
            mbar_wait(bar_bfull(s), sph);
            if (nt == 0) TRACE(1, 500 + sc);
            tc_fence_after();
            const int p0 = 2 * sc, np = min(2, n_full - p0);
            const bool tails = (sc == n_sc - 1);
            const uint32_t b_addr = sB + s * P.b_stage;
            if (elect_one()) {
              if (!skip_mma) {
                for (int k = 0; k < np; ++k) {
                  const uint64_t bd0 = umma_desc(b_addr + (uint32_t)k * rows_b * 128u, 1024, 2);
                  const uint32_t a_tmem = a_base + 32u * (uint32_t)(p0 + k);
#pragma unroll
                  for (int kk = 0; kk < KC / KT; ++kk)
                    tc_mma_f16_ts2(d_tmem, a_tmem + 8u * kk, bd0 + 2u * kk, idesc, (sc | k | kk) != 0);
                }
                if (tails)
                  for (int t = 0; t < P.n_tail; ++t)
                    tc_mma_f16_ts2(d_tmem, a_base + 32u * n_full + 8u * t,
                                   umma_desc(b_addr + (uint32_t)np * rows_b * 128u + (uint32_t)t * rows_b * 32u, 256, 6), idesc, 1u);
              }
              tc_commit<2>(bar_bempty(s));
              if (last_nt) tc_commit<2>(bar_aempty(ab, sc));
              if (tails) tc_commit<2>(bar_accfull(as));
            }
            __syncwarp();
            if (++s == NB) { s = 0; sph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 2) {
    // =========================== row producer: fp32 slots of 128 rows x 32 columns ===========================
    uint32_t slot = 0, ph = 0;
    for (int tile = group; tile < P.n_row_tiles; tile += n_groups) {
      const long long r0 = ((long long)tile * 2 + cta_rank) * TM;
      const int row0 = (int)(r0 < P.N ? r0 : P.N - 1);       // a tile past the end reads (and ignores) the last row
      // a slot = 128 rows x 128 bytes: 32 fp32 columns (two slots per 64-column panel) or 64 16-bit columns (one)
      const int n_main = (Z32 ? 2 : 1) * n_full, n_slots = n_main + P.n_tail;
      for (int u = 0; u < n_slots; ++u) {
        mbar_wait_idle(bar_zempty(slot), ph ^ 1u);
        TRACE(2, u);
        if (P.flags & kDbgNoZ) {
          if (elect_one()) mbar_arrive(bar_zfull(slot));
        } else if (elect_one()) {
          if (u < n_main) {
            mbar_expect_tx(bar_zfull(slot), TME_ZSLOT);
            tma_load_2d<1>(sZ + slot * TME_ZSLOT, &tmZ, u * (Z32 ? TME_ZCOLS : KC), row0, bar_zfull(slot));
          } else {
            mbar_expect_tx(bar_zfull(slot), TM * KT * (uint32_t)sizeof(ZT));
            tma_load_2d<1>(sZ + slot * TME_ZSLOT, &tmZt, n_full * KC + (u - n_main) * KT, row0, bar_zfull(slot));
          }
        }
        __syncwarp();
        if (++slot == NZ) { slot = 0; ph ^= 1u; }
      }
    }
  } else if (warp >= EPI_WARP0 && warp < TME_CONV_WARP0) {
    // =========================== epilogue ===========================
    // as in tc_search_kernel; a warp's 16-column pieces are the codes [32 g + 16 eh, +16) inside the tile
    const int q = warp & 3;
    const int eh = (warp - EPI_WARP0) >> 2;
    const int r = q * 32 + lane;
    uint32_t* xrow = xch + (size_t)r * 12;
    uint32_t it = 0, ti = 0;
    const uint32_t e2_saddr = smem_u32(e2s) + 64u * (uint32_t)eh;      // ||e||^2 of code 16 eh
    for (int tile = group; tile < P.n_row_tiles; tile += n_groups, ++ti) {
      // Only the two key constants of the row stay in registers across the sweep (the key/top-2 chains need
      // every register they can get); the full per-row constants are rebuilt from the statistics afterwards.
      mbar_wait_idle(bar_rsfull(ti % RS_RING), (ti / RS_RING) & 1u);
      float key_cS, key_S;
      {
        const RowInfo r0 = make_rowinfo(rowstat[(ti % RS_RING) * TM + r].x, 0.f, 1.f, 0.f, 0.f, P.hdr->e2min, P.hdr->scale_e, P.n_ksteps);
        key_cS = r0.cS; key_S = r0.S;
      }
      uint32_t m1[16], m2[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) { m1[j] = 0xFFFFFFFFu; m2[j] = 0xFFFFFFFFu; }

      for (int nt = 0; nt < P.n_ntiles; ++nt, ++it) {
        const uint32_t as = it & 1u, around = it >> 1;
        const int s0 = nt * P.ntile;
        const int e_valid = min(s0 + ((nt == P.n_ntiles - 1) ? P.n_last : P.ntile), P.K);
        const int g0 = (s0 - 16 * eh + 31) >> 5, g1 = (e_valid - 16 * eh + 31) >> 5;     // pieces g0 .. g1-1
        if (warp == EPI_WARP0) TRACE(3, 100 + nt);
        mbar_wait_idle(bar_accfull(as), around & 1u);
        if (warp == EPI_WARP0) TRACE(3, 200 + nt);
        tc_fence_after();
        const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + acc_col_of(as) + (uint32_t)(16 * eh - s0);
        epi_sweep(taddr0, e2_saddr, g0, g1, e_valid, eh, key_cS, key_S, 0u, (P.flags & kDbgSkipEpi) != 0, m1, m2);
        tc_fence_before();
        __syncwarp();
        if (warp == EPI_WARP0) TRACE(3, 300 + nt);
        if (lane == 0) {
          if (leader) mbar_arrive(bar_accempty(as));
          else mbar_arrive_cluster(bar_accempty(as), 0);
        }
      }

      const Cand none{0xFFFFFFFFu, 0xFFFFFFFFu, 0};
      Cand c1 = none, c2 = none, c3 = none;
      uint32_t k4 = 0xFFFFFFFFu;
#pragma unroll
      for (int j = 0; j < 16; ++j) cand_insert(c1, c2, c3, k4, Cand{m1[j], m2[j], eh * 16 + j});
      if (eh == 1) {
        xrow[0] = c1.key; xrow[1] = c1.key2; xrow[2] = (uint32_t)c1.j;
        xrow[3] = c2.key; xrow[4] = c2.key2; xrow[5] = (uint32_t)c2.j;
        xrow[6] = c3.key; xrow[7] = c3.key2; xrow[8] = (uint32_t)c3.j;
        xrow[9] = k4;
      }
      asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
      if (eh == 0) {
        cand_insert(c1, c2, c3, k4, Cand{xrow[0], xrow[1], (int)xrow[2]});
        cand_insert(c1, c2, c3, k4, Cand{xrow[3], xrow[4], (int)xrow[5]});
        cand_insert(c1, c2, c3, k4, Cand{xrow[6], xrow[7], (int)xrow[8]});
        k4 = min(k4, xrow[9]);
        const long long row = ((long long)tile * 2 + cta_rank) * TM + r;
        const bool valid = row < P.N;
        const float2 st = rowstat[(ti % RS_RING) * TM + r];
        const RowInfo ri = make_rowinfo(st.x, st.y, 1.f, P.hdr->sfrac, P.hdr->rsub, P.hdr->e2min, P.hdr->scale_e, P.n_ksteps);
        finish_row(c1, c2, c3, k4, ri, row, valid, P.K, P.ntab, P.flags, P.idx, P.pair_list, P.chain_list, P.full_list, P.counters);
      }
      asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
    }
  } else if (warp >= TME_CONV_WARP0) {
    // =========================== converters ===========================
    // warp cw owns the 16 TMEM lanes (rows) 32 (cw % 4) + 16 (cw / 4) ...; in the 16x256b store pattern
    // lane i holds, per K step of 16, the elements 4 (i % 4) .. +3 of the rows i / 4 and i / 4 + 8.
    // A slot row is 128 bytes (32 fp32) with the 16-byte chunks XOR-swizzled by row % 8, so the eight rows
    // a quarter-warp reads hit all banks.
    const int cw = warp - TME_CONV_WARP0;
    const int lane0_row = 32 * (cw & 3) + 16 * (cw >> 2);
    const int ra_l = lane0_row + (lane >> 2), rb_l = ra_l + 8;
    const uint32_t t_lane = tmem_base + ((uint32_t)lane0_row << 16);
    const uint32_t kq = (uint32_t)(lane & 3), sw = (uint32_t)(ra_l & 7);            // (rb_l & 7) == (ra_l & 7)
    const uint32_t off_a0 = (uint32_t)ra_l * 128u + ((kq ^ sw) << 4), off_a1 = (uint32_t)ra_l * 128u + (((4u + kq) ^ sw) << 4);
    const bool odd = (ra_l & 1) != 0;
    const uint32_t off_f = odd ? off_a1 : off_a0, off_s = odd ? off_a0 : off_a1;
    const uint32_t off_t = (uint32_t)ra_l * 64u + kq * 16u;                        // tail slot: 64-byte rows, no swizzle
    // per-row sums as two partial sums each (even / odd elements), so that the squares and the rounding
    // residuals v - fp16(v) (exact in fp32) take packed FMAs
    float2 z2a = make_float2(0.f, 0.f), r2a = z2a, z2b = z2a, r2b = z2a;
    const float2 neg1 = make_float2(-1.f, -1.f);
    auto cvt = [&](const float4 v, float2& z2, float2& r2, uint32_t& w0, uint32_t& w1) {
      const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
      const float2 v01 = make_float2(v.x, v.y), v23 = make_float2(v.z, v.w);
      const float2 e01 = ffma2(__half22float2(h01), neg1, v01), e23 = ffma2(__half22float2(h23), neg1, v23);
      z2 = ffma2(v23, v23, ffma2(v01, v01, z2));
      r2 = ffma2(e23, e23, ffma2(e01, e01, r2));
      w0 = *reinterpret_cast<const uint32_t*>(&h01);
      w1 = *reinterpret_cast<const uint32_t*>(&h23);
    };
    uint32_t slot = 0, ph = 0, ti = 0;
    // one 64-column panel of the tile (a "chunk": two fp32 slots or one 16-bit slot) -> 16 packed fp16 words
    auto convert_main = [&](uint32_t (&w)[16]) {
      if constexpr (!Z32) {
        // 16-bit rows: one slot per panel, a row is 64 elements = 128 bytes (SWIZZLE_128B); K step s of row r is
        // the 16-byte chunks 2s, 2s+1 (xor r % 8); this lane takes 8 bytes (4 elements) of it
        mbar_wait(bar_zfull(slot), ph);
        const unsigned char* zs = gbase + sp.z_off + slot * TME_ZSLOT;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t sf = 2u * h + (odd ? 1u : 0u), ss = 2u * h + (odd ? 0u : 1u);   // odd rows: second K step first
          const uint32_t of = (uint32_t)ra_l * 128u + (((2u * sf + (kq >> 1)) ^ sw) << 4) + 8u * (kq & 1u);
          const uint32_t os = (uint32_t)ra_l * 128u + (((2u * ss + (kq >> 1)) ^ sw) << 4) + 8u * (kq & 1u);
          const uint2 af = *reinterpret_cast<const uint2*>(zs + of), bf = *reinterpret_cast<const uint2*>(zs + of + 8 * 128);
          const uint2 as = *reinterpret_cast<const uint2*>(zs + os), bs = *reinterpret_cast<const uint2*>(zs + os + 8 * 128);
          uint32_t f0, f1, f2, f3, s0, s1, s2, s3;
          cvt(unpack16<ZT>(af), z2a, r2a, f0, f1);
          cvt(unpack16<ZT>(bf), z2b, r2b, f2, f3);
          cvt(unpack16<ZT>(as), z2a, r2a, s0, s1);
          cvt(unpack16<ZT>(bs), z2b, r2b, s2, s3);
          w[8 * h + 0] = odd ? s0 : f0; w[8 * h + 1] = odd ? s1 : f1;
          w[8 * h + 2] = odd ? s2 : f2; w[8 * h + 3] = odd ? s3 : f3;
          w[8 * h + 4] = odd ? f0 : s0; w[8 * h + 5] = odd ? f1 : s1;
          w[8 * h + 6] = odd ? f2 : s2; w[8 * h + 7] = odd ? f3 : s3;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_zempty(slot));
        if (++slot == NZ) { slot = 0; ph ^= 1u; }
      } else {
        // fp32 rows: the panel is two slots.  Both are awaited and loaded before any arithmetic, and the two halves
        // accumulate their norms separately: a converter warp is bound by the latency of its own dependent
        // instructions (5.1 us per tile with nothing else running, ncu: 85 instructions per slot at ~5.5 cycles
        // each), so the eight conversions of a panel have to be independent work the scheduler can interleave.
        const uint32_t s1 = (slot + 1 == NZ) ? 0u : slot + 1u, ph1 = (slot + 1 == NZ) ? (ph ^ 1u) : ph;
        mbar_wait(bar_zfull(slot), ph);
        mbar_wait(bar_zfull(s1), ph1);
        if (P.flags & kDbgSkipConv) {              // trace builds: hand-shakes only
#pragma unroll
          for (int i = 0; i < 16; ++i) w[i] = 0u;
          __syncwarp();
          if (lane == 0) { mbar_arrive(bar_zempty(slot)); mbar_arrive(bar_zempty(s1)); }
          slot += 2;
          if (slot >= NZ) { slot -= NZ; ph ^= 1u; }
          return;
        }
        const unsigned char* zs0 = gbase + sp.z_off + slot * TME_ZSLOT;
        const unsigned char* zs1 = gbase + sp.z_off + s1 * TME_ZSLOT;
        // odd rows fetch their second K step first: the two rows of a quarter-warp then sit in different
        // halves of the swizzled 128-byte line (no bank conflict); the words are put back in order below
        float4 af[2], bf[2], as[2], bs[2];
        af[0] = *reinterpret_cast<const float4*>(zs0 + off_f); bf[0] = *reinterpret_cast<const float4*>(zs0 + off_f + 8 * 128);
        as[0] = *reinterpret_cast<const float4*>(zs0 + off_s); bs[0] = *reinterpret_cast<const float4*>(zs0 + off_s + 8 * 128);
        af[1] = *reinterpret_cast<const float4*>(zs1 + off_f); bf[1] = *reinterpret_cast<const float4*>(zs1 + off_f + 8 * 128);
        as[1] = *reinterpret_cast<const float4*>(zs1 + off_s); bs[1] = *reinterpret_cast<const float4*>(zs1 + off_s + 8 * 128);
        float2 za[2], ra[2], zb[2], rb[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          za[h] = ra[h] = zb[h] = rb[h] = make_float2(0.f, 0.f);
          uint32_t f0, f1, f2, f3, s0, s1w, s2, s3;
          cvt(af[h], za[h], ra[h], f0, f1);
          cvt(bf[h], zb[h], rb[h], f2, f3);
          cvt(as[h], za[h], ra[h], s0, s1w);
          cvt(bs[h], zb[h], rb[h], s2, s3);
          w[8 * h + 0] = odd ? s0 : f0; w[8 * h + 1] = odd ? s1w : f1;
          w[8 * h + 2] = odd ? s2 : f2; w[8 * h + 3] = odd ? s3 : f3;
          w[8 * h + 4] = odd ? f0 : s0; w[8 * h + 5] = odd ? f1 : s1w;
          w[8 * h + 6] = odd ? f2 : s2; w[8 * h + 7] = odd ? f3 : s3;
        }
        z2a.x += za[0].x + za[1].x; z2a.y += za[0].y + za[1].y;
        r2a.x += ra[0].x + ra[1].x; r2a.y += ra[0].y + ra[1].y;
        z2b.x += zb[0].x + zb[1].x; z2b.y += zb[0].y + zb[1].y;
        r2b.x += rb[0].x + rb[1].x; r2b.y += rb[0].y + rb[1].y;
        __syncwarp();
        if (lane == 0) { mbar_arrive(bar_zempty(slot)); mbar_arrive(bar_zempty(s1)); }
        slot += 2;
        if (slot >= NZ) { slot -= NZ; ph ^= 1u; }
      }
    };
    auto publish = [&](int ab, int sc) {          // the super-chunk's TMEM stores are done: hand it to the MMA issuer
      if (cw == 0) TRACE(4, 300 + sc);
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(bar_aconv(ab, sc));
        else mbar_arrive_cluster(bar_aconv(ab, sc), 0);
      }
      if (cw == 0) TRACE(4, 400 + sc);
    };
    // With at least two super-chunks, super-chunk 0 (panels 0 and 1) is converted into REGISTERS before the A
    // buffer is free: the converters otherwise idle ~3 us per tile waiting for the last code tile's MMAs to release
    // it, and the conversion that follows is the critical path of the tile (profiles/r2_trace_tc_tmem_k400.txt).
    const bool ahead = (n_sc >= 2) && P.ahead != 0;
    for (int tile = group; tile < P.n_row_tiles; tile += n_groups, ++ti) {
      const int ab = (int)(ti & abm);
      const uint32_t t_buf = t_lane + (uint32_t)ab * a_cols;
      const uint32_t a_par = ((ti >> absh) & 1u) ^ 1u;
      int c_first = 0;
      if (ahead) {
        uint32_t w0[16], w1[16];
        convert_main(w0);
        convert_main(w1);
        if (cw == 0) TRACE(4, 100);
        mbar_wait(bar_aempty(ab, 0), a_par);      // the MMAs of the buffer's previous row tile have read these panels
        if (cw == 0) TRACE(4, 200);
        tc_fence_after();
        tc_st_16x256b_x4(t_buf, w0);
        tc_st_16x256b_x4(t_buf + 32u, w1);
        publish(ab, 0);
        c_first = 2;
      }
      for (int c = c_first; c < n_chunks; ++c) {
        // barriers work on super-chunks: panels 2 sc and 2 sc + 1, the tails belong to the last one
        const int sc = min(c >> 1, n_sc - 1);
        const bool sc_first = (c < n_full) && ((c & 1) == 0);
        const bool sc_last = (c == n_chunks - 1) || (c < n_full - 1 && (c & 1) == 1) || (c == n_full - 1 && sc < n_sc - 1);
        if (sc_first) {
          if (cw == 0) TRACE(4, 100 + sc);
          mbar_wait(bar_aempty(ab, sc), a_par);   // the MMAs of the buffer's previous row tile have read these panels
          if (cw == 0) TRACE(4, 200 + sc);
          tc_fence_after();
        }
        if (c < n_full) {
          uint32_t w[16];
          convert_main(w);
          tc_st_16x256b_x4(t_buf + 32u * c, w);
        } else if (!Z32) {
          mbar_wait(bar_zfull(slot), ph);
          const unsigned char* zs = gbase + sp.z_off + slot * TME_ZSLOT;                 // tail slot: 32-byte rows, no swizzle
          const uint32_t ot = (uint32_t)ra_l * 32u + kq * 8u;
          const uint2 a0 = *reinterpret_cast<const uint2*>(zs + ot), b0 = *reinterpret_cast<const uint2*>(zs + ot + 8 * 32);
          uint32_t w0, w1, w2, w3;
          cvt(unpack16<ZT>(a0), z2a, r2a, w0, w1);
          cvt(unpack16<ZT>(b0), z2b, r2b, w2, w3);
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_zempty(slot));
          if (++slot == NZ) { slot = 0; ph ^= 1u; }
          tc_st_16x256b_x1(t_buf + 32u * n_full + 8u * (c - n_full), w0, w1, w2, w3);
        } else {
          mbar_wait(bar_zfull(slot), ph);
          const unsigned char* zs = gbase + sp.z_off + slot * TME_ZSLOT;
          const float4 a0 = *reinterpret_cast<const float4*>(zs + off_t), b0 = *reinterpret_cast<const float4*>(zs + off_t + 8 * 64);
          uint32_t w0, w1, w2, w3;
          cvt(a0, z2a, r2a, w0, w1);
          cvt(b0, z2b, r2b, w2, w3);
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_zempty(slot));
          if (++slot == NZ) { slot = 0; ph ^= 1u; }
          tc_st_16x256b_x1(t_buf + 32u * n_full + 8u * (c - n_full), w0, w1, w2, w3);
        }
        if (sc_last) publish(ab, sc);
      }
      // row statistics of the finished tile: sum over the four lanes that share a row
      float sza = z2a.x + z2a.y, sra = r2a.x + r2a.y, szb = z2b.x + z2b.y, srb = r2b.x + r2b.y;
      sza += __shfl_xor_sync(0xffffffffu, sza, 1); sra += __shfl_xor_sync(0xffffffffu, sra, 1);
      szb += __shfl_xor_sync(0xffffffffu, szb, 1); srb += __shfl_xor_sync(0xffffffffu, srb, 1);
      sza += __shfl_xor_sync(0xffffffffu, sza, 2); sra += __shfl_xor_sync(0xffffffffu, sra, 2);
      szb += __shfl_xor_sync(0xffffffffu, szb, 2); srb += __shfl_xor_sync(0xffffffffu, srb, 2);
      if ((lane & 3) == 0) {
        rowstat[(ti % RS_RING) * TM + ra_l] = make_float2(sza, sra);
        rowstat[(ti % RS_RING) * TM + rb_l] = make_float2(szb, srb);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_rsfull(ti % RS_RING));
      z2a = r2a = z2b = r2b = make_float2(0.f, 0.f);
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// ------------------------------------------------------------------------------------------
// row preparation: fp16 operand rows + per-row constants
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float ld_f32(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_f32(const __half* p) { return __half2float(*p); }
__device__ __forceinline__ float ld_f32(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename ZT>
__global__ void __launch_bounds__(256) row_prep_kernel(const ZT* __restrict__ z, long long N, int D, int Dp,
                                                       const CbHeader* __restrict__ hdr, __half* __restrict__ z16,
                                                       RowInfo* __restrict__ rowinfo, int* counters, int n_ksteps) {
  if (blockIdx.x == 0 && threadIdx.x < 4) counters[threadIdx.x] = 0;
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= N) return;
  const ZT* zr = z + (size_t)row * D;
  float am = 0.f, s2 = 0.f;
  for (int j = lane; j < D; j += 32) {
    float v = ld_f32(zr + j);
    am = fmaxf(am, fabsf(v));
    s2 = fmaf(v, v, s2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, o));
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  float sc = 1.f;
  if (am > 0.f && isfinite(am)) {
    int e;
    frexpf(am, &e);
    sc = ldexpf(1.f, max(min(9 - e, 100), -100));      // row max lands in [256, 512)
  }
  const float inv = 1.f / sc;
  float r2 = 0.f;
  __half* o = z16 + (size_t)row * Dp;
  for (int j = lane; j < Dp; j += 32) {
    float v = (j < D) ? ld_f32(zr + j) : 0.f;
    __half h = __float2half_rn(v * sc);
    float r = v - __half2float(h) * inv;
    r2 = fmaf(r, r, r2);
    o[j] = h;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) r2 += __shfl_xor_sync(0xffffffffu, r2, off);
  if (lane == 0) {
    const RowInfo ri = make_rowinfo(s2, r2, inv, hdr->sfrac, hdr->rsub, hdr->e2min, hdr->scale_e, n_ksteps);
    rowinfo[row] = ri;
  }
}

// exact re-rank of the two or three candidate codes of each listed row (fp64).  LPE lanes share an entry, so a
// warp keeps 32 / LPE entries in flight: the pass is bound by the latency of the row fetch (random rows, long
// gone from L2), not by arithmetic.  The third code is only read when there is one.  Lowest index wins exact ties.
template <typename ZT, int LPE>
__device__ __forceinline__ void pair_entries(const ZT* __restrict__ z, const float* __restrict__ E, int D, int Dz,
                                             const int* __restrict__ pair_list, int n, int* __restrict__ idx) {
  constexpr int EPW = 32 / LPE;                          // entries per warp
  constexpr int U = 2;                                   // float4 columns in flight per array (registers -> occupancy)
  const int lane = threadIdx.x & 31, sub = lane / LPE, sl = lane % LPE;
  const int wstride = (int)((gridDim.x * blockDim.x) >> 5);
  // rows have Dz <= D columns (zero beyond: a folded codebook is wider than the rows it is searched with)
  const bool vec = (D % 4 == 0) && (Dz % 4 == 0) && sizeof(ZT) == 4 && ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(E)) & 15) == 0;
  for (int e0 = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * EPW; e0 < n; e0 += wstride * EPW) {
    const int e = e0 + sub;
    const bool live = e < n;
    const int4 ent = live ? reinterpret_cast<const int4*>(pair_list)[e] : make_int4(0, 0, 0, -1);
    const int row = ent.x, a = ent.y, b = ent.z, c = ent.w;
    const bool has_c = c >= 0;
    const ZT* zr = z + (size_t)row * Dz;
    const float* ea = E + (size_t)a * D;
    const float* eb = E + (size_t)b * D;
    const float* ec = E + (size_t)(has_c ? c : a) * D;
    double da = 0.0, db = 0.0, dc = 0.0;
    if (live) {
      if (vec) {
        const float* zf = reinterpret_cast<const float*>(zr);
        for (int j0 = sl * 4; j0 < D; j0 += LPE * 4 * U) {
          float4 zv[U], av[U], bv[U], cv[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int j = j0 + LPE * 4 * u;
            zv[u] = av[u] = bv[u] = cv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < D) {
              if (j < Dz) zv[u] = __ldg(reinterpret_cast<const float4*>(zf + j));
              av[u] = __ldg(reinterpret_cast<const float4*>(ea + j));
              bv[u] = __ldg(reinterpret_cast<const float4*>(eb + j));
              if (has_c) cv[u] = __ldg(reinterpret_cast<const float4*>(ec + j));
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const float zz[4] = {zv[u].x, zv[u].y, zv[u].z, zv[u].w};
            const float aa[4] = {av[u].x, av[u].y, av[u].z, av[u].w};
            const float bb[4] = {bv[u].x, bv[u].y, bv[u].z, bv[u].w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const double zd = (double)zz[t];
              const double xa = zd - (double)aa[t], xb = zd - (double)bb[t];
              da = fma(xa, xa, da);
              db = fma(xb, xb, db);
            }
            if (has_c) {
              const float cc[4] = {cv[u].x, cv[u].y, cv[u].z, cv[u].w};
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const double xc = (double)zz[t] - (double)cc[t];
                dc = fma(xc, xc, dc);
              }
            }
          }
        }
      } else {
        for (int j = sl; j < D; j += LPE) {
          const double zv = j < Dz ? (double)ld_f32(zr + j) : 0.0;
          const double xa = zv - (double)__ldg(ea + j), xb = zv - (double)__ldg(eb + j);
          da = fma(xa, xa, da);
          db = fma(xb, xb, db);
          if (has_c) {
            const double xc = zv - (double)__ldg(ec + j);
            dc = fma(xc, xc, dc);
          }
        }
      }
    }
#pragma unroll
    for (int o = LPE / 2; o > 0; o >>= 1) {
      da += __shfl_xor_sync(0xffffffffu, da, o);
      db += __shfl_xor_sync(0xffffffffu, db, o);
      dc += __shfl_xor_sync(0xffffffffu, dc, o);
    }
    if (live && sl == 0) {
      int best = a;
      double dbest = da;
      if (db < dbest || (db == dbest && b < best)) { best = b; dbest = db; }
      if (has_c && (dc < dbest || (dc == dbest && c < best))) { best = c; dbest = dc; }
      idx[row] = best;
    }
  }
}

// exact re-rank over one chain (codes j, j+32, ... < K) plus up to two extra codes (one warp per
// entry, fp64, lanes split the dimensions); lowest index wins exact ties
// BLOCK = false: one warp per entry.  BLOCK = true (large K: a chain is K / 32 codes): one CTA per entry, the
// warps split the chain and combine through shared memory.
template <typename ZT, bool BLOCK>
__device__ __forceinline__ void chain_entries(const ZT* __restrict__ z, const float* __restrict__ E, int K, int D, int Dz,
                                              const int* __restrict__ chain_list, int n, int* __restrict__ idx,
                                              double* sh_v, int* sh_i) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wstride = BLOCK ? (int)gridDim.x : (int)((gridDim.x * blockDim.x) >> 5);
  for (int e = BLOCK ? (int)blockIdx.x : (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); e < n; e += wstride) {
    const int4 ent = reinterpret_cast<const int4*>(chain_list)[e];
    const ZT* zr = z + (size_t)ent.x * Dz;
    double best = INFINITY;
    int besti = 0x7fffffff;
    const int nchain = (K - ent.y + 31) / 32;
    for (int t = BLOCK ? warp : 0; t < nchain + 2; t += BLOCK ? 8 : 1) {
      const int k = t < nchain ? ent.y + 32 * t : (t == nchain ? ent.z : ent.w);
      if (k < 0 || k >= K) continue;                                // warp-uniform
      const float* er = E + (size_t)k * D;
      double s = 0.0;
      for (int j0 = lane; j0 < D; j0 += 32 * 8) {
        float zv[8], ev[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = j0 + 32 * u;
          zv[u] = j < Dz ? ld_f32(zr + j) : 0.f;
          ev[u] = j < D ? __ldg(er + j) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const double df = (double)zv[u] - (double)ev[u];
          s = fma(df, df, s);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (s < best || (s == best && k < besti)) { best = s; besti = k; }
    }
    if constexpr (BLOCK) {
      __syncthreads();
      if (lane == 0) { sh_v[warp] = best; sh_i[warp] = besti; }
      __syncthreads();
      if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w)
          if (sh_v[w] < best || (sh_v[w] == best && sh_i[w] < besti)) { best = sh_v[w]; besti = sh_i[w]; }
        if (besti != 0x7fffffff) idx[ent.x] = besti;
      }
    } else {
      if (lane == 0 && besti != 0x7fffffff) idx[ent.x] = besti;
    }
  }
}

// ------------------------------------------------------------------------------------------
// refine pass: the listed WHOLE rows, a second time on the tensor cores at fp32 accuracy
// ------------------------------------------------------------------------------------------
// A whole-row re-rank costs K x D fp64 operations; one row in a hundred ends up there (K = 400, iid rows) and
// those rows were most of the re-rank time (and ALL of the step for a degenerate EMA codebook, where most rows
// are uncertain at fp16 accuracy).  The listed rows therefore take a second tensor-core pass first:
//   refine_prep_kernel   gathers the listed rows and the codes, each scaled by its OWN power of two (largest
//                        entry in [256, 512)), split into fp16 hi + lo terms laid out for the 3-term product
//                        (g2v_gemm.cu):  A' = [hi | lo | hi],  B' = [hi | hi | lo]
//   tc_gemm_kernel       dots[i, k] = A'_i . B'_k for the listed rows (row count read on the device)
//   rerank_kernel        per listed row: d^_k = e2_k - 2 dots[i, k] / (s_i t_k); every code whose lower bound
//                        d^_k - err_k does not exceed min_k (d^_k + err_k) is evaluated exactly in fp64
//                        (usually one or two codes), lowest index winning exact ties.
// Error of one dot product, relative to |z| |e_k| (u = 2^-11, fp16 unit roundoff; x = scaled entry, |x| < 512):
//   hi = rn(x), lo = rn(x - hi):  |lo| <= u |x| + 2^-25,  |x - hi - lo| <= u^2 |x| + 2^-25   (2^-25: fp16 subnormal grid)
//   with the largest entry >= 256 the 2^-25 terms add < 2.6e-9 to either ratio:  lam = 4.8829e-4,  rho = 2.411e-7
//   dropped lo.lo term + residuals:  lam^2 + 2 rho + rho^2 <= 7.3e-7
//   tensor-core accumulation (same model as the first pass): 1.003 (2^-19 + (k-steps + 2) 2^-23)
// The comparison itself runs in fp32: e2_k carries one rounding (2^-24 e2_k), the fma and the +-err one each.
struct RefineRow {
  float inv_s;      // 1 / s_i  (power of two)
  float znorm;      // upper bound of |z_i|
  float pad0, pad1;
};

constexpr int64_t kRefineMinRows = 32768;          // below this a step is launch-bound: keep the two-launch re-rank
// The listed rows take the pass when there is enough fp64 work to save (rows x codes): three dependent launches
// cost ~35 us, K x D fp64 operations per row cost ~0.25 ns per code.  Evaluated on the device by every kernel
// involved, from the same counter, so that they agree.  Below the threshold fewer than kFull64Cap rows are listed
// (K >= 64), i.e. rerank_kernel's own fp64 path covers all of them.
constexpr long long kRefineMinWork = 131072;
__device__ __forceinline__ int refine_count(int listed, int K, int cap) {
  const int n = min(max(listed, 0), cap);
  return (long long)n * K >= kRefineMinWork ? n : 0;
}
constexpr size_t kRefineBudget = (size_t)4 << 30;  // bytes of workspace the pass may claim (of 180 GB)

struct RefineWs {
  long long cap;        // listed rows the pass can take (0 = pass disabled); the rest goes to the batched fp32+fp64 kernel
  int kp, ldc;          // fp16 per operand segment (D rounded up to 64), row pitch of the dots
  size_t invt, rows, A, B, C, total;      // offsets relative to the start of the refine region
};
inline size_t al256r(size_t v) { return (v + 255) / 256 * 256; }
RefineWs refine_ws(int64_t N, int K, int D) {
  RefineWs r = {};
  if (N < kRefineMinRows || D > 512 || K < 64) return r;
  r.kp = round_up(D, 64);
  r.ldc = round_up(K, 4);
  const size_t per_row = sizeof(RefineRow) + (size_t)3 * r.kp * 2 + (size_t)r.ldc * 4;
  long long cap = (long long)(kRefineBudget / per_row);
  if (cap > N) cap = N;
  cap = cap / 256 * 256;
  if (cap < 256) return r;
  r.cap = cap;
  r.invt = 0;
  r.rows = r.invt + al256r((size_t)K * 4);
  r.A = r.rows + al256r((size_t)cap * sizeof(RefineRow));
  r.B = r.A + al256r((size_t)cap * 3 * r.kp * 2);
  r.C = r.B + al256r((size_t)K * 3 * r.kp * 2);
  r.total = r.C + al256r((size_t)cap * r.ldc * 4);
  return r;
}

// one warp per code (tasks [0, K)) or listed row (tasks [K, K + n)); D <= 512: a lane holds 4 x 4 entries
template <typename ZT>
__global__ void __launch_bounds__(256) refine_prep_kernel(const ZT* __restrict__ z, int Dz, const float* __restrict__ E, int K, int D,
                                                          int kp, const int* __restrict__ full_list,
                                                          const int* __restrict__ full_count, int cap, __half* __restrict__ A16,
                                                          __half* __restrict__ B16, float* __restrict__ invt,
                                                          RefineRow* __restrict__ rows, int* __restrict__ n_eff) {
  const int lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0) *n_eff = refine_count(*full_count, K, cap);   // the GEMM's row count
  const int gwarp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), nwarps = (int)((gridDim.x * blockDim.x) >> 5);
  const int n = refine_count(*full_count, K, cap);
  if (n == 0) return;
  for (int task = gwarp; task < K + n; task += nwarps) {
    const bool is_code = task < K;
    const int i = is_code ? task : task - K;
    float v[4][4];
    float amax = 0.f, n2 = 0.f;
    if (is_code) {
      const float* src = E + (size_t)i * D;
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int j = 4 * (lane + 32 * u) + t;
          v[u][t] = j < D ? __ldg(src + j) : 0.f;
        }
    } else {
      const ZT* src = z + (size_t)full_list[i] * Dz;
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int j = 4 * (lane + 32 * u) + t;
          v[u][t] = j < Dz ? ld_f32(src + j) : 0.f;
        }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        amax = fmaxf(amax, fabsf(v[u][t]));        // fmaxf drops a NaN: the split below still carries it into the dots
        n2 = fmaf(v[u][t], v[u][t], n2);
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
      n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    }
    float sc = 1.f;
    if (amax > 0.f && isfinite(amax)) {
      int e;
      frexpf(amax, &e);
      sc = ldexpf(1.f, max(min(9 - e, 120), -120));
    }
    __half* dst = (is_code ? B16 : A16) + (size_t)i * 3 * kp;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = 4 * (lane + 32 * u);
      if (c < kp) {
        __half2 h[2], l[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const float x0 = v[u][2 * q] * sc, x1 = v[u][2 * q + 1] * sc;
          h[q] = __floats2half2_rn(x0, x1);
          const float2 hf = __half22float2(h[q]);
          l[q] = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
        }
        uint2 hi, lo;
        hi.x = *reinterpret_cast<const uint32_t*>(&h[0]); hi.y = *reinterpret_cast<const uint32_t*>(&h[1]);
        lo.x = *reinterpret_cast<const uint32_t*>(&l[0]); lo.y = *reinterpret_cast<const uint32_t*>(&l[1]);
        *reinterpret_cast<uint2*>(dst + c) = hi;
        *reinterpret_cast<uint2*>(dst + kp + c) = is_code ? hi : lo;
        *reinterpret_cast<uint2*>(dst + 2 * kp + c) = is_code ? lo : hi;
      }
    }
    if (lane == 0) {
      if (is_code) {
        invt[i] = 1.f / sc;
      } else {
        RefineRow r;
        r.inv_s = 1.f / sc;
        r.znorm = sqrtf(n2) * 1.0001f;
        r.pad0 = r.pad1 = 0.f;
        rows[i] = r;
      }
    }
  }
}

struct RefineArgs {
  const float* dots;          // [cap][ldc] raw accumulators, null = pass disabled
  const float* invt;          // [K]
  const RefineRow* rows;      // [cap]
  const float* e2;            // [K] |e_k|^2 (fp64 accumulated, rounded once)
  int ldc, cap;
  float beta;                 // dot-product error relative to |z| |e_k|
};

// fp64 distance of one row to one code, lanes split the dimensions (all lanes get the sum)
template <typename ZT>
__device__ __forceinline__ double exact_dist_warp(const ZT* __restrict__ zr, const float* __restrict__ er, int D, int Dz, int lane) {
  double s = 0.0;
  for (int j0 = lane; j0 < D; j0 += 32 * 8) {
    float zv[8], ev[8];
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      const int j = j0 + 32 * v;
      zv[v] = j < Dz ? ld_f32(zr + j) : 0.f;
      ev[v] = j < D ? __ldg(er + j) : 0.f;
    }
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      const double df = (double)zv[v] - (double)ev[v];
      s = fma(df, df, s);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}

// K <= 32 KI: a lane keeps the lower bounds of its KI codes in registers -- one pass over the row's dots instead of
// two (the collapse regime lists most rows: 574 k rows took 1.8 ms with the two-pass kernel below)
template <typename ZT, int KI>
__global__ void __launch_bounds__(256, 4) refine_rows_reg_kernel(const ZT* __restrict__ z, const float* __restrict__ E, int K, int D,
                                                                 int Dz, const int* __restrict__ full_list,
                                                                 const int* __restrict__ full_count, const RefineArgs R,
                                                                 int* __restrict__ idx, unsigned long long* stats) {
  const int lane = threadIdx.x & 31;
  const int n = refine_count(*full_count, K, R.cap);
  unsigned long long* n_exact = stats ? stats + G2V_STAT_REFINE_EXACT : nullptr;
  if (stats && blockIdx.x == 0 && threadIdx.x == 0 && n) atomicAdd(stats + G2V_STAT_REFINE_ROWS, (unsigned long long)n);
  const int wstride = (int)((gridDim.x * blockDim.x) >> 5);
  unsigned long long exact = 0;
  for (int e = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); e < n; e += wstride) {
    const int row = full_list[e];
    const RefineRow ri = R.rows[e];
    const float* dr = R.dots + (size_t)e * R.ldc;
    const float c1 = (2.f * R.beta + 4.7683716e-7f) * ri.znorm, c2 = 4.7683716e-7f;    // 2^-21: the fp32 steps of the comparison
    float raw[KI], lo[KI];
#pragma unroll
    for (int i = 0; i < KI; ++i) raw[i] = (lane + 32 * i < K) ? __ldcs(dr + lane + 32 * i) : 0.f;
    float U = INFINITY;
#pragma unroll
    for (int i = 0; i < KI; ++i) {
      const int k = lane + 32 * i;
      lo[i] = INFINITY;
      if (k < K) {
        const float e2k = __ldg(R.e2 + k);
        const float dot = (raw[i] * ri.inv_s) * __ldg(R.invt + k);
        const float d = fmaf(-2.f, dot, e2k);
        const float err = fmaf(c1, sqrtf(e2k) * 1.0001f, c2 * e2k);
        lo[i] = d - err;
        U = fminf(U, d + err);                  // a NaN bound is ignored here and becomes a candidate below
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) U = fminf(U, __shfl_xor_sync(0xffffffffu, U, o));
    const ZT* zr = z + (size_t)row * Dz;
    double best = INFINITY;
    int besti = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < KI; ++i) {
      const bool cand = (lane + 32 * i < K) && !(lo[i] > U);
      unsigned mask = __ballot_sync(0xffffffffu, cand);
      while (mask) {
        const int kk = 32 * i + __ffs((int)mask) - 1;
        mask &= mask - 1;
        const double s = exact_dist_warp<ZT>(zr, E + (size_t)kk * D, D, Dz, lane);
        if (s < best) { best = s; besti = kk; }         // ascending k: the first of equal distances stays
        ++exact;
      }
    }
    if (lane == 0) idx[row] = besti == 0x7fffffff ? 0 : besti;    // every distance NaN: index 0, as torch.argmin
  }
  if (n_exact && lane == 0 && exact) atomicAdd(n_exact, exact);
}

// one warp per listed row: candidates from the fp32-accurate dots, then fp64 on the candidates
template <typename ZT>
__global__ void __launch_bounds__(256, 4) refine_rows_kernel(const ZT* __restrict__ z, const float* __restrict__ E, int K, int D, int Dz,
                                                             const int* __restrict__ full_list, const int* __restrict__ full_count,
                                                             const RefineArgs R, int* __restrict__ idx, unsigned long long* stats) {
  const int lane = threadIdx.x & 31;
  const int n = refine_count(*full_count, K, R.cap);
  unsigned long long* n_exact = stats ? stats + G2V_STAT_REFINE_EXACT : nullptr;
  if (stats && blockIdx.x == 0 && threadIdx.x == 0 && n) atomicAdd(stats + G2V_STAT_REFINE_ROWS, (unsigned long long)n);
  const int wstride = (int)((gridDim.x * blockDim.x) >> 5);
  unsigned long long exact = 0;
  for (int e = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); e < n; e += wstride) {
    const int row = full_list[e];
    const RefineRow ri = R.rows[e];
    const float* dr = R.dots + (size_t)e * R.ldc;
    const float c1 = (2.f * R.beta + 4.7683716e-7f) * ri.znorm, c2 = 4.7683716e-7f;    // 2^-21: the fp32 steps of the comparison
    auto bounds = [&](int k, float& lo, float& hi) {
      const float e2k = __ldg(R.e2 + k);
      const float dot = (__ldg(dr + k) * ri.inv_s) * __ldg(R.invt + k);
      const float d = fmaf(-2.f, dot, e2k);
      const float err = fmaf(c1, sqrtf(e2k) * 1.0001f, c2 * e2k);
      lo = d - err;
      hi = d + err;
    };
    constexpr int UK = 4;                       // 32-code groups in flight (the loads of a row are a dependent chain otherwise)
    float U = INFINITY;
    for (int k0 = lane; k0 < K; k0 += 32 * UK) {
      float lo[UK], hi[UK];
#pragma unroll
      for (int u = 0; u < UK; ++u) {
        hi[u] = INFINITY;
        if (k0 + 32 * u < K) bounds(k0 + 32 * u, lo[u], hi[u]);
      }
#pragma unroll
      for (int u = 0; u < UK; ++u) U = fminf(U, hi[u]);      // a NaN bound is ignored here and becomes a candidate below
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) U = fminf(U, __shfl_xor_sync(0xffffffffu, U, o));
    const ZT* zr = z + (size_t)row * Dz;
    double best = INFINITY;
    int besti = 0x7fffffff;
    for (int k0 = 0; k0 < K; k0 += 32 * UK) {
      bool cand[UK];
#pragma unroll
      for (int u = 0; u < UK; ++u) {
        const int k = k0 + 32 * u + lane;
        cand[u] = false;
        if (k < K) {
          float lo, hi;
          bounds(k, lo, hi);
          cand[u] = !(lo > U);
        }
      }
#pragma unroll
      for (int u = 0; u < UK; ++u) {
        unsigned mask = __ballot_sync(0xffffffffu, cand[u]);
        while (mask) {
          const int kk = k0 + 32 * u + __ffs((int)mask) - 1;
          mask &= mask - 1;
          const float* er = E + (size_t)kk * D;
          double s = 0.0;
          for (int j0 = lane; j0 < D; j0 += 32 * 8) {
            float zv[8], ev[8];
#pragma unroll
            for (int v = 0; v < 8; ++v) {
              const int j = j0 + 32 * v;
              zv[v] = j < Dz ? ld_f32(zr + j) : 0.f;
              ev[v] = j < D ? __ldg(er + j) : 0.f;
            }
#pragma unroll
            for (int v = 0; v < 8; ++v) {
              const double df = (double)zv[v] - (double)ev[v];
              s = fma(df, df, s);
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          if (s < best) { best = s; besti = kk; }       // ascending k: the first of equal distances stays
          ++exact;
        }
      }
    }
    if (lane == 0) idx[row] = besti == 0x7fffffff ? 0 : besti;    // every distance NaN: index 0, as torch.argmin
  }
  if (n_exact && lane == 0 && exact) atomicAdd(n_exact, exact);
}

// One launch for every exact re-rank of the tensor-core pass: the listed whole rows first (fp64 over all K
// codes; each row split into `slices` CTAs that combine through `scratch`, the last one to arrive writes the
// index), then the chain entries, then the two/three-candidate entries.  All CTAs walk all three lists, so
// the short long-latency lists overlap with the long cheap one instead of running as three serial kernels.
struct RerankSlot {
  double v;
  int i;
  int pad;
};
constexpr int kRerankMaxSlices = 32;
#ifndef G2V_PAIR_LANES
#define G2V_PAIR_LANES 8
#endif
constexpr int kPairLanes = G2V_PAIR_LANES;        // lanes per two/three-candidate entry (measured: see DESIGN.md)
constexpr size_t kRerankScratchBytes = (size_t)kFull64Cap * kRerankMaxSlices * sizeof(RerankSlot);
constexpr size_t kRerankArriveBytes = (size_t)kFull64Cap * sizeof(int);

template <typename ZT>
__global__ void __launch_bounds__(256, 4) rerank_kernel(const ZT* __restrict__ z, const float* __restrict__ E, int K, int D, int Dz,
                                                     const int* __restrict__ pair_list, const int* __restrict__ chain_list,
                                                     const int* __restrict__ full_list, const int* __restrict__ counters,
                                                     int slices, RerankSlot* scratch, int* arrive, int* __restrict__ idx,
                                                     unsigned long long* stats, const int refined /* cap of the refine pass, 0 = none */) {
  extern __shared__ __align__(16) unsigned char rr_smem[];
  float* zs = reinterpret_cast<float*>(rr_smem);            // one row, zero padded to a multiple of 32
  __shared__ double bv[8];
  __shared__ int bi[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_pair = counters[0], n_chain = counters[2];
  // whole rows: straight fp64 here unless the refine pass takes them (refine_rows_kernel, on its own stream)
  const int n_full = (refined && refine_count(counters[1], K, refined) > 0) ? 0 : min(counters[1], kFull64Cap);

  // ---- whole rows ----
  const int Dp = (D + 31) & ~31;
  const int per = ((K + slices - 1) / slices + 3) & ~3;     // codes per slice
  constexpr int KC4 = 4;                                     // codes in flight per warp (latency hiding)
  for (int item = blockIdx.x; item < n_full * slices; item += gridDim.x) {
    const int e = item / slices, sl = item - e * slices;
    const int row = full_list[e];
    const int k0 = sl * per, k1 = min(K, k0 + per);
    __syncthreads();
    for (int j = threadIdx.x; j < Dp; j += blockDim.x) zs[j] = j < Dz ? ld_f32(z + (size_t)row * Dz + j) : 0.f;
    __syncthreads();
    double best = INFINITY;
    int besti = 0x7fffffff;
    for (int kb = k0 + warp * KC4; kb < k1; kb += 8 * KC4) {
      double acc[KC4];
#pragma unroll
      for (int c = 0; c < KC4; ++c) acc[c] = 0.0;
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
              acc[c] = fma(df, df, acc[c]);
            }
          }
        }
      }
#pragma unroll
      for (int c = 0; c < KC4; ++c) {
        double t = acc[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (kb + c < k1 && t < best) { best = t; besti = kb + c; }     // ascending k inside a warp: first wins
      }
    }
    if (lane == 0) { bv[warp] = best; bi[warp] = besti; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double v = bv[0];
      int id = bi[0];
      for (int w = 1; w < 8; ++w)
        if (bv[w] < v || (bv[w] == v && bi[w] < id)) { v = bv[w]; id = bi[w]; }
      if (slices == 1) {
        idx[row] = id == 0x7fffffff ? 0 : id;       // every distance NaN: index 0, as torch.argmin
      } else {
        RerankSlot* slot = scratch + (size_t)e * slices;
        slot[sl].v = v;
        slot[sl].i = id;
        __threadfence();
        if (atomicAdd(arrive + e, 1) == slices - 1) {       // the last slice of this row: combine
          __threadfence();
          const volatile RerankSlot* vs = slot;
          v = vs[0].v; id = vs[0].i;
          for (int q = 1; q < slices; ++q) {
            const double qv = vs[q].v;
            const int qi = vs[q].i;
            if (qv < v || (qv == v && qi < id)) { v = qv; id = qi; }
          }
          idx[row] = id == 0x7fffffff ? 0 : id;
          arrive[e] = 0;                                     // ready for the next search on this workspace
        }
      }
    }
  }
  // ---- chains, then candidate pairs / triples ----
  if (K >= 2048) chain_entries<ZT, true>(z, E, K, D, Dz, chain_list, n_chain, idx, bv, bi);
  else chain_entries<ZT, false>(z, E, K, D, Dz, chain_list, n_chain, idx, bv, bi);
  pair_entries<ZT, kPairLanes>(z, E, D, Dz, pair_list, n_pair, idx);
  if (stats && blockIdx.x == 0 && threadIdx.x == 0) {
    atomicAdd(stats + G2V_STAT_PAIR_RECHECK, (unsigned long long)(n_pair + n_chain));
    atomicAdd(stats + G2V_STAT_FALLBACK_ROWS, (unsigned long long)counters[1]);
    if (n_full > 0) atomicAdd(stats + G2V_STAT_FULL_RECHECK, (unsigned long long)n_full);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
inline size_t al256(size_t v) { return (v + 255) / 256 * 256; }

// Experiment knobs, read from the environment ONCE per process (-1 = not set).  None of them changes results.
struct Tuning {
  int tmem_mode, a_bufs, bstages, zslots, fused, cg, tmem16, ahead;
  unsigned dbg_flags;          // -DG2V_TRACE=1 builds only
  long long* trace;            // -DG2V_TRACE=1 builds only: device buffer the role timeline is written to
};
const Tuning& tuning() {
  static const Tuning t = [] {
    auto geti = [](const char* name) { const char* e = getenv(name); return e ? atoi(e) : -1; };
    Tuning u;
    u.tmem_mode = geti("G2V_TC_TMEM");
    u.a_bufs = geti("G2V_TC_ABUFS");
    u.bstages = geti("G2V_TC_BSTAGES");
    u.zslots = geti("G2V_TC_ZSLOTS");
    u.fused = geti("G2V_TC_FUSED");
    u.cg = geti("G2V_TC_CG");
    u.tmem16 = geti("G2V_TC_TMEM16");
    u.ahead = geti("G2V_TC_AHEAD");
    u.dbg_flags = 0;
    u.trace = nullptr;
#if G2V_TRACE
    if (const char* e = getenv("G2V_TC_DEBUG")) u.dbg_flags = ((unsigned)atoi(e) & 63u) << 16;
    if (const char* e = getenv("G2V_TC_TRACE")) u.trace = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
#endif
    return u;
  }();
  return t;
}

// one side stream + fork / join events per (host thread, device), created on first use and kept
struct SideStream {
  cudaStream_t s;
  cudaEvent_t fork, join;
};
SideStream* side_stream() {
  constexpr int kMaxDev = 64;
  thread_local SideStream tab[kMaxDev];
  thread_local bool made[kMaxDev] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return nullptr;
  if (!made[dev]) {
    SideStream t;
    if (cudaStreamCreateWithFlags(&t.s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&t.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&t.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    tab[dev] = t;
    made[dev] = true;
  }
  return &tab[dev];
}

struct TcWs {
  size_t z16, rowinfo, pairs, fulls, chains, counters, scratch, refine, total;
  RefineWs rf;
};
TcWs tc_ws(int64_t N, int K, int D) {
  const int Dp = round_up(D, 16);
  TcWs w;
  w.z16 = 0;
  w.rowinfo = al256((size_t)N * Dp * 2);
  w.pairs = w.rowinfo + al256((size_t)N * sizeof(RowInfo));
  w.fulls = w.pairs + al256((size_t)N * 16);
  w.chains = w.fulls + al256((size_t)N * 4);
  w.counters = w.chains + al256((size_t)N * 16);       // 256 bytes of list counters, then the re-rank arrive counters
  w.scratch = w.counters + 256 + al256(kRerankArriveBytes);
  w.refine = w.scratch + al256(kRerankScratchBytes);
  w.rf = refine_ws(N, K, D);
  w.total = w.refine + w.rf.total;
  return w;
}

// exact re-rank of the rows the fast pass could not certify (candidate list, chain list, whole rows).
// `counters` is followed by the arrive counters (zero on entry, left zero) and `scratch` by tc_ws().
template <typename ZT>
int run_recheck(const ZT* z, int z_dtype, const float* E, const void* cb, int64_t N, int K, int D, int Dz, int* pairs, int* chains,
                int* fulls, int* counters, void* scratch, int32_t* idx, unsigned long long* stats, unsigned flags,
                cudaStream_t st, const RefineWs& rf, char* rbase) {
  if (flags & G2V_NO_RECHECK) return G2V_OK;
  // The refine pass of the listed whole rows runs on a side stream next to the pair / chain re-rank (they write
  // disjoint rows of idx): three small dependent launches that would otherwise sit in front of the main kernel.
  const bool refine = rf.cap > 0 && !(flags & G2V_NO_REFINE);
  SideStream* side = nullptr;
  if (refine) {
    side = side_stream();
    if (!side) return G2V_ERR_CUDA;
    G2V_CUDA_CHECK(cudaEventRecord(side->fork, st));
    G2V_CUDA_CHECK(cudaStreamWaitEvent(side->s, side->fork, 0));
    float* invt = reinterpret_cast<float*>(rbase + rf.invt);
    RefineRow* rows = reinterpret_cast<RefineRow*>(rbase + rf.rows);
    __half* A16 = reinterpret_cast<__half*>(rbase + rf.A);
    __half* B16 = reinterpret_cast<__half*>(rbase + rf.B);
    float* dots = reinterpret_cast<float*>(rbase + rf.C);
    // codes + a guess of the listed rows (a few per cent of N); the grid-stride loop takes whatever is listed
    const long long tasks = (long long)K + std::min<long long>(rf.cap, N / 16 + 1024);
    const int pgrid = (int)std::max<long long>(1, std::min<long long>((tasks + 7) / 8, (long long)num_sms() * 8));
    int* n_eff = counters + 8;              // inside the 256 bytes of list counters that every search zeroes
    refine_prep_kernel<ZT><<<pgrid, 256, 0, side->s>>>(z, Dz, E, K, D, rf.kp, fulls, counters + 1, (int)rf.cap, A16, B16, invt, rows, n_eff);
    G2V_LAUNCH_CHECK("refine_prep_kernel");
    const int rc = launch_gemm_prepared(A16, n_eff, rf.cap, B16, K, (long long)3 * rf.kp, dots, rf.ldc, side->s);
    if (rc) return rc;
    RefineArgs RA;
    RA.dots = dots; RA.invt = invt; RA.rows = rows;
    RA.e2 = reinterpret_cast<const float*>(reinterpret_cast<const char*>(cb) + cb_e2_offset());
    RA.ldc = rf.ldc; RA.cap = (int)rf.cap;
    RA.beta = 7.3e-7f + 1.003f * (1.9073486e-6f + (float)(3 * rf.kp / 16 + 2) * 1.1920929e-7f);
    if (K <= 512) refine_rows_reg_kernel<ZT, 16><<<num_sms() * 4, 256, 0, side->s>>>(z, E, K, D, Dz, fulls, counters + 1, RA, idx, stats);
    else if (K <= 1024) refine_rows_reg_kernel<ZT, 32><<<num_sms() * 4, 256, 0, side->s>>>(z, E, K, D, Dz, fulls, counters + 1, RA, idx, stats);
    else refine_rows_kernel<ZT><<<num_sms() * 4, 256, 0, side->s>>>(z, E, K, D, Dz, fulls, counters + 1, RA, idx, stats);
    G2V_LAUNCH_CHECK("refine_rows_kernel");
    G2V_CUDA_CHECK(cudaEventRecord(side->join, side->s));
  }
  // a whole row costs K x D fp64 operations on one CTA: split it so that a handful of rows does not become
  // the latency of the step (~64 codes per warp pass)
  int slices = K / 256;
  slices = slices < 1 ? 1 : (slices > kRerankMaxSlices ? kRerankMaxSlices : slices);
  if (K >= 96 && slices < 4) slices = 4;
  int* arrive = counters + 64;
  const size_t smem = (size_t)((D + 31) / 32 * 32) * sizeof(float);
  // every list holds at most N entries and a CTA's eight warps take one entry each per pass: a small batch gets
  // a small grid (a fixed 4-CTAs-per-SM grid costs ~25 us of launch + drain for a 128-row step)
  const long long want = (N + 7) / 8, cap = (long long)num_sms() * 4;
  const int rgrid = (int)(want < 1 ? 1 : (want < cap ? want : cap));
  rerank_kernel<ZT><<<rgrid, 256, smem, st>>>(z, E, K, D, Dz, pairs, chains, fulls, counters, slices,
                                                     reinterpret_cast<RerankSlot*>(scratch), arrive, idx, stats, refine ? (int)rf.cap : 0);
  G2V_LAUNCH_CHECK("rerank_kernel");
  const int rc = launch_full_recheck(z, z_dtype, E, cb, K, D, Dz, fulls, counters + 1, N, idx, stats, true, st,
                                     refine ? (int64_t)rf.cap : (int64_t)kFull64Cap);
  if (refine) G2V_CUDA_CHECK(cudaStreamWaitEvent(st, side->join, 0));
  return rc;
}

// "tmem" variant: geometry, or false if the shape does not qualify (fp32 rows read through TMA: D % 4 == 0
// and a 16-byte aligned base; more than one 128-row tile so that a CTA pair has work).
bool plan_tmem(const void* z, int z_dtype, int64_t N, int K, int D, int Dz, int a_bufs, TmeParams* R) {
  const int Dp = round_up(D, 16);
  // rows are read through TMA: 16-byte aligned base and row pitch (G2V_TC_TMEM16=0 keeps 16-bit rows on the
  // row_prep + shared-memory-operand path)
  if (z_dtype != G2V_F32 && tuning().tmem16 == 0) return false;
  if (Dz % (z_dtype == G2V_F32 ? 4 : 8) != 0 || (reinterpret_cast<uintptr_t>(z) & 15) != 0) return false;
  if (N <= TM || Dp > kMaxDp || Dp < KC) return false;
  // a_bufs == 2: the fp16 rows of the next tile are converted while the MMAs still read this one's; what is
  // left of tensor memory holds two (then much narrower) accumulator stages
  const int acc_col0 = round_up(a_bufs * (Dp / 2), 16);
  if (512 - acc_col0 < 2 * 32) return false;
  const int nt_max = std::min(256, ((512 - acc_col0) / 2) & ~15);
  int best_nt = 0, best_n = 1 << 30, best_pad = 1 << 30, best_last = 0;
  for (int nt = nt_max; nt >= 32 && nt >= nt_max - 64; nt -= 16) {
    const int n = (K + nt - 1) / nt;
    const int last = round_up(K - (n - 1) * nt, 16);
    if (last > nt) continue;
    const int pad = (n - 1) * nt + last;
    if (n < best_n || (n == best_n && pad < best_pad)) { best_nt = nt; best_n = n; best_pad = pad; best_last = last; }
  }
  if (best_nt == 0) return false;
  R->N = N; R->K = K; R->D = D; R->Dp = Dp; R->Dz = Dz;
  R->n_full = Dp / KC; R->n_tail = (Dp % KC) / KT; R->n_chunks = R->n_full + R->n_tail;
  if (R->n_chunks > MAX_CHUNKS) return false;
  R->ntile = best_nt; R->n_ntiles = best_n; R->n_last = best_last; R->Kpad = best_pad;
  R->a_bufs = a_bufs; R->acc_col0 = acc_col0;
  R->ahead = tuning().ahead == 0 ? 0 : 1;      // G2V_TC_AHEAD=0: convert only into a free A buffer (round-1 behaviour)
  R->n_ksteps = Dp / KT;
  R->b_stage = (uint32_t)round_up((best_nt / 2) * (std::min(2, R->n_full) * KC + R->n_tail * KT) * 2, 1024);
  R->n_row_tiles = (int)((N + 2 * TM - 1) / (2 * TM));
  // codebook ring: ~1.5 us of L2 latency at 288 MMA cycles per full panel of 144 codes (~105 KB in flight,
  // whatever the stage size); the rest goes to the row slots
  int nb = std::max(5, std::min(TME_MAX_BST, (int)((5u * 21504u + R->b_stage - 1) / R->b_stage)));
  if (tuning().bstages >= 0) nb = std::max(2, std::min(TME_MAX_BST, tuning().bstages));
  int nz = TME_MAX_ZSLOTS;
  if (tuning().zslots >= 0) nz = std::max(2, std::min(TME_MAX_ZSLOTS, tuning().zslots));
  while (nz > 2 && tme_plan(nb, R->b_stage, nz, R->Kpad).total + 1024 > 227 * 1024) --nz;
  R->nb = nb; R->nz = nz;
  return tme_plan(nb, R->b_stage, nz, R->Kpad).total + 1024 <= 227 * 1024;
}

template <typename ZT>
int launch_tmem(TmeParams& R, const ZT* z, const __half* e16, int Kp, cudaStream_t st) {
  alignas(64) CUtensorMap tmZ, tmZt, tmB, tmBt, tmBl, tmBlt;
  int rc;
  constexpr bool z32 = sizeof(ZT) == 4;
  if ((rc = make_map(&tmZ, z, (uint64_t)R.N, (uint64_t)R.Dz, z32 ? TME_ZCOLS : KC, TM, CU_TENSOR_MAP_SWIZZLE_128B, z32))) return rc;
  if ((rc = make_map(&tmZt, z, (uint64_t)R.N, (uint64_t)R.Dz, KT, TM, CU_TENSOR_MAP_SWIZZLE_NONE, z32))) return rc;
  const uint32_t brow = (uint32_t)R.ntile / 2, blast = (uint32_t)R.n_last / 2;
  if ((rc = make_map(&tmB, e16, (uint64_t)Kp, (uint64_t)R.Dp, KC, brow, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_map(&tmBt, e16, (uint64_t)Kp, (uint64_t)R.Dp, KT, brow, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;
  if ((rc = make_map(&tmBl, e16, (uint64_t)Kp, (uint64_t)R.Dp, KC, blast, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_map(&tmBlt, e16, (uint64_t)Kp, (uint64_t)R.Dp, KT, blast, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;
  const size_t smem = tme_plan(R.nb, R.b_stage, R.nz, R.Kpad).total + 1024;
  { const int rc_attr = set_dyn_smem(reinterpret_cast<const void*>(&tc_tmem_kernel<ZT>), smem); if (rc_attr) return rc_attr; }
  const int max_groups = num_sms() / 2;
  const int groups = R.n_row_tiles < max_groups ? R.n_row_tiles : max_groups;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(groups * 2);
  cfg.blockDim = dim3(TME_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  G2V_CUDA_CHECK(cudaLaunchKernelEx(&cfg, tc_tmem_kernel<ZT>, tmZ, tmZt, tmB, tmBt, tmBl, tmBlt, R));
  G2V_LAUNCH_CHECK("tc_tmem_kernel");
  return G2V_OK;
}

template <typename ZT>
int run_tc(const ZT* z, int z_dtype, const float* E, const void* cb, int64_t N, int K, int D, int Dz, int32_t* idx,
           unsigned long long* stats, void* ws, unsigned flags, cudaStream_t st) {
  const int Dp = round_up(D, 16), Kp = round_up(K, 256);
  const unsigned variant = flags & G2V_TC_VARIANT_MASK;      // test / benchmark aid: pin the sweep kernel
  const TcWs w = tc_ws(N, K, D);
  char* base = reinterpret_cast<char*>(ws);
  __half* z16 = reinterpret_cast<__half*>(base + w.z16);
  RowInfo* rowinfo = reinterpret_cast<RowInfo*>(base + w.rowinfo);
  int* pairs = reinterpret_cast<int*>(base + w.pairs);
  int* fulls = reinterpret_cast<int*>(base + w.fulls);
  int* chains = reinterpret_cast<int*>(base + w.chains);
  int* counters = reinterpret_cast<int*>(base + w.counters);
  const auto* hdr = reinterpret_cast<const CbHeader*>(cb);
  const float* e2 = reinterpret_cast<const float*>(reinterpret_cast<const char*>(cb) + cb_e2_offset());
  const __half* e16 = reinterpret_cast<const __half*>(reinterpret_cast<const char*>(cb) + cb_e16_offset(K));

  {   // fp32 rows, few code tiles: A operand in tensor memory, CTA pairs (G2V_TC_TMEM=0 switches the variant off,
      // =2 forces it for any K)
    TmeParams R;
    int mode = tuning().tmem_mode >= 0 ? tuning().tmem_mode : 1;
    if (variant == G2V_TC_VARIANT_TMEM) mode = 2;
    else if (variant != G2V_TC_VARIANT_AUTO) mode = 0;
    // fp32 rows: any K (measured, 1 M rows: K=2048 1.77 vs 2.18 ms for row_prep + tc_search_kernel, K=16384 equal);
    // 16-bit rows: up to 16 code tiles, beyond that the cheaper 16-bit row_prep + 256-code stages win
    bool use = mode != 0 && plan_tmem(z, z_dtype, N, K, D, Dz, 1, &R) &&
               (mode == 2 || R.n_ntiles <= (z_dtype == G2V_F32 ? (1 << 30) : 16));
    if (use) {      // G2V_TC_ABUFS=1|2: single / double A operand buffer
      int a_bufs = kTmeDefaultABufs;
      if (tuning().a_bufs >= 0) a_bufs = tuning().a_bufs == 2 ? 2 : 1;
      TmeParams R2;
      if (a_bufs == 2 && plan_tmem(z, z_dtype, N, K, D, Dz, 2, &R2)) R = R2;
    }
    if (use) {
      R.hdr = hdr; R.e2 = e2; R.idx = idx;
      R.ntab = reinterpret_cast<const float*>(reinterpret_cast<const char*>(cb) + cb_tab_offset());
      R.pair_list = pairs; R.full_list = fulls; R.chain_list = chains; R.counters = counters; R.flags = flags;
      R.flags |= tuning().dbg_flags;
      R.trace = tuning().trace;
      G2V_CUDA_CHECK(cudaMemsetAsync(counters, 0, 256 + kRerankArriveBytes, st));
      cudaEvent_t pev0, pev1;
      profile_take(&pev0, &pev1);
      if (pev0) G2V_CUDA_CHECK(cudaEventRecord(pev0, st));
      const int rc = launch_tmem<ZT>(R, z, e16, Kp, st);
      if (rc) return rc;
      if (pev1) G2V_CUDA_CHECK(cudaEventRecord(pev1, st));
      return run_recheck(z, z_dtype, E, cb, N, K, D, Dz, pairs, chains, fulls, counters, base + w.scratch, idx, stats, flags, st, w.rf, base + w.refine);
    }
  }
  if (Dz != D) {
    set_error_detail("rows narrower than the codebook (Dz=%d, D=%d) are only searched by tc_tmem_kernel", Dz, D);
    return G2V_ERR_UNSUPPORTED;
  }
  TcParams P;
  P.N = N; P.K = K; P.D = D; P.Dp = Dp;
  P.n_full = Dp / KC;
  P.n_tail = (Dp % KC) / KT;
  P.n_chunks = P.n_full + P.n_tail;
  // Two regimes (measured, see profiles/):
  //  * K <= 512, fp32 rows: HBM-bound.  One kernel reads the fp32 rows once and converts them on the
  //    fly ("fused"); single-CTA MMAs, because a CTA pair needs a cluster-scope release per operand
  //    panel from the peer's converter warps (~2 us of latency per row tile).
  //  * larger K or 16-bit rows: tensor-bound.  row_prep_kernel + CTA pairs (cta_group::2: M = 256 per
  //    MMA, each CTA streams half of every codebook stage through a deep ring).
  // G2V_TC_CG=1|2 and G2V_TC_FUSED=0|1 override the choice.
  P.n_ntiles = (K + TN - 1) / TN;
  bool fused = (z_dtype == G2V_F32) && (D % 4 == 0) && (D >= KC) && P.n_ntiles <= 2 &&
               ((reinterpret_cast<uintptr_t>(z) & 15) == 0);
  const bool fusable = (z_dtype == G2V_F32) && (D % 4 == 0) && (D >= KC) && ((reinterpret_cast<uintptr_t>(z) & 15) == 0);
  if (variant == G2V_TC_VARIANT_FUSED) fused = fusable;
  else if (variant == G2V_TC_VARIANT_PREP) fused = false;
  else if (tuning().fused >= 0) {
    if (tuning().fused == 0) fused = false;
    else fused = (z_dtype == G2V_F32) && (D % 4 == 0) && (D >= KC) && ((reinterpret_cast<uintptr_t>(z) & 15) == 0);
  }
  int cg = fused ? 1 : ((N > TM) ? 2 : 1);
  if (tuning().cg >= 0) cg = (tuning().cg == 1) ? 1 : ((N > TM) ? 2 : 1);
  if (fused && smem_plan(P.n_full, P.n_tail, cg, 2, 3).total + 1024 > 227 * 1024) fused = false;   // needs >= 3 staging slots
  P.n_last_mma = round_up(K - (P.n_ntiles - 1) * TN, 16 * cg);
  P.n_row_tiles = (int)((N + (long long)TM * cg - 1) / ((long long)TM * cg));   // tiles of 128*cg rows
  P.rowinfo = rowinfo; P.e2 = e2; P.idx = idx;
  P.ntab = reinterpret_cast<const float*>(reinterpret_cast<const char*>(cb) + cb_tab_offset());
  P.pair_list = pairs; P.full_list = fulls; P.chain_list = chains; P.counters = counters; P.flags = flags;
  P.trace = tuning().trace;
  P.flags |= tuning().dbg_flags;

  const int n_ksteps = P.n_full * (KC / KT) + P.n_tail;
  P.n_ksteps = n_ksteps;
  P.hdr = hdr;
  if (fused) {
    G2V_CUDA_CHECK(cudaMemsetAsync(counters, 0, 256 + kRerankArriveBytes, st));
  } else {
    G2V_CUDA_CHECK(cudaMemsetAsync(counters, 0, 256 + kRerankArriveBytes, st));
    const long long threads = N * 32;
    const int grid = (int)((threads + 255) / 256);
    row_prep_kernel<ZT><<<grid, 256, 0, st>>>(z, N, D, Dp, hdr, z16, rowinfo, counters, n_ksteps);
    G2V_LAUNCH_CHECK("row_prep_kernel");
  }

  alignas(64) CUtensorMap tmA, tmAt, tmB, tmBt, tmBl, tmBlt, tmPf;
  int rc;
  // main maps need a 64-wide box; if D < 64 there are no full panels and the main maps are unused
  const uint32_t main_box = (Dp >= KC) ? KC : KT;
  const CUtensorMapSwizzle main_sw = (Dp >= KC) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B;
  if (fused) {
    if ((rc = make_map(&tmA, z, (uint64_t)N, (uint64_t)D, KC, ZROWS, CU_TENSOR_MAP_SWIZZLE_NONE, true))) return rc;
    if ((rc = make_map(&tmAt, z, (uint64_t)N, (uint64_t)D, KT, TM, CU_TENSOR_MAP_SWIZZLE_NONE, true))) return rc;
    if ((rc = make_map(&tmPf, z, (uint64_t)N, (uint64_t)D, KC, TM, CU_TENSOR_MAP_SWIZZLE_NONE, true))) return rc;
  } else {
    if ((rc = make_map(&tmA, z16, (uint64_t)N, (uint64_t)Dp, main_box, TM, main_sw))) return rc;
    if ((rc = make_map(&tmPf, z16, (uint64_t)N, (uint64_t)Dp, Dp < 64 ? (uint32_t)Dp : 64u, TM, CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
  }
  if (!fused && (rc = make_map(&tmAt, z16, (uint64_t)N, (uint64_t)Dp, KT, TM, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;
  const uint32_t brow = TN / cg, blast = (uint32_t)P.n_last_mma / cg;     // code rows each CTA fetches per stage
  if ((rc = make_map(&tmB, e16, (uint64_t)Kp, (uint64_t)Dp, main_box, brow, main_sw))) return rc;
  if ((rc = make_map(&tmBt, e16, (uint64_t)Kp, (uint64_t)Dp, KT, brow, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;
  if ((rc = make_map(&tmBl, e16, (uint64_t)Kp, (uint64_t)Dp, main_box, blast, main_sw))) return rc;
  if ((rc = make_map(&tmBlt, e16, (uint64_t)Kp, (uint64_t)Dp, KT, blast, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;

  // ring depth: whatever shared memory is left after the resident row tile, capped at MAX_STAGES
  // (the ring has to cover the TMA round trip, ~2.5k cycles, at 512 MMA cycles per full panel)
  int nstage = fused ? (cg == 2 ? 4 : 2) : MAX_STAGES, nzslot = 0;
  if (tuning().bstages >= 2 && tuning().bstages <= MAX_STAGES) nstage = tuning().bstages;
  while (nstage > 2 && smem_plan(P.n_full, P.n_tail, cg, nstage).total + 1024 > 227 * 1024) --nstage;
  if (fused) {      // the rest of shared memory becomes the fp32 staging ring (>= 3 slots or give up fusing)
    nzslot = MAX_ZSLOTS;
    if (tuning().zslots >= 2 && tuning().zslots <= MAX_ZSLOTS) nzslot = tuning().zslots;
    while (nzslot > 0 && smem_plan(P.n_full, P.n_tail, cg, nstage, nzslot).total + 1024 > 227 * 1024) --nzslot;
    if (nzslot < 2) return G2V_ERR_UNSUPPORTED;   // cannot happen: checked when `fused` was decided
  }
  P.nstage = nstage;
  P.nzslot = nzslot;
  const SmemPlan sp = smem_plan(P.n_full, P.n_tail, cg, nstage, nzslot);
  const size_t smem = sp.total + 1024;
  cudaEvent_t pev0, pev1;
  profile_take(&pev0, &pev1);
  if (pev0) G2V_CUDA_CHECK(cudaEventRecord(pev0, st));
  const int max_groups = num_sms() / cg;
  const int groups = P.n_row_tiles < max_groups ? P.n_row_tiles : max_groups;
  if (cg == 1) {
    if (fused) {
      { const int rc_attr = set_dyn_smem(reinterpret_cast<const void*>(&tc_search_kernel<1, true>), smem); if (rc_attr) return rc_attr; }
      tc_search_kernel<1, true><<<groups, NTHREADS_FUSED, smem, st>>>(tmA, tmAt, tmB, tmBt, tmBl, tmBlt, tmPf, P);
    } else {
      { const int rc_attr = set_dyn_smem(reinterpret_cast<const void*>(&tc_search_kernel<1, false>), smem); if (rc_attr) return rc_attr; }
      tc_search_kernel<1, false><<<groups, NTHREADS, smem, st>>>(tmA, tmAt, tmB, tmBt, tmBl, tmBlt, tmPf, P);
    }
  } else {
    if (fused)
      { const int rc_attr = set_dyn_smem(reinterpret_cast<const void*>(&tc_search_kernel<2, true>), smem); if (rc_attr) return rc_attr; }
    else
      { const int rc_attr = set_dyn_smem(reinterpret_cast<const void*>(&tc_search_kernel<2, false>), smem); if (rc_attr) return rc_attr; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(groups * 2);
    cfg.blockDim = dim3(fused ? NTHREADS_FUSED : NTHREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (fused) G2V_CUDA_CHECK(cudaLaunchKernelEx(&cfg, tc_search_kernel<2, true>, tmA, tmAt, tmB, tmBt, tmBl, tmBlt, tmPf, P));
    else G2V_CUDA_CHECK(cudaLaunchKernelEx(&cfg, tc_search_kernel<2, false>, tmA, tmAt, tmB, tmBt, tmBl, tmBlt, tmPf, P));
  }
  G2V_LAUNCH_CHECK("tc_search_kernel");
  if (pev1) G2V_CUDA_CHECK(cudaEventRecord(pev1, st));

  return run_recheck(z, z_dtype, E, cb, N, K, D, Dz, pairs, chains, fulls, counters, base + w.scratch, idx, stats, flags, st, w.rf, base + w.refine);
}

}  // namespace

bool tc_supported(int K, int D) {
  const int Dp = round_up(D, 16);
  if (Dp > kMaxDp || K > kMaxK) return false;
  if ((long long)K * D < 16384) return false;       // tiny problems: the fp32 path is already bandwidth-bound
  const SmemPlan sp = smem_plan(Dp / KC, (Dp % KC) / KT, 1, 2);
  return sp.total + 1024 <= 227 * 1024;
}

size_t tc_workspace_bytes(int64_t N, int K, int D, int z_dtype) {
  (void)K; (void)z_dtype;
  return tc_ws(N, K, D).total;
}

int launch_search_tc(const void* z, int z_dtype, const float* E, const void* cb, int64_t N, int K, int D,
                     int32_t* idx, unsigned long long* stats, void* ws, size_t ws_bytes, unsigned flags,
                     cudaStream_t st, int Dz) {
  if (Dz <= 0) Dz = D;
  if (!tc_supported(K, D)) return G2V_ERR_UNSUPPORTED;
  if (ws_bytes < tc_ws(N, K, D).total) return G2V_ERR_WORKSPACE;
  switch (z_dtype) {
    case G2V_F32: return run_tc(reinterpret_cast<const float*>(z), z_dtype, E, cb, N, K, D, Dz, idx, stats, ws, flags, st);
    case G2V_F16: return run_tc(reinterpret_cast<const __half*>(z), z_dtype, E, cb, N, K, D, Dz, idx, stats, ws, flags, st);
    case G2V_BF16: return run_tc(reinterpret_cast<const __nv_bfloat16*>(z), z_dtype, E, cb, N, K, D, Dz, idx, stats, ws, flags, st);
    default: return G2V_ERR_DTYPE;
  }
}

}  // namespace g2v
