// tc_pair_kernel: tc_tmem_kernel with a DOUBLE-BUFFERED A operand.  Included by g2v_tc.cu inside
// namespace g2v { namespace { ... } } (it uses that file's PTX wrappers, epilogue and TmeParams).
//
// Why: with one A buffer the per-tile dependency cycle conv -> MMA(first code tile) -> MMA -> MMA -> conv leaves
// the tensor pipe idle while the next tile's rows are converted, and the row stream idle while the converters
// wait for the buffer (profiles/r1_trace_tc_tmem_k400.txt: 8.4 us per 256-row tile for 5.1 us of MMAs).  Tensor
// memory cannot hold two A buffers next to two useful accumulator stages (2 x 200 + 2 x 144 columns > 512), so the
// second buffer lives in shared memory:
//
//   even row tiles   A in tensor memory, columns [0, Dp/2)   (tcgen05.st.16x256b by the converters, MMA kind ".ts")
//   odd  row tiles   A in shared memory, SWIZZLE_128B K-major panels (st.shared.v2 by the same converter
//                    registers + fence.proxy.async, MMA reads it through a shared-memory descriptor)
//
// The converters of tile t+1 therefore run while the MMAs of tile t execute, and neither waits for the other
// except through the per-super-chunk "buffer free" / "panel converted" barriers of tile t-1 / t+1.
// Shared memory per CTA: 100 KB A buffer + 4 codebook stages of ONE 64-column panel each (+ the 16-column tails
// with the last panel) + 4 row slots of 16 KB; the row tiles are pulled into L2 two tiles ahead by TMA prefetches,
// so the short row ring only has to cover L2 latency, not HBM latency.
constexpr int PAIR_MAX_BST = 8;

struct PairPlan {
  uint32_t a_off, b_off, z_off, e2_off, xch_off, rs_off, bar_off, tmem_off, total;
};
constexpr int PAIR_NBARS = 2 * PAIR_MAX_BST + 2 * TME_MAX_ZSLOTS + 6 * MAX_CHUNKS + 4 + RS_RING;
__host__ __device__ inline PairPlan pair_plan(int n_full, int n_tail, int nb, uint32_t b_stage, int nz, int Kpad) {
  PairPlan p;
  p.a_off = 0;
  p.b_off = (uint32_t)n_full * A_PANEL + (uint32_t)n_tail * A_TAIL;           // 16 KB / 4 KB units: 1024-aligned
  p.z_off = p.b_off + (uint32_t)nb * b_stage;
  p.e2_off = p.z_off + (uint32_t)nz * TME_ZSLOT;
  p.xch_off = p.e2_off + (((uint32_t)Kpad * 4u + 127u) & ~127u);
  p.rs_off = p.xch_off + TM * 12 * 4;
  p.bar_off = p.rs_off + RS_RING * TM * 8;
  p.tmem_off = p.bar_off + 8 * PAIR_NBARS;
  p.total = p.tmem_off + 16;
  return p;
}

__device__ __forceinline__ void sts_v2(uint32_t saddr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(saddr), "r"(a), "r"(b) : "memory");
}

template <typename ZT>
__global__ void __launch_bounds__(TME_THREADS, 1)
tc_pair_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmZt,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBt,
               const __grid_constant__ CUtensorMap tmBl, const __grid_constant__ CUtensorMap tmBlt, const TmeParams P) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* gbase = smem_dyn + (base - raw);
  const PairPlan sp = pair_plan(P.n_full, P.n_tail, P.nb, P.b_stage, P.nz, P.Kpad);
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int n_groups = gridDim.x / 2, group = blockIdx.x / 2;
  const uint32_t NB = (uint32_t)P.nb, NZ = (uint32_t)P.nz;

  const uint32_t sA = base + sp.a_off, sB = base + sp.b_off, sZ = base + sp.z_off;
  float* e2s = reinterpret_cast<float*>(gbase + sp.e2_off);
  uint32_t* xch = reinterpret_cast<uint32_t*>(gbase + sp.xch_off);
  float2* rowstat = reinterpret_cast<float2*>(gbase + sp.rs_off);
  const uint32_t bars = base + sp.bar_off;
  auto bar_bfull = [&](int s) { return bars + 8u * s; };
  auto bar_bempty = [&](int s) { return bars + 8u * (PAIR_MAX_BST + s); };
  auto bar_zfull = [&](int s) { return bars + 8u * (2 * PAIR_MAX_BST + s); };
  auto bar_zempty = [&](int s) { return bars + 8u * (2 * PAIR_MAX_BST + TME_MAX_ZSLOTS + s); };
  constexpr int A0 = 2 * PAIR_MAX_BST + 2 * TME_MAX_ZSLOTS;
  // per A buffer b (0 = tensor memory, 1 = shared memory) and super-chunk c (two 64-column panels; tails with the last)
  auto bar_aconv = [&](int b, int c) { return bars + 8u * (A0 + b * MAX_CHUNKS + c); };              // this CTA's converters wrote it
  auto bar_apeer = [&](int b, int c) { return bars + 8u * (A0 + (2 + b) * MAX_CHUNKS + c); };        // leader: the peer's converters did
  auto bar_aempty = [&](int b, int c) { return bars + 8u * (A0 + (4 + b) * MAX_CHUNKS + c); };       // the MMAs finished reading it
  auto bar_accfull = [&](int a) { return bars + 8u * (A0 + 6 * MAX_CHUNKS + a); };
  auto bar_accempty = [&](int a) { return bars + 8u * (A0 + 6 * MAX_CHUNKS + 2 + a); };
  auto bar_rsfull = [&](int s) { return bars + 8u * (A0 + 6 * MAX_CHUNKS + 4 + s); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + sp.tmem_off);

  constexpr bool Z32 = sizeof(ZT) == 4;     // fp32 rows; otherwise bf16 / fp16 rows
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_chunks = P.n_chunks, n_full = P.n_full;
  auto acc_col_of = [&](uint32_t a) { return (uint32_t)P.acc_col0 + a * (uint32_t)P.ntile; };   // accumulator stage a
  const int n_sc = (n_full + 1) >> 1;      // super-chunks (A hand-overs) per row tile
  // shared-memory A buffer: full panel c at sA + c * A_PANEL, tail t at sA + n_full * A_PANEL + t * A_TAIL
  const uint32_t sAt = sA + (uint32_t)n_full * A_PANEL;

  if (threadIdx.x == 0) {
    for (int s = 0; s < PAIR_MAX_BST; ++s) { mbar_init(bar_bfull(s), 1); mbar_init(bar_bempty(s), 1); }
    for (int s = 0; s < TME_MAX_ZSLOTS; ++s) { mbar_init(bar_zfull(s), 1); mbar_init(bar_zempty(s), TME_CONV_WARPS); }
    for (int b = 0; b < 2; ++b)
      for (int c = 0; c < MAX_CHUNKS; ++c) { mbar_init(bar_aconv(b, c), TME_CONV_WARPS); mbar_init(bar_apeer(b, c), 1); mbar_init(bar_aempty(b, c), 1); }
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

  if (warp == 0) {
    // =========================== codebook producer (each CTA fetches its half of every stage) ===========================
    // a stage = ONE 64-column panel of a code tile; the 16-column tails ride with the last panel
    uint32_t s = 0, sph = 0;     // ring position and its phase
    for (int tile = group; tile < P.n_row_tiles; tile += n_groups) {
      for (int nt = 0; nt < P.n_ntiles; ++nt) {
        const bool last = (nt == P.n_ntiles - 1);
        const uint32_t rows = (uint32_t)(last ? P.n_last : P.ntile);        // codes of the tile; this CTA fetches half
        const int row = nt * P.ntile + (int)(cta_rank * (rows / 2));
        for (int c = 0; c < n_full; ++c) {
          mbar_wait(bar_bempty(s), sph ^ 1u);
          const int nt_here = (c == n_full - 1) ? P.n_tail : 0;
          if (elect_one()) {
            if (leader) mbar_expect_tx(bar_bfull(s), rows * (uint32_t)(KC + nt_here * KT) * 2u);   // bytes of both CTAs
            uint32_t dst = sB + s * P.b_stage;
            tma_load_2d<2>(dst, last ? &tmBl : &tmB, c * KC, row, bar_bfull(s));
            dst += (rows / 2) * 128u;
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
      for (int tile = group; tile < P.n_row_tiles; tile += n_groups, ++ti) {
        const int ab = (int)(ti & 1u);
        const uint32_t apar = (ti >> 1) & 1u;
        for (int nt = 0; nt < P.n_ntiles; ++nt, ++it) {
          const uint32_t as = it & 1u, around = it >> 1;
          const bool last_nt = (nt == P.n_ntiles - 1);
          const uint32_t idesc = umma_idesc(2 * TM, last_nt ? P.n_last : P.ntile);
          mbar_wait(bar_accempty(as), (around & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc_col_of(as);
          const uint32_t rows_b = (uint32_t)(last_nt ? P.n_last : P.ntile) / 2u;     // code rows per CTA in a stage
          for (int c = 0; c < n_full; ++c) {
            const int sc = c >> 1;
            if (nt == 0 && (c & 1) == 0) {
              mbar_wait(bar_aconv(ab, sc), apar);
              mbar_wait_cluster(bar_apeer(ab, sc), apar);
            }
            mbar_wait(bar_bfull(s), sph);
            tc_fence_after();
            const bool last_panel = (c == n_full - 1);
            const uint32_t b_addr = sB + s * P.b_stage;
            const uint64_t bd0 = umma_desc(b_addr, 1024, 2);
            if (elect_one()) {
              if (ab == 0) {
                const uint32_t a_tmem = tmem_base + 32u * (uint32_t)c;
#pragma unroll
                for (int kk = 0; kk < KC / KT; ++kk)
                  tc_mma_f16_ts2(d_tmem, a_tmem + 8u * kk, bd0 + 2u * kk, idesc, (c | kk) != 0);
                if (last_panel)
                  for (int t = 0; t < P.n_tail; ++t)
                    tc_mma_f16_ts2(d_tmem, tmem_base + 32u * n_full + 8u * t,
                                   umma_desc(b_addr + rows_b * 128u + (uint32_t)t * rows_b * 32u, 256, 6), idesc, 1u);
              } else {
                const uint64_t ad0 = umma_desc(sA + (uint32_t)c * A_PANEL, 1024, 2);
#pragma unroll
                for (int kk = 0; kk < KC / KT; ++kk)
                  tc_mma_f16<2>(d_tmem, ad0 + 2u * kk, bd0 + 2u * kk, idesc, (c | kk) != 0);
                if (last_panel)
                  for (int t = 0; t < P.n_tail; ++t)
                    tc_mma_f16<2>(d_tmem, umma_desc(sAt + (uint32_t)t * A_TAIL, 256, 6),
                                  umma_desc(b_addr + rows_b * 128u + (uint32_t)t * rows_b * 32u, 256, 6), idesc, 1u);
              }
              tc_commit<2>(bar_bempty(s));
              if (last_nt && ((c & 1) == 1 || last_panel)) tc_commit<2>(bar_aempty(ab, sc));
              if (last_panel) tc_commit<2>(bar_accfull(as));
            }
            __syncwarp();
            if (++s == NB) { s = 0; sph ^= 1u; }
          }
        }
      }
    } else {
      // =========================== peer: forward "super-chunk converted" to the leader ===========================
      // (bare remote arrive: the data it orders was published by tcgen05.wait::st / fence.proxy.async before the
      // converters arrived on the local barrier)
      uint32_t ti = 0;
      for (int tile = group; tile < P.n_row_tiles; tile += n_groups, ++ti) {
        for (int sc = 0; sc < n_sc; ++sc) {
          mbar_wait(bar_aconv((int)(ti & 1u), sc), (ti >> 1) & 1u);
          if (elect_one()) mbar_arrive_cluster(bar_apeer((int)(ti & 1u), sc), 0);
          __syncwarp();
        }
      }
    }
  } else if (warp == 2) {
    // =========================== row producer: slots of 128 rows x 128 bytes ===========================
    // The ring is only 4 slots (64 KB): every tile is first pulled into L2 (TMA prefetch, same boxes) kPairPf tiles
    // ahead, so a slot refill costs an L2 hit, not an HBM round trip.
    constexpr int kPairPf = 2;
    const int n_main = (Z32 ? 2 : 1) * n_full, n_slots = n_main + P.n_tail;
    auto prefetch_tile = [&](int t) {
      if (t < P.n_row_tiles && elect_one()) {
        const long long r0 = ((long long)t * 2 + cta_rank) * TM;
        if (r0 < P.N) {
          for (int u = 0; u < n_main; ++u) tma_prefetch_2d(&tmZ, u * (Z32 ? TME_ZCOLS : KC), (int)r0);
          for (int u = n_main; u < n_slots; ++u) tma_prefetch_2d(&tmZt, n_full * KC + (u - n_main) * KT, (int)r0);
        }
      }
      __syncwarp();
    };
    for (int a = 0; a < kPairPf; ++a) prefetch_tile(group + a * n_groups);
    uint32_t slot = 0, ph = 0;
    for (int tile = group; tile < P.n_row_tiles; tile += n_groups) {
      prefetch_tile(tile + kPairPf * n_groups);
      const long long r0 = ((long long)tile * 2 + cta_rank) * TM;
      const int row0 = (int)(r0 < P.N ? r0 : P.N - 1);       // a tile past the end reads (and ignores) the last row
      for (int u = 0; u < n_slots; ++u) {
        mbar_wait(bar_zempty(slot), ph ^ 1u);
        if (elect_one()) {
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
    // =========================== epilogue (as in tc_tmem_kernel) ===========================
    const int q = warp & 3;
    const int eh = (warp - EPI_WARP0) >> 2;
    const int r = q * 32 + lane;
    uint32_t* xrow = xch + (size_t)r * 12;
    uint32_t it = 0, ti = 0;
    const uint32_t e2_saddr = smem_u32(e2s) + 64u * (uint32_t)eh;      // ||e||^2 of code 16 eh
    for (int tile = group; tile < P.n_row_tiles; tile += n_groups, ++ti) {
      mbar_wait(bar_rsfull(ti % RS_RING), (ti / RS_RING) & 1u);
      float key_cS, key_S;
      {
        const RowInfo r0 = make_rowinfo(rowstat[(ti % RS_RING) * TM + r].x, 0.f, 1.f, 0.f, P.hdr->e2min, P.hdr->scale_e, P.n_ksteps);
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
        mbar_wait(bar_accfull(as), around & 1u);
        tc_fence_after();
        const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + acc_col_of(as) + (uint32_t)(16 * eh - s0);
        epi_sweep(taddr0, e2_saddr, g0, g1, e_valid, eh, key_cS, key_S, 0u, false, m1, m2);
        tc_fence_before();
        __syncwarp();
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
        const RowInfo ri = make_rowinfo(st.x, st.y, 1.f, P.hdr->sfrac, P.hdr->e2min, P.hdr->scale_e, P.n_ksteps);
        finish_row(c1, c2, c3, k4, ri, row, valid, P.K, P.ntab, P.flags, P.idx, P.pair_list, P.chain_list, P.full_list, P.counters);
      }
      asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
    }
  } else if (warp >= TME_CONV_WARP0) {
    // =========================== converters ===========================
    // warp cw owns the 16 rows 32 (cw % 4) + 16 (cw / 4) ...; lane i holds, per K step of 16, the elements
    // 4 (i % 4) .. +3 of the rows i / 4 and i / 4 + 8 -- the register layout of tcgen05.st.16x256b, which the
    // shared-memory buffer reuses: the same words go to the SWIZZLE_128B panel as 8-byte stores.
    const int cw = warp - TME_CONV_WARP0;
    const int lane0_row = 32 * (cw & 3) + 16 * (cw >> 2);
    const int ra_l = lane0_row + (lane >> 2), rb_l = ra_l + 8;
    const uint32_t t_lane = tmem_base + ((uint32_t)lane0_row << 16);
    const uint32_t kq = (uint32_t)(lane & 3), sw = (uint32_t)(ra_l & 7);            // (rb_l & 7) == (ra_l & 7)
    const uint32_t off_a0 = (uint32_t)ra_l * 128u + ((kq ^ sw) << 4), off_a1 = (uint32_t)ra_l * 128u + (((4u + kq) ^ sw) << 4);
    const bool odd = (ra_l & 1) != 0;
    const uint32_t off_f = odd ? off_a1 : off_a0, off_s = odd ? off_a0 : off_a1;
    const uint32_t off_t = (uint32_t)ra_l * 64u + kq * 16u;                        // fp32 tail slot: 64-byte rows, no swizzle
    // shared-memory A panel: row r at r * 128, K step s = 16-byte chunks 2s, 2s+1 (xor r % 8); this lane's 8 bytes
    const uint32_t sa_row = (uint32_t)ra_l * 128u + 8u * (kq & 1u);
    uint32_t sa_ks[4];
#pragma unroll
    for (int s4 = 0; s4 < 4; ++s4) sa_ks[s4] = sa_row + (((2u * (uint32_t)s4 + (kq >> 1)) ^ sw) << 4);
    const uint32_t sa_tail = (uint32_t)ra_l * 32u + ((((kq >> 1) & 1u) ^ (((uint32_t)ra_l >> 2) & 1u)) << 4) + 8u * (kq & 1u);
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
    // one full panel's 16 words -> the A buffer (w[8h + 4e + {0,1}]: row a, K step 2h + e; {2,3}: row b)
    auto put_panel = [&](int ab, int c, const uint32_t (&w)[16]) {
      if (ab == 0) {
        tc_st_16x256b_x4(t_lane + 32u * (uint32_t)c, w);
      } else {
        const uint32_t pa = sA + (uint32_t)c * A_PANEL;
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
          sts_v2(pa + sa_ks[s4], w[4 * s4 + 0], w[4 * s4 + 1]);
          sts_v2(pa + sa_ks[s4] + 8u * 128u, w[4 * s4 + 2], w[4 * s4 + 3]);
        }
      }
    };
    auto put_tail = [&](int ab, int t, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
      if (ab == 0) {
        tc_st_16x256b_x1(t_lane + 32u * (uint32_t)n_full + 8u * (uint32_t)t, w0, w1, w2, w3);
      } else {
        const uint32_t pa = sAt + (uint32_t)t * A_TAIL + sa_tail;
        sts_v2(pa, w0, w1);
        sts_v2(pa + 8u * 32u, w2, w3);
      }
    };
    uint32_t slot = 0, ph = 0, ti = 0;
    for (int tile = group; tile < P.n_row_tiles; tile += n_groups, ++ti) {
      const int ab = (int)(ti & 1u);
      for (int c = 0; c < n_chunks; ++c) {
        // barriers work on super-chunks: panels 2 sc and 2 sc + 1, the tails belong to the last one
        const int sc = min(c >> 1, n_sc - 1);
        const bool sc_first = (c < n_full) && ((c & 1) == 0);
        const bool sc_last = (c == n_chunks - 1) || (c < n_full - 1 && (c & 1) == 1) || (c == n_full - 1 && sc < n_sc - 1);
        if (sc_first) {
          mbar_wait(bar_aempty(ab, sc), ((ti >> 1) & 1u) ^ 1u);   // the MMAs of the buffer's previous row tile have read these panels
          tc_fence_after();
        }
        if (!Z32 && c < n_full) {
          // 16-bit rows: one slot per panel, a row is 64 elements = 128 bytes (SWIZZLE_128B); K step s of row r is
          // the 16-byte chunks 2s, 2s+1 (xor r % 8); this lane takes 8 bytes (4 elements) of it
          uint32_t w[16];
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
          put_panel(ab, c, w);
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
          put_tail(ab, c - n_full, w0, w1, w2, w3);
        } else if (c < n_full) {
          uint32_t w[16];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            mbar_wait(bar_zfull(slot), ph);
            const unsigned char* zs = gbase + sp.z_off + slot * TME_ZSLOT;
            // odd rows fetch their second K step first: the two rows of a quarter-warp then sit in different
            // halves of the swizzled 128-byte line (no bank conflict); the words are put back in order below
            const float4 af = *reinterpret_cast<const float4*>(zs + off_f), bf = *reinterpret_cast<const float4*>(zs + off_f + 8 * 128);
            const float4 as = *reinterpret_cast<const float4*>(zs + off_s), bs = *reinterpret_cast<const float4*>(zs + off_s + 8 * 128);
            uint32_t f0, f1, f2, f3, s0, s1, s2, s3;
            cvt(af, z2a, r2a, f0, f1);
            cvt(bf, z2b, r2b, f2, f3);
            cvt(as, z2a, r2a, s0, s1);
            cvt(bs, z2b, r2b, s2, s3);
            w[8 * h + 0] = odd ? s0 : f0; w[8 * h + 1] = odd ? s1 : f1;
            w[8 * h + 2] = odd ? s2 : f2; w[8 * h + 3] = odd ? s3 : f3;
            w[8 * h + 4] = odd ? f0 : s0; w[8 * h + 5] = odd ? f1 : s1;
            w[8 * h + 6] = odd ? f2 : s2; w[8 * h + 7] = odd ? f3 : s3;
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_zempty(slot));
            if (++slot == NZ) { slot = 0; ph ^= 1u; }
          }
          put_panel(ab, c, w);
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
          put_tail(ab, c - n_full, w0, w1, w2, w3);
        }
        if (sc_last) {
          if (ab == 0) {
            tc_wait_st();
            tc_fence_before();
          } else {
            fence_proxy_async();          // generic-proxy stores -> visible to the tensor core's (async proxy) reads
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_aconv(ab, sc));
        }
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
