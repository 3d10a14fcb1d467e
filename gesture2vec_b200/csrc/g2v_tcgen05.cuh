// tcgen05 / TMA / mbarrier PTX wrappers and the tensor-map helpers shared by the sm_100a tensor-core kernels
// (g2v_tc.cu: nearest-code search; g2v_gemm.cu: split-fp16 GEMM).  Everything here is internal to a translation
// unit (anonymous namespace): each .cu that includes it gets its own copies.
#pragma once
#include "g2v_common.cuh"

#include <cuda.h>
#include <stdlib.h>

#include <mutex>

namespace g2v {
namespace {

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// G2V_WAIT_HINT: suspend-time hint (ns) of try_wait -- the warp sleeps in hardware until the phase completes
// or the hint expires, instead of re-polling (each poll costs three issue slots: SYNCS, YIELD, BRA)
#ifndef G2V_WAIT_HINT
#define G2V_WAIT_HINT 0
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#if G2V_WAIT_HINT
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity), "r"((uint32_t)G2V_WAIT_HINT)
      : "memory");
#else
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
#endif
}
// TMA tile load.  CG == 2: the .cta_group::2 form, whose completion bytes are credited to the
// mbarrier at the same offset in the LEADER CTA (peer bit of the shared::cluster address cleared).
// wait with cluster-scope acquire (pairs with a peer CTA's release.cluster arrive)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
#if G2V_WAIT_HINT
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP_C:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra.uni WAIT_DONE_C;\n\t"
      "bra.uni WAIT_LOOP_C;\n\t"
      "WAIT_DONE_C:\n\t"
      "}" ::"r"(bar), "r"(parity), "r"((uint32_t)G2V_WAIT_HINT)
      : "memory");
#else
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP_C:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE_C;\n\t"
      "bra.uni WAIT_LOOP_C;\n\t"
      "WAIT_DONE_C:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
#endif
}
// Wait of a warp that is NOT the one everybody else is waiting for (epilogue / converter / producer warps of
// tc_tmem_kernel).  try_wait comes back at once on this part whatever its time hint says (ncu: 56 M of the 306 M
// warp-instructions of the K=400 sweep are SYNCS / YIELD / BRA polls, 11 M polls by the epilogue warps alone), and
// the polling warps share a scheduler with the warps they wait for.  -DG2V_WAIT_SLEEP_NS=<ns> puts a nanosleep
// behind every failed poll.  Measured (1 M rows, K=400 / 512 / 2048, bf16 512): 20, 40 and 100 ns are all ~1 %
// SLOWER than plain polling (421.5 vs 417.1 us at K=400) -- the polls only use issue slots nobody else wants, so
// the default stays 0.
#ifndef G2V_WAIT_SLEEP_NS
#define G2V_WAIT_SLEEP_NS 0
#endif
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_idle(uint32_t bar, uint32_t parity) {
#if G2V_WAIT_SLEEP_NS > 0
  while (!mbar_try(bar, parity)) asm volatile("nanosleep.u32 %0;" ::"r"((uint32_t)G2V_WAIT_SLEEP_NS));
#else
  mbar_wait(bar, parity);
#endif
}
__device__ __forceinline__ bool mbar_try_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster_idle(uint32_t bar, uint32_t parity) {
#if G2V_WAIT_SLEEP_NS > 0
  while (!mbar_try_cluster(bar, parity)) asm volatile("nanosleep.u32 %0;" ::"r"((uint32_t)G2V_WAIT_SLEEP_NS));
#else
  mbar_wait_cluster(bar, parity);
#endif
}
// make generic-proxy shared-memory writes visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int CG>
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  if constexpr (CG == 1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(tm), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(tm), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
  }
}
// one lane of a converged warp; everything around it stays warp-uniform, so descriptors and
// addresses live in uniform registers (UTCHMMA / UTMALDG take uniform operands)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
// pull a tile into L2 ahead of time (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* tm, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tm), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Arrive on the barrier at the same offset in CTA `cta` of the cluster.
// The relaxed form is a bare remote arrive; use it when what the barrier orders lives in tensor memory or
// was written by the async proxy (complete before the arrive: tcgen05.wait::ld / wait::st / complete_tx).
// The release form costs MEMBAR.ALL.GPU (~2 us under memory load) and is only needed to publish
// generic-proxy shared-memory writes to the other CTA.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t local_bar, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_bar), "r"(cta));
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster_release(uint32_t local_bar, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_bar), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// arrive on `bar` once every tcgen05 op issued so far by this thread has retired; CG == 2 signals
// the barrier at that offset in BOTH CTAs of the pair
template <int CG>
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  if constexpr (CG == 1) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  } else {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  if constexpr (CG == 1) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor, K-major.  layout: 2 = SWIZZLE_128B, 6 = SWIZZLE_32B.
// sbo = byte distance between consecutive 8-row groups.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);   // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                      // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(sbo_bytes >> 4) << 32;       // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                      // descriptor version 1 (sm_100)
  d |= (uint64_t)layout << 61;
  return d;
}
// instruction descriptor: fp16 x fp16 -> fp32, both K-major, M = m (128, or 256 for a CTA pair), N = n
__device__ __forceinline__ uint32_t umma_idesc(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}


// ------------------------------------------------------------------------------------------
// host side: tensor maps, one-time kernel attributes
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// Tensor maps are pure functions of (pointer, extents, box, swizzle, element type); encoding one costs a driver
// call, and a search needs six or seven.  A small per-thread cache keeps the maps of the buffers a training loop
// or a tokeniser reuses every step (codebook aux buffer, workspace, input batch), so the steady state of a
// small-batch step makes no driver calls for them.
struct MapKey {
  const void* gptr;
  uint64_t rows, cols;
  uint32_t box_cols, box_rows;
  int sw, f32;
  bool operator==(const MapKey& o) const {
    return gptr == o.gptr && rows == o.rows && cols == o.cols && box_cols == o.box_cols && box_rows == o.box_rows &&
           sw == o.sw && f32 == o.f32;
  }
};
constexpr int kMapCache = 32;
struct MapCache {
  MapKey key[kMapCache];
  alignas(64) CUtensorMap map[kMapCache];
  int used = 0, next = 0;
};

int make_map(CUtensorMap* tm, const void* gptr, uint64_t rows, uint64_t cols, uint32_t box_cols, uint32_t box_rows,
             CUtensorMapSwizzle sw, bool f32 = false) {
  static thread_local MapCache cache;
  const MapKey k{gptr, rows, cols, box_cols, box_rows, (int)sw, f32 ? 1 : 0};
  for (int i = 0; i < cache.used; ++i)
    if (cache.key[i] == k) {
      *tm = cache.map[i];
      return G2V_OK;
    }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error_detail("cuTensorMapEncodeTiled entry point not available");
    return G2V_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {cols * (f32 ? 4u : 2u)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                  const_cast<void*>(gptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error_detail("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu box=%ux%u)", (int)r,
                     (unsigned long long)rows, (unsigned long long)cols, box_cols, box_rows);
    return G2V_ERR_CUDA;
  }
  const int slot = cache.used < kMapCache ? cache.used++ : (cache.next++ % kMapCache);
  cache.key[slot] = k;
  cache.map[slot] = *tm;
  return G2V_OK;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) only when a launch needs MORE than what the kernel already has on
// that device.  The attribute belongs to the function (all host threads share it) and a call SETS it, it does not
// raise it: the table is therefore process-wide and keeps the maximum -- a per-thread table would let a small launch
// from one thread (e.g. the autograd engine's) lower the limit under a large launch from another.
inline int set_dyn_smem(const void* func, size_t bytes) {
  struct Ent { const void* f; int dev; size_t bytes; };
  static Ent ent[32];
  static int n = 0;
  static std::mutex mu;
  int dev = 0;
  G2V_CUDA_CHECK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  for (int i = 0; i < n; ++i)
    if (ent[i].f == func && ent[i].dev == dev) {
      if (ent[i].bytes >= bytes) return G2V_OK;
      G2V_CUDA_CHECK(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
      ent[i].bytes = bytes;
      return G2V_OK;
    }
  G2V_CUDA_CHECK(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  if (n < 32) ent[n++] = Ent{func, dev, bytes};
  return G2V_OK;
}

}  // namespace
}  // namespace g2v
