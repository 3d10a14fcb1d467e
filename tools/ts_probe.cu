// Bring-up probe: (1) register -> TMEM mapping of tcgen05.st.16x256b.x4, (2) tcgen05.mma with the A operand
// in TMEM (fp16, lane = row, 32-bit column = two K elements) for one CTA and for a CTA pair, at the
// accumulator / operand column offsets the search kernel wants to use.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ts_probe tools/ts_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra.uni WD;\n\tbra.uni WL;\n\tWD:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__device__ __forceinline__ uint32_t umma_idesc(int m, int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }

template <int CG>
__global__ void __launch_bounds__(128, 1) probe(const __half* A, const __half* B, float* D, uint32_t* dump, int N, int a_col, int d_col) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t sB = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* gB = smem_raw + (sB - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    if constexpr (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if constexpr (CG == 1) __syncthreads(); else cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;

  // ---- A rows of this CTA -> TMEM through 16x256b.x4 (64 K elements = 32 columns) ----
  const __half* Ac = A + (size_t)rank * 128 * 64;
  for (int hh = 0; hh < 2; ++hh) {
    const int r0 = 32 * warp + 16 * hh + lane / 4, r1 = r0 + 8, kq = 4 * (lane % 4);
    uint32_t v[16];
    for (int j = 0; j < 4; ++j) {
      const __half* p0 = Ac + (size_t)r0 * 64 + 16 * j + kq;
      const __half* p1 = Ac + (size_t)r1 * 64 + 16 * j + kq;
      v[4 * j + 0] = *reinterpret_cast<const uint32_t*>(p0);
      v[4 * j + 1] = *reinterpret_cast<const uint32_t*>(p0 + 2);
      v[4 * j + 2] = *reinterpret_cast<const uint32_t*>(p1);
      v[4 * j + 3] = *reinterpret_cast<const uint32_t*>(p1 + 2);
    }
    const uint32_t taddr = tmem_base + ((uint32_t)(32 * warp + 16 * hh) << 16) + (uint32_t)a_col;
    asm volatile("tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
                 "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
                 : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  // ---- B rows of this CTA (N / CG codes) -> shared memory, K-major SWIZZLE_128B ----
  const int nb = N / CG;
  for (int t = threadIdx.x; t < nb * 8; t += 128) {
    const int n = t >> 3, ch = t & 7;
    const uint4 val = *reinterpret_cast<const uint4*>(B + ((size_t)rank * nb + n) * 64 + ch * 8);
    *reinterpret_cast<uint4*>(gB + n * 128 + ((ch ^ (n & 7)) << 4)) = val;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if constexpr (CG == 1) __syncthreads(); else cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  if (rank == 0 && threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc(128 * CG, N);
    const uint64_t bd = umma_desc(sB, 1024, 2);
    for (int k = 0; k < 4; ++k) {
      const uint32_t acc = k != 0;
      if constexpr (CG == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_base + d_col),
                     "r"(tmem_base + a_col + 8 * k), "l"(bd + 2u * k), "r"(idesc), "r"(acc)
                     : "memory");
      else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_base + d_col),
                     "r"(tmem_base + a_col + 8 * k), "l"(bd + 2u * k), "r"(idesc), "r"(acc)
                     : "memory");
    }
    if constexpr (CG == 1)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    else
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
  }
  mbar_wait(smem_u32(&bar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int row = 32 * warp + lane;
  const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * warp) << 16);
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(lane_addr + d_col + c0) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) D[((size_t)rank * 128 + row) * N + c0 + j] = __uint_as_float(v[j]);
  }
  for (int c0 = 0; c0 < 32; c0 += 16) {
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(lane_addr + a_col + c0) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) dump[((size_t)rank * 128 + row) * 32 + c0 + j] = v[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if constexpr (CG == 1) __syncthreads(); else cluster_sync_all();
  if (warp == 0) {
    if constexpr (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

template <int CG>
static void run(int N, int a_col, int d_col) {
  const int M = 128 * CG;
  std::vector<__half> A((size_t)M * 64), B((size_t)N * 64);
  std::vector<float> Af(A.size()), Bf(B.size());
  for (size_t i = 0; i < A.size(); ++i) { Af[i] = (float)((int)((i * 7 + (i / 64) * 3) % 17) - 8); A[i] = __float2half(Af[i]); }
  for (size_t i = 0; i < B.size(); ++i) { Bf[i] = (float)((int)((i * 5 + (i / 64) * 11) % 13) - 6); B[i] = __float2half(Bf[i]); }
  __half *dA, *dB; float* dD; uint32_t* dd;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, (size_t)M * N * 4); cudaMalloc(&dd, (size_t)M * 32 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, (size_t)M * N * 4); cudaMemset(dd, 0xff, (size_t)M * 32 * 4);
  const size_t smem = 256 * 128 + 2048;
  cudaFuncSetAttribute(probe<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CG); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, probe<CG>, (const __half*)dA, (const __half*)dB, dD, dd, N, a_col, d_col);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CG=%d N=%d a_col=%d d_col=%d: CUDA error %s\n", CG, N, a_col, d_col, cudaGetErrorString(e)); exit(1); }
  std::vector<float> D((size_t)M * N); std::vector<uint32_t> dump((size_t)M * 32);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(dump.data(), dd, dump.size() * 4, cudaMemcpyDeviceToHost);
  long bad_a = 0, bad_d = 0;
  for (int r = 0; r < M; ++r)
    for (int c = 0; c < 32; ++c) {
      const uint32_t lo = __half_as_ushort(A[(size_t)r * 64 + 2 * c]), hi = __half_as_ushort(A[(size_t)r * 64 + 2 * c + 1]);
      if (dump[(size_t)r * 32 + c] != (lo | (hi << 16))) ++bad_a;
    }
  for (int r = 0; r < M; ++r)
    for (int n = 0; n < N; ++n) {
      float ref = 0.f;
      for (int k = 0; k < 64; ++k) ref += Af[(size_t)r * 64 + k] * Bf[(size_t)n * 64 + k];
      if (D[(size_t)r * N + n] != ref) ++bad_d;
    }
  printf("CG=%d N=%3d a_col=%3d d_col=%3d: A-image mismatches %ld / %d, D mismatches %ld / %d  %s\n", CG, N, a_col, d_col, bad_a, M * 32, bad_d,
         M * N, (bad_a == 0 && bad_d == 0) ? "PASS" : "FAIL");
  if (bad_a) {   // show where thread registers landed for the first quarter
    for (int r = 0; r < 18; ++r) {
      printf("  row %2d:", r);
      for (int c = 0; c < 10; ++c) {
        // decode which (row, k-pair) of A this word holds
        const uint32_t w = dump[(size_t)r * 32 + c];
        int fr = -1, fc = -1;
        for (int rr = 0; rr < M && fr < 0; ++rr)
          for (int cc = 0; cc < 32; ++cc) {
            const uint32_t lo = __half_as_ushort(A[(size_t)rr * 64 + 2 * cc]), hi = __half_as_ushort(A[(size_t)rr * 64 + 2 * cc + 1]);
            if (w == (lo | (hi << 16))) { fr = rr; fc = cc; break; }
          }
        printf(" (%d,%d)", fr, fc);
      }
      printf("\n");
    }
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dd);
}

int main() {
  run<1>(128, 0, 256);
  run<1>(144, 0, 224);
  run<1>(144, 0, 368);
  run<1>(144, 288, 0);
  run<1>(112, 288, 144);
  run<1>(16, 0, 256);
  run<1>(144, 200, 0);
  run<1>(144, 208, 368);
  run<2>(256, 0, 256);
  run<2>(144, 0, 224);
  run<2>(112, 0, 368);
  run<2>(144, 288, 144);
  run<2>(32, 288, 0);
  return 0;
}
