// Microbenchmark: per-SM TMA streaming rate for different box shapes / ring depths (bring-up aid).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tools/tma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra.uni WD;\n\tbra.uni WL;\n\tWD:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

// each CTA streams its own band of rows: tiles of box_rows rows, all column boxes of each tile
__global__ void __launch_bounds__(64, 1) probe(const __grid_constant__ CUtensorMap tm, int box_rows, int box_cols, int ncolbox,
                                                int tiles_per_cta, int nslot, int slot_bytes, float* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t full[16], empty[16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 16; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t sb = (smem_u32(smem) + 1023u) & ~1023u;
  const int total = tiles_per_cta * ncolbox;
  if (warp == 0) {
    if (lane == 0) {
      for (int g = 0; g < total; ++g) {
        const int s = g % nslot, round = g / nslot;
        mbar_wait(smem_u32(&empty[s]), (round & 1) ^ 1);
        mbar_expect_tx(smem_u32(&full[s]), slot_bytes);
        const int tile = blockIdx.x * tiles_per_cta + g / ncolbox, cb = g % ncolbox;
        tma_load_2d(sb + s * slot_bytes, &tm, cb * box_cols, tile * box_rows, smem_u32(&full[s]));
      }
    }
  } else {
    float acc = 0.f;
    for (int g = 0; g < total; ++g) {
      const int s = g % nslot, round = g / nslot;
      mbar_wait(smem_u32(&full[s]), round & 1);
      acc += *reinterpret_cast<const float*>(smem + (sb - smem_u32(smem)) + s * slot_bytes + lane * 4);
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&empty[s]));
    }
    if (acc == 123.456f) sink[0] = acc;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fnp;
  const size_t rows = 1 << 20, cols_b = 1600;           // 1M rows of 1600 bytes
  char* buf; cudaMalloc(&buf, rows * cols_b); cudaMemset(buf, 0, rows * cols_b);
  float* sink; cudaMalloc(&sink, 4);
  struct Cfg { const char* name; int esz; int box_cols; int box_rows; CUtensorMapSwizzle sw; int nslot; };
  Cfg cfgs[] = {
      {"f32 16x128 (64B rows)  x6", 4, 16, 128, CU_TENSOR_MAP_SWIZZLE_NONE, 6},
      {"f32 32x128 (128B rows) x3", 4, 32, 128, CU_TENSOR_MAP_SWIZZLE_NONE, 3},
      {"f32 32x128 (128B rows) x6", 4, 32, 128, CU_TENSOR_MAP_SWIZZLE_NONE, 6},
      {"f32 64x64  (256B rows) x3", 4, 64, 64, CU_TENSOR_MAP_SWIZZLE_NONE, 3},
      {"f32 64x64  (256B rows) x6", 4, 64, 64, CU_TENSOR_MAP_SWIZZLE_NONE, 6},
      {"f32 100x32 (400B rows) x4", 4, 100, 32, CU_TENSOR_MAP_SWIZZLE_NONE, 4},
      {"f32 200x16 (800B rows) x4", 4, 200, 16, CU_TENSOR_MAP_SWIZZLE_NONE, 4},
      {"f16 64x128 sw128       x3", 2, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B, 3},
      {"f16 64x128 sw128       x7", 2, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B, 7},
      {"f16 64x256 sw128       x3", 2, 64, 256, CU_TENSOR_MAP_SWIZZLE_128B, 3},
  };
  for (auto& c : cfgs) {
    const int ncols = cols_b / c.esz;
    const int ncolbox = (ncols + c.box_cols - 1) / c.box_cols;
    alignas(64) CUtensorMap tm;
    cuuint64_t gdim[2] = {(cuuint64_t)ncols, rows}; cuuint64_t gstr[1] = {cols_b};
    cuuint32_t box[2] = {(cuuint32_t)c.box_cols, (cuuint32_t)c.box_rows}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm, c.esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, gdim, gstr, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, c.sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("%s: encode failed %d\n", c.name, (int)r); continue; }
    const int slot_bytes = c.box_cols * c.box_rows * c.esz;
    const int tiles_total = rows / c.box_rows, grid = 148, tiles_per_cta = tiles_total / grid;
    const size_t smem = (size_t)slot_bytes * c.nslot + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      probe<<<grid, 64, smem>>>(tm, c.box_rows, c.box_cols, ncolbox, tiles_per_cta, c.nslot, slot_bytes, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = (double)grid * tiles_per_cta * c.box_rows * cols_b;
    printf("%-28s slot %6d B  in flight %7d B/SM : %7.3f ms  %7.1f GB/s  (%5.1f GB/s per SM)  err=%s\n", c.name, slot_bytes,
           slot_bytes * c.nslot, ms, bytes / ms / 1e6, bytes / ms / 1e6 / grid, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
