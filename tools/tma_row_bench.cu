// Micro-benchmark: what does the TMA unit charge for -- rows or bytes?  One warp per CTA (one CTA per SM) streams
// 12 KB weight tiles from an L2-resident tensor into a 4-stage shared-memory ring in three forms:
//   A: 2D tiled load, box 64 halfs x 96 rows, SWIZZLE_128B        (what conv_umma_kernel does for its B operand)
//   B: 2D tiled load, box 256 halfs x 24 rows, SWIZZLE_NONE       (same bytes, rows four times as wide)
//   C: cp.async.bulk, one contiguous 12 KB copy                   (no tensor map)
// and, for the A-operand side, D: box 64 halfs x 128 rows SWIZZLE_128B.  Prints cycles per tile (issue-to-issue, steady state).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma_row_bench tools/tma_row_bench.cu -lcuda && /tmp/tma_row_bench
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma_2d(const CUtensorMap* map, uint32_t dst, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

constexpr int kMaxStages = 32;
template <int MODE>
__global__ void __launch_bounds__(32, 1) bench(const __grid_constant__ CUtensorMap map, const __half* src, int tile_bytes, int rows_total,
                                               int iters, int kStages, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[kMaxStages];
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int rows_per_tile = MODE == 1 ? tile_bytes / 512 : tile_bytes / 128;
  long long t0 = 0;
  const int n_tiles = rows_total / rows_per_tile;
  int tile = (blockIdx.x * 131) % n_tiles;
  uint32_t phase = 0;
  int stage = 0;
  for (int it = 0; it < iters + kStages; ++it) {
    if (it == kStages) t0 = clock64();
    if (it >= kStages) mbar_wait(smem_u32(&bars[stage]), phase ^ 1u);  // the load issued kStages iterations ago has landed
    {  // warp-uniform loop, only the TMA instructions are predicated on one elected lane (as in conv_umma.cu)
      const uint32_t bar = smem_u32(&bars[0]) + 8u * stage;
      tile += 7;
      if (tile >= n_tiles) tile -= n_tiles;
      if (elect_one()) {
        mbar_expect_tx(bar, (uint32_t)tile_bytes);
        if (MODE == 2) bulk_1d(base + stage * tile_bytes, reinterpret_cast<const uint8_t*>(src) + (size_t)tile * tile_bytes, (uint32_t)tile_bytes, bar);
        else tma_2d(&map, base + stage * tile_bytes, bar, 0, tile * rows_per_tile);
      }
    }
    __syncwarp();
    if (++stage == kStages) {
      stage = 0;
      phase ^= 1u;
    }
  }
  for (int s = 0; s < kStages; ++s) {  // drain
    mbar_wait(smem_u32(&bars[stage]), phase ^ 1u);
    if (++stage == kStages) {
      stage = 0;
      phase ^= 1u;
    }
  }
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

int main() {
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  auto encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  const size_t bytes = 8u << 20;
  __half* d = nullptr;
  cudaMalloc(&d, bytes);
  cudaMemset(d, 0, bytes);
  long long* out;
  cudaMalloc(&out, 148 * sizeof(long long));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int tile_rows128 : {16, 48, 96, 128}) {
    const int tile_bytes = tile_rows128 * 128;
    CUtensorMap mapA, mapB;
    {
      cuuint64_t dims[2] = {64, bytes / 128};
      cuuint64_t strides[1] = {128};
      cuuint32_t box[2] = {64, (cuuint32_t)tile_rows128};
      cuuint32_t estr[2] = {1, 1};
      CUresult r = encode(&mapA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r) printf("encode A failed %d\n", (int)r);
    }
    {
      cuuint64_t dims[2] = {256, bytes / 512};
      cuuint64_t strides[1] = {512};
      cuuint32_t box[2] = {256, (cuuint32_t)(tile_rows128 / 4)};
      cuuint32_t estr[2] = {1, 1};
      CUresult r = encode(&mapB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r) printf("encode B failed %d\n", (int)r);
    }
    const int iters = 2000;
    int kStages = (200 * 1024) / tile_bytes; if (kStages > kMaxStages) kStages = kMaxStages;
    const size_t smem = (size_t)kStages * tile_bytes + 2048;
    cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(bench<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int mode = 0; mode < 3; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) bench<0><<<sms, 32, smem>>>(mapA, d, tile_bytes, (int)(bytes / 128), iters, kStages, out);
        if (mode == 1) bench<1><<<sms, 32, smem>>>(mapB, d, tile_bytes, (int)(bytes / 512), iters, kStages, out);
        if (mode == 2) bench<2><<<sms, 32, smem>>>(mapA, d, tile_bytes, (int)(bytes / 128), iters, kStages, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
      }
      long long h[148];
      cudaMemcpy(h, out, sms * sizeof(long long), cudaMemcpyDeviceToHost);
      double s = 0;
      for (int i = 0; i < sms; ++i) s += (double)h[i];
      const char* names[3] = {"2D box 64 x rows, SWIZZLE_128B", "2D box 256 x rows/4, SWIZZLE_NONE", "cp.async.bulk 1D"};
      printf("stages %2d tile %5d B (%3d rows of 128 B)  %-36s %7.1f cycles / tile  = %5.2f cycles per 128 B, %5.1f B/clk/SM\n", kStages, tile_bytes,
             tile_rows128, names[mode], s / sms / iters, s / sms / iters / tile_rows128, tile_bytes / (s / sms / iters));
    }
  }
  return 0;
}
