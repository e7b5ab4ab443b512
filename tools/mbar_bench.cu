// Micro-benchmark: cost of an mbarrier producer/consumer ring between two warps of one CTA (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mbar_bench tools/mbar_bench.cu && /tmp/mbar_bench
// Variants: wait flavour (try_wait loop with a clock read, plain try_wait loop, test_wait spin), whole warp vs one
// lane waiting, and 0 / 4 extra warps spinning on a barrier that never completes (like idle epilogue warps).
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ uint32_t test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
template <int MODE>
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity) {
  if (MODE == 0) {  // as in conv_umma.cu: try_wait, then a loop with a clock read
    if (try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!try_wait(bar, parity)) {
      if (clock64() - t0 > 4000000000LL) __trap();
    }
  } else if (MODE == 1) {
    while (!try_wait(bar, parity)) {
    }
  } else {
    while (!test_wait(bar, parity)) {
    }
  }
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}

template <int MODE, bool WHOLE_WARP>
__global__ void ring(int stages, int iters, int spinners, long long* out) {
  __shared__ __align__(8) uint64_t bars[2 * 8 + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[8]), never = smem_u32(&bars[16]);
  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(never, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  if (warp == 0) {  // producer
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < iters; ++i) {
      if (WHOLE_WARP) {
        wait<MODE>(empty0 + 8 * stage, phase ^ 1u);
        if (elect_one()) mbar_arrive(full0 + 8 * stage);
        __syncwarp();
      } else if (lane == 0) {
        wait<MODE>(empty0 + 8 * stage, phase ^ 1u);
        mbar_arrive(full0 + 8 * stage);
      }
      if (++stage == stages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {  // consumer
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < iters; ++i) {
      if (WHOLE_WARP) {
        wait<MODE>(full0 + 8 * stage, phase);
        if (elect_one()) mbar_arrive(empty0 + 8 * stage);
        __syncwarp();
      } else if (lane == 0) {
        wait<MODE>(full0 + 8 * stage, phase);
        mbar_arrive(empty0 + 8 * stage);
      }
      if (++stage == stages) { stage = 0; phase ^= 1u; }
    }
    if (lane == 0) {
      out[0] = clock64() - t0;
      mbar_arrive(never);  // release the spinners
    }
  } else if (warp - 2 < spinners) {
    wait<MODE>(never, 0);
  }
}

template <int MODE, bool WW>
void run(const char* name, int stages, int spinners) {
  long long* d;
  cudaMalloc(&d, 8);
  const int iters = 20000;
  ring<MODE, WW><<<1, 192>>>(stages, iters, spinners, d);
  ring<MODE, WW><<<1, 192>>>(stages, iters, spinners, d);
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%-34s stages %d spinners %d : %7.1f cycles / iteration  (%s)\n", name, stages, spinners, (double)h / iters,
         cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int sp = 0; sp <= 4; sp += 4)
    for (int st = 1; st <= 8; st *= 2) {
      run<0, true>("try_wait+clock loop, whole warp", st, sp);
      run<1, true>("try_wait loop, whole warp", st, sp);
      run<2, true>("test_wait spin, whole warp", st, sp);
      run<0, false>("try_wait+clock loop, one lane", st, sp);
      run<1, false>("try_wait loop, one lane", st, sp);
      run<2, false>("test_wait spin, one lane", st, sp);
    }
  return 0;
}
