// TISE_CONV_UMMA: Conv2d(bias=False) + BatchNorm(eval) + ReLU as an implicit GEMM on the 5th-gen
// tensor cores (tcgen05.mma, accumulators in TMEM), operands staged in shared memory by TMA.
//
//   D[m, co] = sum_{r,s,ci} X[n, p*stride - pad_h + r, q*stride - pad_w + s, ci] * W[co, r, s, ci]
//   m = (n * Ho + p) * Wo + q                     (NHWC activations, so D is the NHWC output)
//
// * A operand (activations, fp16): TMA *im2col* loads.  One load brings the 128 consecutive
//   output pixels of an M tile x 64 input channels for one filter tap (r, s) into a 128B-swizzled
//   K-major tile; the TMA unit walks the pixels across image rows / images and zero-fills
//   padding, so the im2col matrix never exists in memory.
// * B operand (weights, fp16 hi + lo planes, K-major [cout][taps * cin_pad]): tiled 2D TMA loads.
// * MMA: one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M = 128, N = cout tile <= 256,
//   K = 16) -- for every K step once against W_hi and once against W_lo into the same fp32 TMEM
//   accumulator ("2 passes": fp16 activations x ~22-bit weights; FID stays within 1e-4 of fp32).
// * Epilogue: 4 warps read the accumulator with tcgen05.ld (one output pixel per thread), apply
//   scale/shift (folded BN) from shared memory, ReLU, convert to fp16, stage 128 x 16|32-channel boxes in
//   swizzled shared memory and hand them to TMA stores (cp.async.bulk.tensor...global.shared::cta) that
//   write channel slices of the NHWC destinations -- coalesced, clipped at the tensor end, and torch.cat of
//   the Inception branches costs nothing.  Output channels may be split into segments with different
//   destinations (sibling 1x1 convolutions of a block run as ONE GEMM over the shared input).
// * Persistent: grid = #SMs, each CTA strides over (m tile, n tile) pairs; TMEM holds two
//   accumulators (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of tile i+1;
//   shared memory holds a 3-6 stage TMA<->MMA ring.
//
// Warp roles (192 threads): warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2-5 epilogue.
#include <cudaTypedefs.h>
#include <stdio.h>

#include "common.cuh"

namespace tise {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                         // fp16 elements = one 128-byte swizzle row
constexpr int kABytes = kBlockM * kBlockK * 2;      // 16 KiB
constexpr int kMaxStages = 8;
constexpr int kAccumCols = 256;                     // TMEM columns per accumulator stage
constexpr int kNumThreads = 192;

constexpr int kEpiBufBytes = 128 * 64;              // one staged store box: 128 rows x 64 bytes
constexpr int kEpiBufs = 3;                         // boxes in flight (TMA store read latency ~1 us)
constexpr int kMaxNt = 256;

struct UmmaParams {
  const float* scale;
  const float* shift;
  long long M;
  int out_f32;
  int n_segs;
  int seg_end[kMaxSegs];   // exclusive end column of each segment within the layer's cout
  int seg_relu[kMaxSegs];
  int box_w;               // channels per store box: 32 or 16 (fp16), 16 (fp32) -> 64- or 32-byte rows
  int ho_wo, wo, stride, pad_h, pad_w;
  int R, S, n_chunks, last_ksteps, cin_pad;
  int nt, n_ntiles, num_tiles;
  int passes, stages, stage_bytes;
};

// ----------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// Spins on try_wait (which itself suspends the thread for a hardware-defined interval).  A pipeline
// bug would otherwise hang the GPU; after ~2 s the kernel traps instead, which surfaces as a launch error.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("conv_umma_kernel: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", (int)blockIdx.x,
             (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t dst, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col(const CUtensorMap* map, uint32_t dst, uint32_t bar, int c, int w, int h,
                                                int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2], {%7, %8};" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}
// One lane of a converged warp (elect.sync): true in exactly one lane.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void epi_bar(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(fminf(a, 65504.f), fminf(b, 65504.f));
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start address >> 4 in [0,14), LBO >> 4 in [16,30) (= 1 for swizzled K-major), SBO >> 4 in [32,46)
// (= 1024 B between 8-row groups), version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = f32 (1 << 4), A = B = f16 (0), both K-major,
// N >> 3 in [17,23), M >> 4 in [24,29).
__device__ __forceinline__ uint32_t make_instr_desc(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
}

// ----------------------------------------------------------------------------- kernel
__global__ void __launch_bounds__(kNumThreads, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_bhi,
                 const __grid_constant__ CUtensorMap tm_blo, const __grid_constant__ CUtensorMap tm_o0,
                 const __grid_constant__ CUtensorMap tm_o1, const __grid_constant__ CUtensorMap tm_o2,
                 const __grid_constant__ CUtensorMap tm_o3, const UmmaParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
  // layout: [stages x (A | B_hi | B_lo)] [3 epilogue store boxes] [scale nt | shift nt] [barriers] [tmem slot]
  const uint32_t epi_base = smem_base + (uint32_t)p.stages * (uint32_t)p.stage_bytes;
  const uint32_t ss_base = epi_base + (uint32_t)kEpiBufs * kEpiBufBytes;
  const uint32_t bar_base = ss_base + 2u * kMaxNt * 4u;
  // barriers: full[stages], empty[stages], tmem_full[2], tmem_empty[2]; then the TMEM base address word
  auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(kMaxStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (uint32_t)(2 * kMaxStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(2 * kMaxStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int kb_per_tile = p.R * p.S * p.n_chunks;
  const uint32_t b_bytes = (uint32_t)p.nt * 128u;
  const uint32_t tx_bytes = (uint32_t)kABytes + (uint32_t)p.passes * b_bytes;

  // Producer and MMA warps run their loops warp-uniformly (all 32 lanes wait on the mbarriers and compute the
  // descriptors, which therefore live in uniform registers); only the TMA / tcgen05 instructions themselves
  // are predicated on one elected lane.  A single-lane `if (lane == 0)` loop makes ptxas wrap every
  // uniform-datapath instruction in an election loop (~700 cycles per K block, measured).
  if (warp == 0) {
    // ===================================================================== TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.n_ntiles, n_tile = tile - m_tile * p.n_ntiles;
      const long long m0 = (long long)m_tile * kBlockM;
      const int img = (int)(m0 / p.ho_wo);
      const int rem = (int)(m0 - (long long)img * p.ho_wo);
      const int oy = rem / p.wo, ox = rem - oy * p.wo;
      const int wc = ox * p.stride - p.pad_w, hc = oy * p.stride - p.pad_h;  // base pixel (tap 0,0)
      for (int r = 0; r < p.R; ++r)
        for (int s = 0; s < p.S; ++s)
          for (int ch = 0; ch < p.n_chunks; ++ch) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t sa = smem_base + (uint32_t)stage * (uint32_t)p.stage_bytes;
            const int kcol = (r * p.S + s) * p.cin_pad + ch * kBlockK;
            if (elect_one()) {
              mbar_expect_tx(full_bar(stage), tx_bytes);
              tma_load_im2col(&tm_a, sa, full_bar(stage), ch * kBlockK, wc, hc, img, (uint16_t)s, (uint16_t)r);
              tma_load_2d(&tm_bhi, sa + kABytes, full_bar(stage), kcol, n_tile * p.nt);
              if (p.passes == 2) tma_load_2d(&tm_blo, sa + kABytes + b_bytes, full_bar(stage), kcol, n_tile * p.nt);
            }
            __syncwarp();
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    const uint32_t idesc = make_instr_desc(p.nt);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(tempty_bar(as), aphase ^ 1u);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * kAccumCols);
      int chunk = 0;
      for (int kb = 0; kb < kb_per_tile; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = smem_base + (uint32_t)stage * (uint32_t)p.stage_bytes;
        const uint64_t adesc = make_smem_desc(sa);
        const uint64_t bhi = make_smem_desc(sa + kABytes);
        const uint64_t blo = make_smem_desc(sa + kABytes + b_bytes);
        const int ksteps = (chunk == p.n_chunks - 1) ? p.last_ksteps : (kBlockK / 16);
        if (elect_one()) {
          for (int k = 0; k < ksteps; ++k) {
            // advance 16 fp16 = 32 bytes along K inside the swizzle row: +2 in the (>>4) address field
            umma_f16(d_tmem, adesc + (uint64_t)(2 * k), bhi + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
            if (p.passes == 2) umma_f16(d_tmem, adesc + (uint64_t)(2 * k), blo + (uint64_t)(2 * k), idesc, 1u);
          }
          umma_commit(empty_bar(stage));  // frees the smem stage once these MMAs have read it
          if (kb == kb_per_tile - 1) umma_commit(tfull_bar(as));  // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++chunk == p.n_chunks) chunk = 0;
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (++as == 2) {
        as = 0;
        aphase ^= 1u;
      }
    }
  } else {
    // ===================================================================== epilogue (warps 2..5)
    const int quarter = warp & 3;  // TMEM lanes [32 * quarter, 32 * quarter + 32) are accessible to this warp
    const int row = quarter * 32 + lane;
    const int etid = threadIdx.x - 64;  // 0..127
    float* s_scale = reinterpret_cast<float*>(smem_raw + (ss_base - smem_u32(smem_raw)));
    float* s_shift = s_scale + kMaxNt;
    const CUtensorMap* omaps[kMaxSegs] = {&tm_o0, &tm_o1, &tm_o2, &tm_o3};
    const int w = p.box_w;
    const int row_bytes = w * (p.out_f32 ? 4 : 2);  // 64 or 32
    // Swizzle<B,4,3> phase of this row inside the staged box: 64-byte rows XOR chunk bits with (row >> 1) & 3,
    // 32-byte rows with (row >> 2) & 1 -- the pattern TMA expects for SWIZZLE_64B / SWIZZLE_32B boxes.
    const uint32_t swz = row_bytes == 64 ? (uint32_t)((row >> 1) & 3) : (uint32_t)((row >> 2) & 1);
    int as = 0;
    uint32_t aphase = 0;
    uint32_t box_it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.n_ntiles, n_tile = tile - m_tile * p.n_ntiles;
      const int m0 = m_tile * kBlockM;
      const int co0 = n_tile * p.nt;
      // every epilogue thread is past its last read of s_scale/s_shift of the previous tile (barrier 2 of its last box)
      for (int c = etid; c < p.nt; c += 128) {
        s_scale[c] = __ldg(p.scale + co0 + c);
        s_shift[c] = __ldg(p.shift + co0 + c);
      }
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * kAccumCols);
      int seg = 0;
      for (int c0 = 0; c0 < p.nt; c0 += w, ++box_it) {
        const int col = co0 + c0;
        while (col >= p.seg_end[seg]) ++seg;
        const int seg_col = col - (seg ? p.seg_end[seg - 1] : 0);
        const bool relu = p.seg_relu[seg] != 0;
        const uint32_t buf = epi_base + (box_it % (uint32_t)kEpiBufs) * kEpiBufBytes;
        if (etid == 0) bulk_wait_read<kEpiBufs - 1>();  // the store that used this buffer kEpiBufs boxes ago has read it
        epi_bar(1);                          // ... and s_scale / s_shift of this tile are visible
        float f[32];
        if (w == 32) {
          uint32_t v[32];
          tmem_ld32(taddr + (uint32_t)c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        } else {
          uint32_t v[16];
          tmem_ld16(taddr + (uint32_t)c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
        }
        if (c0 + w >= p.nt) {  // accumulator fully read: hand it back to the MMA warp before the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(as));
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (j < w) {
            float y = fmaf(f[j], s_scale[c0 + j], s_shift[c0 + j]);
            f[j] = relu ? fmaxf(y, 0.f) : y;
          }
        }
        const uint32_t rbase = buf + (uint32_t)row * (uint32_t)row_bytes;
        if (p.out_f32) {  // w == 16: 64-byte rows, 4 chunks of 4 floats
#pragma unroll
          for (int j = 0; j < 4; ++j)
            st_shared_v4(rbase + (((uint32_t)j ^ swz) << 4), __float_as_uint(f[4 * j]), __float_as_uint(f[4 * j + 1]),
                         __float_as_uint(f[4 * j + 2]), __float_as_uint(f[4 * j + 3]));
        } else if (w == 32) {  // 64-byte rows, 4 chunks of 8 halves
#pragma unroll
          for (int j = 0; j < 4; ++j)
            st_shared_v4(rbase + (((uint32_t)j ^ swz) << 4), pack_h2(f[8 * j], f[8 * j + 1]),
                         pack_h2(f[8 * j + 2], f[8 * j + 3]), pack_h2(f[8 * j + 4], f[8 * j + 5]),
                         pack_h2(f[8 * j + 6], f[8 * j + 7]));
        } else {  // w == 16 fp16: 32-byte rows, 2 chunks
#pragma unroll
          for (int j = 0; j < 2; ++j)
            st_shared_v4(rbase + (((uint32_t)j ^ swz) << 4), pack_h2(f[8 * j], f[8 * j + 1]),
                         pack_h2(f[8 * j + 2], f[8 * j + 3]), pack_h2(f[8 * j + 4], f[8 * j + 5]),
                         pack_h2(f[8 * j + 6], f[8 * j + 7]));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the TMA
        epi_bar(2);
        if (etid == 0) {
          tma_store_2d(omaps[seg], buf, seg_col, m0);  // rows beyond M are clipped by the tensor map
          bulk_commit();
        }
      }
      if (++as == 2) {
        as = 0;
        aphase ^= 1u;
      }
    }
    if (etid == 0) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ----------------------------------------------------------------------------- host side
static PFN_cuTensorMapEncodeTiled_v12000 g_encode_tiled = nullptr;
static PFN_cuTensorMapEncodeIm2col_v12000 g_encode_im2col = nullptr;

static int load_driver_entry_points(tise_ctx* ctx) {
  if (g_encode_tiled && g_encode_im2col) return TISE_OK;
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  TISE_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn || q != cudaDriverEntryPointSuccess) return fail(ctx, TISE_ERR_CUDA, "cuTensorMapEncodeTiled not available");
  g_encode_tiled = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  fn = nullptr;
  TISE_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q));
  if (!fn || q != cudaDriverEntryPointSuccess) return fail(ctx, TISE_ERR_CUDA, "cuTensorMapEncodeIm2col not available");
  g_encode_im2col = (PFN_cuTensorMapEncodeIm2col_v12000)fn;
  return TISE_OK;
}

// Largest divisor of cout that is a multiple of 16 and at most 256.
static int n_tile_of(int cout) {
  for (int nt = 256; nt >= 16; nt -= 16)
    if (cout % nt == 0) return nt;
  return 0;
}

int build_umma_plan(tise_ctx* ctx, const ConvLayer& L, const ConvArgs& a, UmmaPlan* plan) {
  int rc = load_driver_entry_points(ctx);
  if (rc) return rc;
  const int nt = n_tile_of(L.cout);
  if (L.cout % 16 || nt == 0) return fail(ctx, TISE_ERR_INVALID, "umma conv: unsupported cout %d", L.cout);
  if (L.cin % 16) return fail(ctx, TISE_ERR_INVALID, "umma conv: cin %d is not a multiple of 16", L.cin);
  if (a.in_stride % 8 || ((uintptr_t)a.in & 15))
    return fail(ctx, TISE_ERR_INVALID, "umma conv: input view is not 16-byte aligned");
  if (a.n_segs < 1 || a.n_segs > kMaxSegs) return fail(ctx, TISE_ERR_INVALID, "umma conv: %d segments", a.n_segs);
  plan->nt = nt;

  // ---- A: im2col map over the NHWC input view, dims (C, W, H, N)
  {
    cuuint64_t dims[4] = {(cuuint64_t)L.cin, (cuuint64_t)a.w, (cuuint64_t)a.h, (cuuint64_t)a.n};
    cuuint64_t strides[3] = {(cuuint64_t)a.in_stride * 2, (cuuint64_t)a.in_stride * 2 * a.w,
                             (cuuint64_t)a.in_stride * 2 * a.w * a.h};
    int lower[2] = {-L.pad_w, -L.pad_h};
    int upper[2] = {L.pad_w - (L.S - 1), L.pad_h - (L.R - 1)};
    cuuint32_t estr[4] = {1, (cuuint32_t)L.stride, (cuuint32_t)L.stride, 1};
    CUresult r = g_encode_im2col(&plan->tm_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void*)a.in, dims, strides, lower,
                                 upper, (cuuint32_t)kBlockK, (cuuint32_t)kBlockM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
      return fail(ctx, TISE_ERR_CUDA, "cuTensorMapEncodeIm2col failed (%d) for %dx%dx%dx%d k%dx%d s%d", (int)r, a.n, a.h,
                  a.w, L.cin, L.R, L.S, L.stride);
    // Driver <= 13.1 sets a flag that breaks im2col maps over tensors smaller than 128 KiB; CUTLASS
    // (cute/atom/copy_traits_sm90_im2col.hpp) clears bit 21 of the second 64-bit word in that case.
    int drv = 0;
    cudaDriverGetVersion(&drv);
    const size_t bytes = (size_t)a.n * a.h * a.w * a.in_stride * 2;
    if (drv <= 13010 && bytes < 131072) reinterpret_cast<uint64_t*>(&plan->tm_a)[1] &= ~(1ull << 21);
  }
  // ---- B: weights [cout][ktot], box 64 x nt
  for (int hl = 0; hl < 2; ++hl) {
    cuuint64_t dims[2] = {(cuuint64_t)L.ktot, (cuuint64_t)L.cout};
    cuuint64_t strides[1] = {(cuuint64_t)L.ktot * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)nt};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(hl ? &plan->tm_blo : &plan->tm_bhi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                                (void*)(hl ? L.d_wlo : L.d_whi), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, TISE_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for weights", (int)r);
  }
  // ---- outputs: one 2D store map per segment, dims (channels of the segment, all output pixels)
  int box_w = a.out_f32 ? 16 : 32;
  int csum = 0;
  for (int i = 0; i < a.n_segs; ++i) {
    if (a.segs[i].c % 16) return fail(ctx, TISE_ERR_INVALID, "umma conv: segment of %d channels", a.segs[i].c);
    csum += a.segs[i].c;
    if (!a.out_f32 && (csum % 32 || nt % 32)) box_w = 16;
  }
  if (csum != L.cout) return fail(ctx, TISE_ERR_INVALID, "umma conv: segments cover %d of %d channels", csum, L.cout);
  plan->box_w = box_w;
  const int elt = a.out_f32 ? 4 : 2;
  const long long M = (long long)a.n * a.ho * a.wo;
  for (int i = 0; i < kMaxSegs; ++i) {
    const ConvSeg& sg = a.segs[i < a.n_segs ? i : 0];  // unused slots repeat segment 0 (never dereferenced)
    if (((uintptr_t)sg.out & 15) || (sg.out_stride * elt) % 16)
      return fail(ctx, TISE_ERR_INVALID, "umma conv: output slice is not 16-byte aligned");
    cuuint64_t dims[2] = {(cuuint64_t)sg.c, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)sg.out_stride * elt};
    cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)kBlockM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(&plan->tm_out[i], a.out_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                                2, sg.out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                box_w * elt == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, TISE_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for output %d", (int)r, i);
  }
  plan->valid = true;
  return TISE_OK;
}

int launch_conv_umma(tise_ctx* ctx, const ConvLayer& L, const ConvArgs& a, const UmmaPlan& plan, int passes,
                     cudaStream_t st) {
  if (!plan.valid) return fail(ctx, TISE_ERR_STATE, "umma conv: plan not built");
  UmmaParams p;
  p.scale = L.d_scale;
  p.shift = L.d_shift;
  p.M = (long long)a.n * a.ho * a.wo;
  p.out_f32 = a.out_f32;
  p.n_segs = a.n_segs;
  int end = 0;
  for (int i = 0; i < kMaxSegs; ++i) {
    if (i < a.n_segs) end += a.segs[i].c;
    p.seg_end[i] = i < a.n_segs ? end : 0x7fffffff;
    p.seg_relu[i] = i < a.n_segs ? a.segs[i].relu : 0;
  }
  p.box_w = plan.box_w;
  p.ho_wo = a.ho * a.wo;
  p.wo = a.wo;
  p.stride = L.stride;
  p.pad_h = L.pad_h;
  p.pad_w = L.pad_w;
  p.R = L.R;
  p.S = L.S;
  p.n_chunks = L.cin_pad / kBlockK;
  p.last_ksteps = (L.cin - (p.n_chunks - 1) * kBlockK + 15) / 16;
  p.cin_pad = L.cin_pad;
  p.nt = plan.nt;
  p.n_ntiles = L.cout / p.nt;
  const int m_tiles = cdiv(p.M, kBlockM);
  p.num_tiles = m_tiles * p.n_ntiles;
  p.passes = passes;
  p.stage_bytes = kABytes + passes * p.nt * 128;

  static int configured_smem = 0;
  const int smem_limit = ctx->smem_optin > 0 ? ctx->smem_optin : 232448;
  if (configured_smem != smem_limit) {
    TISE_CUDA(ctx, cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit));
    configured_smem = smem_limit;
  }
  const int reserve = 1024 /*alignment*/ + kEpiBufs * kEpiBufBytes + 2 * kMaxNt * 4 + 8 * (2 * kMaxStages + 4) + 16;
  int stages = (smem_limit - reserve) / p.stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return fail(ctx, TISE_ERR_INVALID, "umma conv: tile does not fit shared memory");
  p.stages = stages;
  const int grid = p.num_tiles < ctx->num_sms ? p.num_tiles : ctx->num_sms;
  // always request the full opt-in size: one CTA per SM, so the 512-column TMEM allocation never contends
  conv_umma_kernel<<<grid, kNumThreads, smem_limit, st>>>(plan.tm_a, plan.tm_bhi, plan.tm_blo, plan.tm_out[0],
                                                         plan.tm_out[1], plan.tm_out[2], plan.tm_out[3], p);
  TISE_LAUNCH_CHECK(ctx);
  return TISE_OK;
}

}  // namespace tise
