// Conv2d_1a_3x3 (3 -> 32, 3x3, stride 2, no padding; 299 -> 149) on the tensor cores, fused with the uint8 -> float input
// conversion (ToTensor + affine, tise_set_input_lut) and with BN + ReLU.
//
// K = 27 is too thin for TMA im2col (a 3-channel uint8 image is not a tensor-core operand), so the A tile is built
// by CUDA cores: thread = output pixel, 27 byte loads -> per-channel lookup table -> one 128-byte row
//     [ x_hi (27 fp16) 0 0 0 0 0 | x_lo (27 fp16) 0 0 0 0 0 ]          x = x_hi + x_lo  (~22 mantissa bits)
// written into shared memory in the 128B-swizzled K-major layout tcgen05.mma expects.  The weights sit resident in
// shared memory as 64 rows [ W_hi (32 channels) ; W_lo (32 channels) ], each row [ w (27) 0.. | w (27) 0.. ], so ONE
// MMA sequence of M = 128, N = 64, K = 64 computes (x_hi + x_lo) . W_hi | (x_hi + x_lo) . W_lo in TMEM and the epilogue
// adds the halves: fp32-grade products like the other layers (fp16 activations x (hi + lo) weights) but with the
// network input itself carried at ~22 bits.  864 FMAs per output pixel on the CUDA cores (the previous stem, 47 % of
// the fp32 pipe) become ~130 gather/pack instructions.
//
// Warp roles (672 threads): warps 0-11 three gather groups (a CTA's tiles go round-robin to them, so the chain of one
// tile overlaps those of the next two), warp 12 TMEM allocator + MMA issuer, warps 13-20 two epilogue groups (one per
// TMEM accumulator stage, as in conv_umma.cu).
#include <stdlib.h>

#include "common.cuh"
#include "umma_ptx.cuh"

namespace tise {

// The gather groups set the pace of this kernel (a ~170-instruction dependent chain per 128-pixel tile and warp: 27 table
// look-ups, packing, 8 swizzled stores), so there are THREE of them (round 1: two) feeding two epilogue groups; 672 threads
// leave enough registers (80 used).  Measured per 1024 images: 2 groups 0.84 ms, 3 groups 0.76 ms, 4 groups (800 threads, 72
// registers, spills) 0.85 ms.
constexpr int kSuGather = 3;                                // gather groups of 4 warps
constexpr int kSuMmaWarp = 4 * kSuGather;                   // warp index of the TMEM allocator + MMA issuer
constexpr int kSuThreads = 32 * (kSuMmaWarp + 1) + 256;     // + two epilogue groups of 4 warps
constexpr int kSuStages = 2 * kSuGather;     // A stages: gather group g fills stages g and g + kSuGather
constexpr int kSuATile = 128 * 128;          // 128 pixels x 128 bytes
constexpr int kSuBTile = 64 * 128;           // 64 weight rows x 128 bytes
constexpr int kSuEpiBuf = 128 * 64;          // staged output box: 128 pixels x 32 fp16 channels
constexpr int kSuOut = 149;
constexpr int kSuLutRep = 8;                 // copies of the lookup table, one per (lane & 7): 27 lookups per pixel with
constexpr int kSuLutBytes = 768 * 4 * kSuLutRep;  // arbitrary byte values were 3-4-way bank conflicts on a single copy
constexpr int kSuSmem = 1024 + kSuStages * kSuATile + kSuBTile + 4 * kSuEpiBuf + kSuLutBytes + 2 * 32 * 4 + 8 * (2 * kSuStages + 4) + 16;

struct StemUmmaParams {
  const uint8_t* u8;      // [n][299][299][3]
  const uint8_t* u8_end;  // one past its last byte
  const float* lut;       // [3][256]
  const __half* w16;      // [64][64] rows as described above (unswizzled)
  const float* scale;     // [32] folded BN scale x 2^-e of the weight pre-scaling
  const float* shift;     // [32]
  __half* out;            // [n][149][149][32]
  long long total;        // n * 149 * 149 output pixels
  int num_tiles;
};

__global__ void __launch_bounds__(kSuThreads, 1) stem_umma_kernel(const StemUmmaParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + kSuStages * kSuATile;
  const uint32_t epi_base = b_base + kSuBTile;
  const uint32_t lut_base = epi_base + 4u * kSuEpiBuf;
  const uint32_t ss_base = lut_base + (uint32_t)kSuLutBytes;
  const uint32_t bar_base = ss_base + 2u * 32u * 4u;
  auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(kSuStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (uint32_t)(2 * kSuStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(2 * kSuStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kSuStages + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  uint32_t* s_lut = reinterpret_cast<uint32_t*>(smem_raw + (lut_base - smem_u32(smem_raw)));  // hi | lo << 16 per (channel, value)
  float* s_scale = reinterpret_cast<float*>(smem_raw + (ss_base - smem_u32(smem_raw)));
  float* s_shift = s_scale + 32;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kSuStages; ++s) {
      mbar_init(full_bar(s), 4);   // one arrive per warp of the gather group that owns the stage
      mbar_init(empty_bar(s), 1);  // tcgen05.commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // lookup table with the hi / lo split, scale / shift, and the resident weight tile (128B-swizzled rows)
  for (int i = threadIdx.x; i < 768 * kSuLutRep; i += blockDim.x) {  // entry e of copy r at [e * kSuLutRep + r]
    const float v = __ldg(p.lut + i / kSuLutRep);
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    s_lut[i] = (uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(lo) << 16);
  }
  if (threadIdx.x < 32) {
    s_scale[threadIdx.x] = __ldg(p.scale + threadIdx.x);
    s_shift[threadIdx.x] = __ldg(p.shift + threadIdx.x);
  }
  for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) {  // (row, 16-byte chunk)
    const int r = i >> 3, c = i & 7;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.w16) + i);
    st_shared_v4(b_base + (uint32_t)r * 128u + ((uint32_t)(c ^ (r & 7)) << 4), v.x, v.y, v.z, v.w);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
  if (warp == kSuMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < kSuMmaWarp) {
    // ===================================================================== gather: uint8 pixels -> swizzled fp16 A rows
    const int grp = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    uint32_t phase = 0;  // parity of this group's stages (each is reused every 2 of its tiles)
    int it = 0;
    // The 9 aligned words of a pixel's 3 x 9-byte window are requested ONE TILE AHEAD: the global-load latency (~1000
    // cycles with every SM streaming) was the longest link of this group's per-tile chain (load -> 27 table look-ups ->
    // pack -> 8 stores); (32 table copies instead of 8 -- no bank conflicts at all -- made the kernel 6 % slower.)
    uint32_t wn[9], offn[3];
    auto fetch = [&](int tile_f) {
      const long long m = (long long)tile_f * kBlockM + row;
#pragma unroll
      for (int k = 0; k < 9; ++k) wn[k] = 0u;
#pragma unroll
      for (int r = 0; r < 3; ++r) offn[r] = 0u;
      if (tile_f < p.num_tiles && m < p.total) {
        const unsigned mu = (unsigned)m;  // (n * 149 * 149 < 2^31: checked on the host)
        const int img = (int)(mu / (unsigned)(kSuOut * kSuOut));
        const int rem = (int)(mu - (unsigned)img * (unsigned)(kSuOut * kSuOut));
        const int oy = rem / kSuOut, ox = rem - oy * kSuOut;
        const uint8_t* base = p.u8 + ((size_t)img * kOutSize + 2 * oy) * (kOutSize * 3) + 6 * ox;
        // 9 contiguous bytes per input row at any alignment (rows are 897 bytes): three aligned 32-bit loads + byte permutes
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const uint8_t* rp = base + (size_t)r * (kOutSize * 3);
          const uint32_t off = (uint32_t)((uintptr_t)rp & 3u);
          const uint32_t* wp = reinterpret_cast<const uint32_t*>(rp - off);
          offn[r] = off;
          wn[3 * r] = __ldg(wp);
          wn[3 * r + 1] = __ldg(wp + 1);
          if (reinterpret_cast<const uint8_t*>(wp + 3) <= p.u8_end) {
            wn[3 * r + 2] = __ldg(wp + 2);
          } else {  // the very last pixels of the batch: never read past the end of the caller's buffer
            const uint8_t* bp = reinterpret_cast<const uint8_t*>(wp + 2);
            uint32_t w2 = 0;
            for (int b = 0; b < 4; ++b)
              if (bp + b < p.u8_end) w2 |= (uint32_t)__ldg(bp + b) << (8 * b);
            wn[3 * r + 2] = w2;
          }
        }
      }
    };
    fetch(blockIdx.x + grp * gridDim.x);
    for (int tile = blockIdx.x + grp * gridDim.x; tile < p.num_tiles; tile += kSuGather * gridDim.x, ++it) {
      const int stage = grp + kSuGather * (it & 1);
      if (it > 0 && (it & 1) == 0) phase ^= 1u;
      const long long m = (long long)tile * kBlockM + row;
      uint32_t wc[9], offc[3];
#pragma unroll
      for (int k = 0; k < 9; ++k) wc[k] = wn[k];
#pragma unroll
      for (int r = 0; r < 3; ++r) offc[r] = offn[r];
      fetch(tile + kSuGather * gridDim.x);  // next tile of this group: in flight during the look-ups below
      uint32_t v[27];
      if (m < p.total) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const uint32_t sel = 0x3210u + 0x1111u * offc[r];  // bytes off .. off + 3 of a register pair
          const uint32_t b0 = __byte_perm(wc[3 * r], wc[3 * r + 1], sel), b1 = __byte_perm(wc[3 * r + 1], wc[3 * r + 2], sel),
                         b2 = __byte_perm(wc[3 * r + 2], 0u, sel);
#pragma unroll
          for (int b = 0; b < 9; ++b) {
            const uint32_t word = b < 4 ? b0 : (b < 8 ? b1 : b2);
            const uint32_t u = (word >> (8 * (b & 3))) & 0xffu;
            v[r * 9 + b] = s_lut[((b % 3) * 256 + u) * kSuLutRep + (lane & (kSuLutRep - 1))];  // k = r * 9 + b, b = s * 3 + channel
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < 27; ++k) v[k] = 0u;
      }
      mbar_wait(empty_bar(stage), phase ^ 1u);
      const uint32_t dst = a_base + (uint32_t)stage * kSuATile + (uint32_t)row * 128u;
      const uint32_t sw = (uint32_t)(row & 7);
#pragma unroll
      for (int c = 0; c < 4; ++c) {  // chunk c: x_hi[8c .. 8c+7]; chunk 4 + c: x_lo[8c .. 8c+7]
        uint32_t h[4], l[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int k0 = 8 * c + 2 * q, k1 = k0 + 1;
          const uint32_t a = k0 < 27 ? v[k0] : 0u, b = k1 < 27 ? v[k1] : 0u;
          h[q] = __byte_perm(a, b, 0x5410);
          l[q] = __byte_perm(a, b, 0x7632);
        }
        st_shared_v4(dst + (((uint32_t)c ^ sw) << 4), h[0], h[1], h[2], h[3]);
        st_shared_v4(dst + (((uint32_t)(4 + c) ^ sw) << 4), l[0], l[1], l[2], l[3]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar(stage));
    }
  } else if (warp == kSuMmaWarp) {
    // ===================================================================== MMA issuer
    const uint64_t dhi = smem_desc_hi(2 /*SWIZZLE_128B*/, 1024);
    const uint32_t hi = (uint32_t)(dhi >> 32);
    const uint32_t a_lo0 = (uint32_t)make_smem_desc(a_base, dhi), b_lo = (uint32_t)make_smem_desc(b_base, dhi);
    const uint32_t idesc = make_instr_desc(64);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int grp = it % kSuGather, git = it / kSuGather;  // this tile was built by gather group grp as its git-th tile
      const int stage = grp + kSuGather * (git & 1);
      const uint32_t fphase = (uint32_t)(git >> 1) & 1u;   // the stage's git / 2-th use
      const int as = it & 1;                                // accumulator stage = epilogue group
      const uint32_t aph = (uint32_t)(it >> 1) & 1u;        // the accumulator's it / 2-th use
      mbar_wait(tempty_bar(as), aph ^ 1u);
      mbar_wait(full_bar(stage), fphase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d = tmem_base + (uint32_t)(as * 64);
        const uint32_t a_lo = a_lo0 + (uint32_t)stage * (kSuATile >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_lh(d, a_lo + 2u * k, b_lo + 2u * k, hi, idesc, (uint32_t)k);
        umma_commit(empty_bar(stage));
        umma_commit(tfull_bar(as));
      }
    }
  } else {
    // ===================================================================== epilogue (two groups, one per accumulator stage)
    const int grp = (warp - (kSuMmaWarp + 1)) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int etid = threadIdx.x - (kSuMmaWarp + 1) * 32 - grp * 128;
    const uint32_t my_epi = epi_base + (uint32_t)grp * 2u * kSuEpiBuf;
    const uint32_t wr_off = (uint32_t)row * 64u;
    const uint32_t wr_swz = (uint32_t)((row >> 1) & 3);
    const int st_chunk = etid & 3, st_row0 = etid >> 2;
    const int as = grp;
    uint32_t aphase = 0;
    uint32_t box_it = 0;
    for (int tile = blockIdx.x + grp * gridDim.x; tile < p.num_tiles; tile += 2 * gridDim.x, ++box_it) {
      const long long m0 = (long long)tile * kBlockM;
      mbar_wait(tfull_bar(as), aphase);
      aphase ^= 1u;
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * 64);
      const uint32_t buf = my_epi + (box_it & 1u) * kSuEpiBuf;
      float f[32];
      {
        uint32_t a[32], b[32];
        tmem_ld32(taddr, a);
        tmem_ld32(taddr + 32u, b);  // products with W_lo
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(a[j]) + __uint_as_float(b[j]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
#pragma unroll
      for (int q = 0; q < 8; ++q) {  // 128-bit broadcast loads of scale / shift
        const float4 sc = *reinterpret_cast<const float4*>(s_scale + 4 * q);
        const float4 sh = *reinterpret_cast<const float4*>(s_shift + 4 * q);
        f[4 * q + 0] = fmaf(f[4 * q + 0], sc.x, sh.x);
        f[4 * q + 1] = fmaf(f[4 * q + 1], sc.y, sh.y);
        f[4 * q + 2] = fmaf(f[4 * q + 2], sc.z, sh.z);
        f[4 * q + 3] = fmaf(f[4 * q + 3], sc.w, sh.w);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)  // ReLU and saturation are part of the conversion
        st_shared_v4(buf + wr_off + (((uint32_t)j ^ wr_swz) << 4), pack_h2_sat<true>(f[8 * j], f[8 * j + 1]),
                     pack_h2_sat<true>(f[8 * j + 2], f[8 * j + 3]), pack_h2_sat<true>(f[8 * j + 4], f[8 * j + 5]),
                     pack_h2_sat<true>(f[8 * j + 6], f[8 * j + 7]));
      epi_bar(1 + grp);  // box staged; also orders the reads of this buffer (two tiles ago) before these writes
      char* obase = reinterpret_cast<char*>(p.out) + (size_t)st_chunk * 16;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = st_row0 + 32 * i;
        const long long mo = m0 + r;
        if (mo < p.total) {
          const uint4 val = ld_shared_v4(buf + (uint32_t)r * 64u + (((uint32_t)st_chunk ^ (uint32_t)((r >> 1) & 3)) << 4));
          *reinterpret_cast<uint4*>(obase + (size_t)mo * 64) = val;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kSuMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
  }
}

// Host-side packing of the 32 x 27 stem weights into the 64 x 64 fp16 operand described at the top of this file.
// `scale` (folded BN, 32 floats) is divided by the per-channel power of two that brings the weights into fp16 range.
void pack_stem_umma(const float* w_k_major /*[27][32]*/, float* scale /*[32], in/out*/, __half* out /*[64][64]*/) {
  for (int i = 0; i < 64 * 64; ++i) out[i] = __float2half_rn(0.f);
  for (int co = 0; co < 32; ++co) {
    float maxabs = 0.f;
    for (int k = 0; k < 27; ++k) maxabs = fmaxf(maxabs, fabsf(w_k_major[k * 32 + co]));
    int e = 0;
    if (maxabs > 0.f && isfinite(maxabs)) e = 11 - ilogbf(maxabs);
    for (int k = 0; k < 27; ++k) {
      const float v = ldexpf(w_k_major[k * 32 + co], e);
      const __half h = __float2half_rn(v);
      const __half l = __float2half_rn(v - __half2float(h));
      out[co * 64 + k] = h;         // W_hi row: against x_hi ...
      out[co * 64 + 32 + k] = h;    // ... and against x_lo
      out[(32 + co) * 64 + k] = l;  // W_lo row
      out[(32 + co) * 64 + 32 + k] = l;
    }
    scale[co] = ldexpf(scale[co], -e);
  }
}

int launch_stem_umma(tise_ctx* ctx, const uint8_t* d_u8, int n, const __half* d_w16, const float* d_scale,
                     const float* d_shift, __half* d_out, cudaStream_t st) {
  static PerDeviceFlag configured;
  if (!configured.is(ctx, 1)) {
    TISE_CUDA(ctx, cudaFuncSetAttribute(stem_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSuSmem));
    configured.set(ctx, 1);
  }
  StemUmmaParams p;
  p.u8 = d_u8;
  p.u8_end = d_u8 + (size_t)n * kOutSize * kOutSize * 3;
  p.lut = ctx->d_lut;
  p.w16 = d_w16;
  p.scale = d_scale;
  p.shift = d_shift;
  p.out = d_out;
  p.total = (long long)n * kSuOut * kSuOut;
  if (p.total + 4 * kBlockM >= (1ll << 31)) return fail(ctx, TISE_ERR_INVALID, "stem: batch too large (%d images)", n);
  p.num_tiles = (int)((p.total + kBlockM - 1) / kBlockM);
  const int grid = p.num_tiles < ctx->num_sms ? p.num_tiles : ctx->num_sms;
  stem_umma_kernel<<<grid, kSuThreads, kSuSmem, st>>>(p);
  TISE_LAUNCH_CHECK(ctx);
  return TISE_OK;
}

}  // namespace tise
