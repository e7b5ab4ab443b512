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
//   scale/shift (folded BN) from shared memory, ReLU, convert to fp16 and stage 128 x 32-channel boxes in
//   bank-conflict-free swizzled shared memory; the same 128 threads then write the box out with
//   16-byte stores in which 4 consecutive lanes cover one pixel's 64 contiguous bytes (full sectors),
//   into channel slices of the NHWC destinations -- torch.cat of the Inception branches costs nothing.
//   Output channels may be split into segments with different destinations (sibling 1x1 convolutions
//   of a block run as ONE GEMM over the shared input).  (TMA stores of such narrow boxes were measured
//   at ~12 cycles per 64-byte row and made the epilogue the bottleneck; see profiles/.)
// * K blocks are grouped: one mbarrier stage carries `group` consecutive (tap, chunk) blocks, so layers with few
//   channels (small N, small K per tap) do not pay the ~500-cycle barrier round trip per 64-deep block.
// * Persistent: grid = #SMs, each CTA strides over (m tile, n tile) pairs; TMEM holds two
//   accumulators (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of tile i+1;
//   shared memory holds a 2-8 stage TMA<->MMA ring.
//
// Warp roles (192 threads): warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2-5 epilogue.
#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "umma_ptx.cuh"

namespace tise {

constexpr int kMaxStages = 8;
constexpr int kAccumCols = 256;  // TMEM columns per accumulator stage
constexpr int kNumThreads = 320;  // TMA producer warp, MMA warp, 2 epilogue groups of 4 warps
constexpr int kMaxNt = 256;
constexpr int kMaxKb = 96;      // K blocks per tile: taps * ceil(cin / 64) (<= 63 in Inception-v3)
constexpr int kMaxCout = 1344;  // widest (fused) layer: Mixed_7b 320 + 384 + 448 + 192
constexpr int kEpiBufBytes = 128 * 64;  // staged output box: 128 pixels x 64 bytes (32 fp16 or 16 fp32 channels)

struct UmmaParams {
  const float* scale;
  const float* shift;
  void* seg_out[kMaxSegs];   // destination of each output-channel segment (channel offset applied)
  int seg_stride[kMaxSegs];  // elements per destination pixel
  int seg_end[kMaxSegs];     // exclusive end column of the segment within the layer's cout
  int seg_relu[kMaxSegs];
  long long M;
  int ho_wo, wo, stride, pad_h, pad_w;
  int R, S, n_chunks, cin_pad;
  int kc;                  // channels per K chunk: 64, 32 or 16 (128-, 64-, 32-byte swizzled rows)
  int ksteps, last_ksteps; // K = 16 MMA steps per chunk (the last chunk of a tap may be shorter: cin = 48)
  int a_bytes, b_bytes;    // bytes of the A tile (128 x kc) and of one B plane (nt x kc) of a K block
  int swz_type, sbo;       // UMMA descriptor layout type and stride between 8-row groups for this kc
  int wide;                // hi | lo planes side by side form ONE MMA of N = 2 nt (nt <= 128, 2 passes)
  int onebox;              // hi | lo rows of a tile arrive as ONE TMA box of 2 nt rows (nt <= 128, 2 passes)
  int nt, n_ntiles, num_tiles;
  int passes, stages, stage_bytes;
  int group, kb_bytes;  // K blocks per pipeline stage, bytes of one K block (A + B planes)
  int cs;               // CTAs per cluster: 1, or 2 = CTA pair (cta_group::2 MMAs of M = 256 over two consecutive M tiles;
                        // each CTA stages its own activation rows and half of the weight rows)
  int m_tiles;
  int msub;             // M tiles (1, 2) per work item: they share every weight load and barrier round trip (cs == 1 only)
  int acc_cols;         // TMEM columns of one M tile's accumulator: 2 nt (wide) or nt
  int pdl;              // launched with programmatic stream serialization (see umma_ptx.cuh)
  unsigned long long* dbg_out;  // debug & 64: per-CTA cycle counters of the three roles (see launch_conv_umma)
  int debug;            // timing ablations (TISE_UMMA_DEBUG): 1 skip A loads, 2 skip B loads, 4 skip MMAs, 8 skip epilogue
};

// ----------------------------------------------------------------------------- MMA issue loop
struct MmaLoopArgs {
  uint32_t idesc, desc_hi, desc_lo0;
  uint32_t stage_step, kb_step;     // descriptor address-field steps (bytes >> 4) between stages / K blocks
  uint32_t b_off, bplane, a1_off;   // W_hi tile, W_lo plane and second A tile relative to the first A tile
  uint32_t acc_cols;
  int group, n_stages, kb_per_tile;
  bool uniform_ks;
  int ks_const;
  uint32_t bar_base, tmem_base;
  int first, count, step;           // work items first, first + step, ... < count
  int dbg;
  unsigned long long* dbg_out;
};

// mbarrier wait whose duration is added to `acc` in the instrumented (debug & 64) runs of the kDebug instantiation
template <bool kDebug>
__device__ __forceinline__ void timed_wait(uint32_t bar, uint32_t parity, int dbg, unsigned long long& acc) {
  if (kDebug && (dbg & 64)) {
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    acc += (unsigned long long)(clock64() - t0);
  } else {
    mbar_wait(bar, parity);
  }
}

// kTwoMma: W_lo is a second MMA into the same accumulator (otherwise hi | lo form one MMA of N = 2 nt, or 1 pass).
// kTwoM: two M tiles per work item.  kPair: cta_group::2 MMAs issued by the leader CTA of a pair.
// All 32 lanes run the loop; one elected lane issues (see the kernel).
template <bool kPair>
__device__ __forceinline__ void issue_mma(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t acc) {
  if (kPair) umma_f16_lh_pair(d, a_lo, b_lo, hi, idesc, acc);
  else umma_f16_lh(d, a_lo, b_lo, hi, idesc, acc);
}

template <bool kPair, bool kDebug, bool kTwoMma, bool kTwoM>
__device__ __forceinline__ void mma_loop(const MmaLoopArgs& a, const uint4* kbtab) {
  const uint32_t full0 = a.bar_base, empty0 = a.bar_base + 8u * kMaxStages, tfull0 = a.bar_base + 8u * (2 * kMaxStages),
                 tempty0 = a.bar_base + 8u * (2 * kMaxStages + 2);
  const uint32_t hi = a.desc_hi, idesc = a.idesc;
  int stage = 0;
  uint32_t phase = 0;
  int as = 0;
  uint32_t aphase = 0;
  unsigned long long w_full = 0, w_tempty = 0;
  const long long t_begin = kDebug ? clock64() : 0;
  for (int st = a.first; st < a.count; st += a.step) {
    timed_wait<kDebug>(tempty0 + 8u * (uint32_t)as, aphase ^ 1u, a.dbg, w_tempty);  // epilogue has drained this accumulator
    tc_fence_after();
    const uint32_t d0 = a.tmem_base + (uint32_t)(as * kAccumCols), d1 = d0 + a.acc_cols;
    // one iteration per pipeline stage (1-2 K blocks): one wait, one election, one commit
    for (int kb0 = 0; kb0 < a.kb_per_tile; kb0 += a.group) {
      const int n = min(a.group, a.kb_per_tile - kb0);
      timed_wait<kDebug>(full0 + 8u * (uint32_t)stage, phase, a.dbg, w_full);
      tc_fence_after();
      if (elect_one()) {
        // descriptors differ from K block to K block only in the 14-bit start-address field (low word)
        uint32_t alo = a.desc_lo0 + (uint32_t)stage * a.stage_step;
        for (int j = 0; j < n; ++j, alo += a.kb_step) {
          const int kb = kb0 + j;
          const int ksteps = a.uniform_ks ? a.ks_const : (int)kbtab[kb].w;
          const uint32_t blo = alo + a.b_off;
          // advance 16 fp16 = 32 bytes along K inside the swizzle row: +2 in the (>> 4) address field
#pragma unroll 4
          for (int k = 0; k < ksteps; ++k) {
            const uint32_t acc = (uint32_t)(kb | k);
            issue_mma<kPair>(d0, alo + 2u * k, blo + 2u * k, hi, idesc, acc);
            if (kTwoMma) issue_mma<kPair>(d0, alo + 2u * k, blo + a.bplane + 2u * k, hi, idesc, 1u);
            if (kTwoM) {  // second M tile against the same weights
              issue_mma<kPair>(d1, alo + a.a1_off + 2u * k, blo + 2u * k, hi, idesc, acc);
              if (kTwoMma) issue_mma<kPair>(d1, alo + a.a1_off + 2u * k, blo + a.bplane + 2u * k, hi, idesc, 1u);
            }
          }
        }
        // frees the smem stage once these MMAs have read it (pair: in both CTAs, each has its own producer)
        const uint32_t eb = empty0 + 8u * (uint32_t)stage;
        if (kDebug && !kPair && (a.dbg & 16)) mbar_arrive(eb);  // timing experiment only (needs debug & 4)
        else if (kPair) umma_commit_pair(eb); else umma_commit(eb);
        if (kb0 + n == a.kb_per_tile) {  // accumulator complete -> epilogue (pair: of both CTAs)
          const uint32_t tb = tfull0 + 8u * (uint32_t)as;
          if (kDebug && !kPair && (a.dbg & 32)) mbar_arrive(tb);
          else if (kPair) umma_commit_pair(tb); else umma_commit(tb);
        }
      }
      if (++stage == a.n_stages) {
        stage = 0;
        phase ^= 1u;
      }
    }
    if (++as == 2) {
      as = 0;
      aphase ^= 1u;
    }
  }
  if (kDebug && (a.dbg & 64) && (threadIdx.x & 31) == 0) {
    a.dbg_out[blockIdx.x * 8 + 2] = (unsigned long long)(clock64() - t_begin);
    a.dbg_out[blockIdx.x * 8 + 3] = w_full;
    a.dbg_out[blockIdx.x * 8 + 4] = w_tempty;
  }
}

// ----------------------------------------------------------------------------- kernel
template <bool kF32, bool kPair, bool kDebug>
__global__ void __launch_bounds__(kNumThreads, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 const __grid_constant__ CUtensorMap tm_b2, const __grid_constant__ CUtensorMap tm_bs,
                 const UmmaParams p) {
  // The producer and MMA warps are single instruction streams in which every dependent instruction costs its full
  // latency (~10 cycles, ~25 for a constant-bank load), so their loops must be short: a private copy of the
  // parameters keeps them in registers (the "memory" clobbers of the PTX wrappers would otherwise force a reload
  // from the constant bank on every use), and the timing-ablation branches exist only in the kDebug instantiation.
  const int dbg = kDebug ? p.debug : 0;
  constexpr int pcs = kPair ? 2 : 1;  // CTAs per work item
  const int msub = p.msub;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
  // layout: [stages x group x (A | B_hi | B_lo)] [2 staged output boxes] [scale cout | shift cout] [barriers] [tmem slot]
  const uint32_t epi_base = smem_base + (uint32_t)p.stages * (uint32_t)p.stage_bytes;
  const uint32_t ss_base = epi_base + 4u * kEpiBufBytes;  // two staged boxes per epilogue group
  const uint32_t bar_base = ss_base + 2u * kMaxCout * 4u;
  // barriers: full[stages], empty[stages], tmem_full[2], tmem_empty[2]; then the TMEM base address word
  auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(kMaxStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (uint32_t)(2 * kMaxStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(2 * kMaxStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 4);
  // per-K-block table {channel offset, tap s | tap r << 16, weight column, MMA K steps}: built once so that the
  // producer / MMA loops need no counters, divisions or branches per K block
  uint4* kbtab = reinterpret_cast<uint4*>(smem_raw + (tmem_slot + 16u - smem_u32(smem_raw)));
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
      mbar_init(tempty_bar(s), kPair ? 8 : 4);  // one arrive per epilogue warp (pair: of both CTAs, on the leader's barrier)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  for (int kb = threadIdx.x; kb < p.R * p.S * p.n_chunks; kb += blockDim.x) {
    const int ch = kb % p.n_chunks, tap = kb / p.n_chunks;
    const int r = tap / p.S, sx = tap - r * p.S;
    kbtab[kb] = make_uint4((uint32_t)(ch * p.kc), (uint32_t)sx | ((uint32_t)r << 16), (uint32_t)(kb * p.kc),
                           (uint32_t)((kDebug && (p.debug & 4)) ? 0 : (ch == p.n_chunks - 1 ? p.last_ksteps : p.ksteps)));
  }
  if (warp == 1) {
    if (kPair) {  // both CTAs of the pair issue it, with the same destination offset
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();  // the peer's barriers are initialised before any load or commit can reach them
  tc_fence_after();
  if (p.pdl) {  // everything above overlapped the previous kernel's tail; its outputs are needed from here on
    pdl_wait();
    pdl_trigger();
  }
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int crank = kPair ? (int)cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / pcs, num_clusters = gridDim.x / pcs;
  const int mg = pcs * msub;                                       // M tiles per work item
  const int num_super = ((p.m_tiles + mg - 1) / mg) * p.n_ntiles;  // (group of M tiles, N tile) work items

  const int kb_per_tile = p.R * p.S * p.n_chunks;
  const uint32_t a_bytes = (uint32_t)p.a_bytes, b_bytes = (uint32_t)p.b_bytes;
  const uint32_t a_tot = (uint32_t)msub * a_bytes;  // the A tiles of a K block lie side by side, then the weight planes
  // pair: each CTA stages half of the rows of every weight plane
  const uint32_t half_bytes = (uint32_t)(p.nt / 2 * p.kc * 2);
  const uint32_t kb_tx = ((dbg & 1) ? 0u : a_tot) + ((dbg & 2) ? 0u : (uint32_t)p.passes * (kPair ? half_bytes : b_bytes));
  const bool wide2 = p.wide && p.passes == 2;

  // Producer and MMA warps run their loops warp-uniformly (all 32 lanes wait on the mbarriers and compute the
  // descriptors, which therefore live in uniform registers); only the TMA / tcgen05 instructions themselves
  // are predicated on one elected lane.  A single-lane `if (lane == 0)` loop makes ptxas wrap every
  // uniform-datapath instruction in an election loop (~700 cycles per K block, measured).
  if (warp == 0) {
    // ===================================================================== TMA producer
    const int group = p.group, n_stages = p.stages, nt = p.nt;
    const uint32_t stage_bytes = (uint32_t)p.stage_bytes, kb_bytes = (uint32_t)p.kb_bytes;
    const bool two_loads = p.passes == 2 && !p.onebox, onebox = p.onebox != 0, two_pass = p.passes == 2;
    const int half_rows = p.nt / 2;
    // pair: every load of both CTAs completes on the leader's full barrier, which alone carries the expect_tx
    const uint32_t full_base = kPair ? mapa_u32(bar_base, 0) : bar_base;
    int stage = 0;
    uint32_t phase = 0;
    unsigned long long w_empty = 0, n_kb = 0;
    const long long t_begin = kDebug ? clock64() : 0;
    for (int st = cluster_id; st < num_super; st += num_clusters) {
      const int m_group = st / p.n_ntiles, n_tile = st - m_group * p.n_ntiles;
      const int m_tile = (m_group * pcs + crank) * msub;  // may be past the end in the last group: loads zero-fill, nothing is stored
      int img[2], wc[2], hc[2];  // base pixel (tap 0,0) of each M tile
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
        // 32-bit arithmetic (the host guarantees M + 512 < 2^31): a 64-bit division is a ~100-instruction dependent
        // chain, and this warp's instruction stream is what bounds most layers
        const unsigned m0 = (unsigned)(m_tile + sub) * (unsigned)kBlockM;
        img[sub] = (int)(m0 / (unsigned)p.ho_wo);
        const int rem = (int)(m0 - (unsigned)img[sub] * (unsigned)p.ho_wo);
        const int oy = rem / p.wo, ox = rem - oy * p.wo;
        wc[sub] = ox * p.stride - p.pad_w;
        hc[sub] = oy * p.stride - p.pad_h;
      }
      const int brow = n_tile * 2 * nt;  // first hi row of this tile's weight block; the lo rows follow
      // one iteration per pipeline stage (1-2 K blocks): one wait, one election, one expect_tx
      for (int kb0 = 0; kb0 < kb_per_tile; kb0 += group) {
        const int n = min(group, kb_per_tile - kb0);
        timed_wait<kDebug>(empty_bar(stage), phase ^ 1u, dbg, w_empty);
        if (kDebug) n_kb += (unsigned long long)n;
        if (elect_one()) {
          const uint32_t fb = full_base + 8u * (uint32_t)stage;
          if (crank == 0) {
            if (kb_tx) mbar_expect_tx(full_bar(stage), (uint32_t)n * kb_tx * (uint32_t)pcs); else mbar_arrive(full_bar(stage));
          }
          uint32_t sa = smem_base + (uint32_t)stage * stage_bytes;
          for (int j = 0; j < n; ++j, sa += kb_bytes) {
            const uint4 e = kbtab[kb0 + j];
            const uint16_t ow = (uint16_t)(e.y & 0xffffu), oh = (uint16_t)(e.y >> 16);
            if (!(kDebug && (dbg & 1))) {
              if (kPair) {
                tma_load_im2col_pair(&tm_a, sa, fb, (int)e.x, wc[0], hc[0], img[0], ow, oh);
                if (msub == 2) tma_load_im2col_pair(&tm_a, sa + a_bytes, fb, (int)e.x, wc[1], hc[1], img[1], ow, oh);
              } else {
                tma_load_im2col(&tm_a, sa, fb, (int)e.x, wc[0], hc[0], img[0], ow, oh);
                if (msub == 2) tma_load_im2col(&tm_a, sa + a_bytes, fb, (int)e.x, wc[1], hc[1], img[1], ow, oh);
              }
            }
            if (!(kDebug && (dbg & 2))) {
              if (kPair) {
                if (wide2) {  // N = 2 nt: the leader holds the W_hi rows, its peer the W_lo rows
                  tma_load_2d_pair(&tm_b, sa + a_tot, fb, (int)e.z, brow + crank * nt);
                } else {      // N = nt per pass: rows [0, nt / 2) of each plane in the leader, the rest in its peer
                  tma_load_2d_pair(&tm_bs, sa + a_tot, fb, (int)e.z, brow + crank * half_rows);
                  if (two_pass) tma_load_2d_pair(&tm_bs, sa + a_tot + half_bytes, fb, (int)e.z, brow + nt + crank * half_rows);
                }
              } else if (!two_loads) {
                tma_load_2d(onebox ? &tm_b2 : &tm_b, sa + a_tot, fb, (int)e.z, brow);  // onebox: hi and lo rows together
              } else {
                tma_load_2d(&tm_b, sa + a_tot, fb, (int)e.z, brow);
                tma_load_2d(&tm_b, sa + a_tot + b_bytes, fb, (int)e.z, brow + nt);
              }
            }
          }
        }
        if (++stage == n_stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
    if (kDebug && (dbg & 64) && lane == 0) {
      p.dbg_out[blockIdx.x * 8 + 0] = (unsigned long long)(clock64() - t_begin);
      p.dbg_out[blockIdx.x * 8 + 1] = w_empty;
      p.dbg_out[blockIdx.x * 8 + 7] = n_kb;
    }
  } else if (warp == 1 && crank == 0) {
    // ===================================================================== MMA issuer (pair: the leader CTA only)
    const bool two_mma = p.passes == 2 && !wide2;
    MmaLoopArgs ma;
    ma.idesc = make_instr_desc(wide2 ? 2 * p.nt : p.nt, kPair ? 256 : 128);
    const uint64_t dhi = smem_desc_hi(p.swz_type, p.sbo);
    ma.desc_hi = (uint32_t)(dhi >> 32);
    ma.desc_lo0 = (uint32_t)make_smem_desc(smem_base, dhi);  // low word of the descriptor of stage 0, K block 0
    ma.stage_step = (uint32_t)p.stage_bytes >> 4;
    ma.kb_step = (uint32_t)p.kb_bytes >> 4;
    ma.b_off = a_tot >> 4;
    ma.bplane = (kPair ? half_bytes : b_bytes) >> 4;
    ma.a1_off = a_bytes >> 4;
    ma.acc_cols = (uint32_t)p.acc_cols;
    ma.group = p.group;
    ma.n_stages = p.stages;
    ma.kb_per_tile = kb_per_tile;
    // K steps per block: constant unless the channel count leaves a short last chunk (then read from the table)
    ma.uniform_ks = p.n_chunks == 1 || p.last_ksteps == p.ksteps;
    ma.ks_const = ((kDebug && (p.debug & 4)) ? 0 : (p.n_chunks == 1 ? p.last_ksteps : p.ksteps));
    ma.bar_base = bar_base;
    ma.tmem_base = tmem_base;
    ma.first = cluster_id;
    ma.count = num_super;
    ma.step = num_clusters;
    ma.dbg = dbg;
    ma.dbg_out = p.dbg_out;
    // one instantiation per (W_lo as a second MMA, second M tile): the loop body is straight-line code
    if (two_mma) {
      if (msub == 2) mma_loop<kPair, kDebug, true, true>(ma, kbtab);
      else mma_loop<kPair, kDebug, true, false>(ma, kbtab);
    } else {
      if (msub == 2) mma_loop<kPair, kDebug, false, true>(ma, kbtab);
      else mma_loop<kPair, kDebug, false, false>(ma, kbtab);
    }
  } else if (warp >= 2) {
    // ===================================================================== epilogue (warps 2..5 and 6..9)
    // Two groups of 4 warps, one per TMEM accumulator stage: group g drains the work items with (index & 1) == g, so
    // two tiles are in the epilogue at once (the single group was the bottleneck of every layer with few K blocks).
    const int grp = (warp - 2) >> 2;
    constexpr int W = kF32 ? 16 : 32;  // channels per staged box (64-byte rows)
    constexpr int CPC = kF32 ? 4 : 8;  // channels per 16-byte chunk
    const int quarter = warp & 3;      // TMEM lanes [32 * quarter, 32 * quarter + 32) are accessible to this warp
    const int row = quarter * 32 + lane;
    const int etid = threadIdx.x - 64 - grp * 128;  // 0..127 within the group
    float* s_scale = reinterpret_cast<float*>(smem_raw + (ss_base - smem_u32(smem_raw)));
    float* s_shift = s_scale + kMaxCout;
    // the whole layer's folded-BN scale / shift, once per CTA (a per-tile reload exposed ~1 us of L2 latency per tile);
    // both groups write all (identical) values, so each only needs its own barrier
    for (int c = etid; c < p.n_ntiles * p.nt; c += 128) {
      s_scale[c] = __ldg(p.scale + c);
      s_shift[c] = __ldg(p.shift + c);
    }
    epi_bar(1 + grp);
    const uint32_t my_epi = epi_base + (uint32_t)grp * 2u * kEpiBufBytes;
    // hands an accumulator stage back to the MMA warp -- which, in a pair, lives in the leader CTA
    const uint32_t tempty_leader = kPair ? mapa_u32(tempty_bar(0), 0) : 0u;
    auto release_accum = [&](int s) {
      if (kPair && crank != 0) mbar_arrive_cluster(tempty_leader + 8u * (uint32_t)s);
      else mbar_arrive(tempty_bar(s));
    };
    // staged box: row r at r * 64 bytes, its 16-byte chunk j stored at slot j ^ ((r >> 1) & 3): conflict-free both
    // for the row-per-thread writes and for the 4-lanes-per-row reads of the store phase
    const uint32_t wr_off = (uint32_t)row * 64u;
    const uint32_t wr_swz = (uint32_t)((row >> 1) & 3);
    const int st_chunk = etid & 3, st_row0 = etid >> 2;  // store phase: chunk st_chunk of rows st_row0 + 32 i
    const int as = grp;  // this group's accumulator stage
    uint32_t aphase = 0;
    uint32_t box_it = 0;
    unsigned long long w_tfull = 0;
    const long long t_begin = kDebug ? clock64() : 0;
    for (int st = cluster_id + grp * num_clusters; st < num_super; st += 2 * num_clusters) {
      const int m_group = st / p.n_ntiles, n_tile = st - m_group * p.n_ntiles;
      const int co0 = n_tile * p.nt;
      timed_wait<kDebug>(tfull_bar(as), aphase, dbg, w_tfull);
      tc_fence_after();
      if (dbg & 8) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release_accum(as);
      }
      for (int sub = 0; sub < msub; ++sub) {
      const long long m0 = (long long)((m_group * pcs + crank) * msub + sub) * kBlockM;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * kAccumCols + sub * p.acc_cols);
      for (int c0 = 0; c0 < ((dbg & 8) ? 0 : p.nt); c0 += W, ++box_it) {
        const uint32_t buf = my_epi + (box_it & 1u) * kEpiBufBytes;
        float f[W];
        {
          uint32_t v[W];
          if (kDebug && (dbg & 256)) {  // timing ablation: no TMEM reads
#pragma unroll
            for (int j = 0; j < W; ++j) v[j] = 0;
          } else if (kF32) {
            tmem_ld16(taddr + (uint32_t)c0, reinterpret_cast<uint32_t(&)[16]>(v));
          } else {
            tmem_ld32(taddr + (uint32_t)c0, reinterpret_cast<uint32_t(&)[32]>(v));
          }
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < W; ++j) f[j] = __uint_as_float(v[j]);
          if (wide2 && !(kDebug && (dbg & 256))) {  // the W_lo products sit nt columns to the right
            if (kF32) {
              tmem_ld16(taddr + (uint32_t)(p.nt + c0), reinterpret_cast<uint32_t(&)[16]>(v));
            } else {
              tmem_ld32(taddr + (uint32_t)(p.nt + c0), reinterpret_cast<uint32_t(&)[32]>(v));
            }
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < W; ++j) f[j] += __uint_as_float(v[j]);
          }
        }
        if (c0 + W >= p.nt && sub == msub - 1) {  // accumulators fully read: hand them back to the MMA warp before the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) release_accum(as);
        }
        // scale / shift / ReLU (the ReLU flag can change every 16 channels: segment boundaries); for fp16 outputs the
        // ReLU is part of the conversion instruction below
        bool relu_h[2] = {false, false};
#pragma unroll
        for (int h = 0; h < W / 16; ++h) {
          const int col = co0 + c0 + 16 * h;
          int sg = 0;
          while (sg < kMaxSegs - 1 && col >= p.seg_end[sg]) ++sg;
          const bool relu = p.seg_relu[sg] != 0;
          relu_h[h] = relu;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 sc = *reinterpret_cast<const float4*>(s_scale + co0 + c0 + 16 * h + 4 * q);
            const float4 sh = *reinterpret_cast<const float4*>(s_shift + co0 + c0 + 16 * h + 4 * q);
            float* y = f + 16 * h + 4 * q;
            y[0] = fmaf(y[0], sc.x, sh.x);
            y[1] = fmaf(y[1], sc.y, sh.y);
            y[2] = fmaf(y[2], sc.z, sh.z);
            y[3] = fmaf(y[3], sc.w, sh.w);
            if (kF32 && relu) {
              y[0] = fmaxf(y[0], 0.f);
              y[1] = fmaxf(y[1], 0.f);
              y[2] = fmaxf(y[2], 0.f);
              y[3] = fmaxf(y[3], 0.f);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t dst = buf + wr_off + (((uint32_t)j ^ wr_swz) << 4);
          if (kF32)
            st_shared_v4(dst, __float_as_uint(f[4 * j]), __float_as_uint(f[4 * j + 1]), __float_as_uint(f[4 * j + 2]),
                         __float_as_uint(f[4 * j + 3]));
          else if (relu_h[(8 * j) / 16])
            st_shared_v4(dst, pack_h2_sat<true>(f[8 * j], f[8 * j + 1]), pack_h2_sat<true>(f[8 * j + 2], f[8 * j + 3]),
                         pack_h2_sat<true>(f[8 * j + 4], f[8 * j + 5]), pack_h2_sat<true>(f[8 * j + 6], f[8 * j + 7]));
          else
            st_shared_v4(dst, pack_h2_sat<false>(f[8 * j], f[8 * j + 1]), pack_h2_sat<false>(f[8 * j + 2], f[8 * j + 3]),
                         pack_h2_sat<false>(f[8 * j + 4], f[8 * j + 5]), pack_h2_sat<false>(f[8 * j + 6], f[8 * j + 7]));
        }
        epi_bar(1 + grp);  // box staged; also orders the reads of this buffer (box_it - 2) before these writes
        // store phase: thread -> (chunk, 4 rows); 4 consecutive lanes write one pixel's 64 contiguous bytes
        const int ccol = c0 + st_chunk * CPC;
        if (ccol < p.nt && !(kDebug && (dbg & 128))) {  // (debug & 128: timing ablation without the global stores)
          const int gcol = co0 + ccol;
          int sg = 0;
          while (sg < kMaxSegs - 1 && gcol >= p.seg_end[sg]) ++sg;
          const int scol = gcol - (sg ? p.seg_end[sg - 1] : 0);
          char* obase = reinterpret_cast<char*>(p.seg_out[sg]) + (size_t)scol * (kF32 ? 4 : 2);
          const size_t pitch = (size_t)p.seg_stride[sg] * (kF32 ? 4 : 2);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = st_row0 + 32 * i;
            const long long m = m0 + r;
            if (m < p.M) {
              const uint4 v = ld_shared_v4(buf + (uint32_t)r * 64u + (((uint32_t)st_chunk ^ (uint32_t)((r >> 1) & 3)) << 4));
              *reinterpret_cast<uint4*>(obase + (size_t)m * pitch) = v;
            }
          }
        }
      }
      }
      aphase ^= 1u;
    }
    if (kDebug && (dbg & 64) && threadIdx.x == 64) {
      p.dbg_out[blockIdx.x * 8 + 5] = (unsigned long long)(clock64() - t_begin);
      p.dbg_out[blockIdx.x * 8 + 6] = w_tfull;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();  // no CTA leaves while its peer may still signal its barriers or read its operands
  if (warp == 1) {
    tc_fence_after();
    if (kPair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
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

static CUtensorMapSwizzle swizzle_of(int kc) {
  return kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// Programmatic dependent launch between consecutive tcgen05 conv kernels (default on; TISE_PDL=0 disables).
bool umma_pdl_mode() {
  static int pdl = -1;
  if (pdl < 0) pdl = (getenv("TISE_PDL") && atoi(getenv("TISE_PDL")) == 0) ? 0 : 1;
  return pdl != 0;
}

// CTA-pair (cta_group::2) execution is the default; TISE_UMMA_PAIR=0 selects the single-CTA form of the kernel.
bool umma_pair_mode() {
  static int pair = -1;
  if (pair < 0) pair = (getenv("TISE_UMMA_PAIR") && atoi(getenv("TISE_UMMA_PAIR")) == 0) ? 0 : 1;
  return pair != 0;
}

int build_umma_plan(tise_ctx* ctx, const ConvLayer& L, const ConvArgs& a, UmmaPlan* plan) {
  int rc = load_driver_entry_points(ctx);
  if (rc) return rc;
  if (L.cout > kMaxCout) return fail(ctx, TISE_ERR_INVALID, "umma conv: unsupported cout %d", L.cout);
  if (L.R * L.S * (L.cin_pad / L.kc) > kMaxKb) return fail(ctx, TISE_ERR_INVALID, "umma conv: too many K blocks");
  if (a.in_stride % 8 || ((uintptr_t)a.in & 15))
    return fail(ctx, TISE_ERR_INVALID, "umma conv: input view is not 16-byte aligned");
  if (a.n_segs < 1 || a.n_segs > kMaxSegs) return fail(ctx, TISE_ERR_INVALID, "umma conv: %d segments", a.n_segs);

  // ---- A: im2col map over the NHWC input view, dims (C, W, H, N), box = kc channels x 128 pixels
  {
    cuuint64_t dims[4] = {(cuuint64_t)L.cin, (cuuint64_t)a.w, (cuuint64_t)a.h, (cuuint64_t)a.n};
    const cuuint64_t pitch_w = (cuuint64_t)(a.in_row_pitch ? a.in_row_pitch : a.w);  // physical pixels per row
    cuuint64_t strides[3] = {(cuuint64_t)a.in_stride * 2, (cuuint64_t)a.in_stride * 2 * pitch_w,
                             (cuuint64_t)a.in_stride * 2 * pitch_w * a.h};
    int lower[2] = {-L.pad_w, -L.pad_h};
    int upper[2] = {L.pad_w - (L.S - 1), L.pad_h - (L.R - 1)};
    cuuint32_t estr[4] = {1, (cuuint32_t)L.stride, (cuuint32_t)L.stride, 1};
    CUresult r = g_encode_im2col(&plan->tm_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void*)a.in, dims, strides, lower,
                                 upper, (cuuint32_t)L.kc, (cuuint32_t)kBlockM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 swizzle_of(L.kc), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
      return fail(ctx, TISE_ERR_CUDA, "cuTensorMapEncodeIm2col failed (%d) for %dx%dx%dx%d k%dx%d s%d", (int)r, a.n, a.h,
                  a.w, L.cin, L.R, L.S, L.stride);
    // Driver <= 13.1 sets a flag that breaks im2col maps over tensors smaller than 128 KiB; CUTLASS
    // (cute/atom/copy_traits_sm90_im2col.hpp) clears bit 21 of the second 64-bit word in that case.
    int drv = 0;
    cudaDriverGetVersion(&drv);
    const size_t bytes = (size_t)a.n * a.h * pitch_w * a.in_stride * 2;
    if (drv <= 13010 && bytes < 131072) reinterpret_cast<uint64_t*>(&plan->tm_a)[1] &= ~(1ull << 21);
  }
  // ---- B: packed weights [2 cout][ktot]; boxes of kc x nt rows (one plane) and kc x 2 nt rows (hi + lo)
  for (int two = 0; two < 2; ++two) {
    const int rows = two ? 2 * L.nt : L.nt;
    if (rows > 256) {
      plan->tm_b2 = plan->tm_b;
      continue;
    }
    cuuint64_t dims[2] = {(cuuint64_t)L.ktot, (cuuint64_t)2 * L.cout};
    cuuint64_t strides[1] = {(cuuint64_t)L.ktot * 2};
    cuuint32_t box[2] = {(cuuint32_t)L.kc, (cuuint32_t)rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(two ? &plan->tm_b2 : &plan->tm_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)L.d_w, dims,
                                strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_of(L.kc),
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, TISE_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for weights", (int)r);
  }
  // weight slices of nt / 2 rows for the cluster multicast
  {
    cuuint64_t dims[2] = {(cuuint64_t)L.ktot, (cuuint64_t)2 * L.cout};
    cuuint64_t strides[1] = {(cuuint64_t)L.ktot * 2};
    cuuint32_t box[2] = {(cuuint32_t)L.kc, (cuuint32_t)(L.nt / 2)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(&plan->tm_bs, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)L.d_w, dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_of(L.kc), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, TISE_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for weight slices", (int)r);
  }
  int csum = 0;
  for (int i = 0; i < a.n_segs; ++i) {
    if (a.segs[i].c % 16) return fail(ctx, TISE_ERR_INVALID, "umma conv: segment of %d channels", a.segs[i].c);
    if (((uintptr_t)a.segs[i].out & 15) || (a.segs[i].out_stride * (a.out_f32 ? 4 : 2)) % 16)
      return fail(ctx, TISE_ERR_INVALID, "umma conv: output slice is not 16-byte aligned");
    csum += a.segs[i].c;
  }
  if (csum != L.cout) return fail(ctx, TISE_ERR_INVALID, "umma conv: segments cover %d of %d channels", csum, L.cout);
  plan->valid = true;
  return TISE_OK;
}

int launch_conv_umma(tise_ctx* ctx, const ConvLayer& L, const ConvArgs& a, const UmmaPlan& plan, int passes,
                     cudaStream_t st) {
  if (!plan.valid) return fail(ctx, TISE_ERR_STATE, "umma conv: plan not built");
  UmmaParams p;
  p.scale = L.d_scale;
  p.shift = (passes == 1 && L.d_shift1) ? L.d_shift1 : L.d_shift;
  p.M = (long long)a.n * a.ho * a.wo;
  if (p.M + 4 * kBlockM >= (1ll << 31)) return fail(ctx, TISE_ERR_INVALID, "umma conv: batch too large (%lld output pixels)", p.M);
  int end = 0;
  for (int i = 0; i < kMaxSegs; ++i) {
    const ConvSeg& sg = a.segs[i < a.n_segs ? i : a.n_segs - 1];
    if (i < a.n_segs) end += sg.c;
    p.seg_out[i] = sg.out;
    p.seg_stride[i] = sg.out_stride;
    p.seg_end[i] = i < a.n_segs ? end : 0x7fffffff;
    p.seg_relu[i] = sg.relu;
  }
  p.ho_wo = a.ho * a.wo;
  p.wo = a.wo;
  p.stride = L.stride;
  p.pad_h = L.pad_h;
  p.pad_w = L.pad_w;
  p.R = L.R;
  p.S = L.S;
  p.kc = L.kc;
  p.n_chunks = L.cin_pad / L.kc;
  p.ksteps = L.kc / 16;
  p.last_ksteps = (L.cin - (p.n_chunks - 1) * L.kc + 15) / 16;
  p.cin_pad = L.cin_pad;
  p.nt = L.nt;
  p.n_ntiles = L.cout / p.nt;
  p.a_bytes = kBlockM * L.kc * 2;
  p.b_bytes = L.nt * L.kc * 2;
  p.swz_type = L.kc == 64 ? 2 : (L.kc == 32 ? 4 : 6);
  p.sbo = 8 * L.kc * 2;
  const int m_tiles = cdiv(p.M, kBlockM);
  p.m_tiles = m_tiles;
  p.num_tiles = m_tiles * p.n_ntiles;
  // CTA pairs (cta_group::2): one MMA of M = 256 spans two SMs, each staging its own 128 activation rows and HALF of
  // the weight rows.  Shared-memory traffic (TMA fills + operand reads, 128 B/clk/SM) is what bounds the single-CTA
  // form of this kernel (measured: profiles/); the pair halves the weight share of both.
  const int cs = umma_pair_mode() ? 2 : 1;
  p.cs = cs;
  p.passes = passes;
  // Two M tiles (256 output pixels) per work item share every weight load, barrier round trip and commit: the
  // TMA fill rate out of L2 (~60-85 B/clk/SM measured), not the tensor pipe, bounds M = 128 tiles, whose weight
  // bytes (2 planes) outweigh the activation bytes.  Needs both accumulators of a stage in 256 TMEM columns.
  static int msub_env = -1;
  if (msub_env < 0) msub_env = getenv("TISE_UMMA_MSUB") ? atoi(getenv("TISE_UMMA_MSUB")) : 0;
  int msub = 1;
  if (cs == 1 && 2 * p.nt <= kAccumCols && cdiv(m_tiles, 2) * p.n_ntiles >= ctx->num_sms) msub = 2;  // enough work items for every SM
  // pairs: two M tiles per CTA (512 output pixels per work item) where a K block would otherwise hold < 300 tensor
  // cycles against ~600 cycles of producer / MMA-issue instructions (nt <= 64: the 5x5 convolutions)
  if (cs == 2 && passes == 2 && 4 * p.nt <= kAccumCols && cdiv(m_tiles, 4) * p.n_ntiles >= ctx->num_sms / 2) msub = 2;
  // single pass: the accumulator of an M tile takes nt columns, so two M tiles fit for nt <= 128.  Measured per layer at
  // 1024 images (profiles/layers_r02_msub_sweep.txt): conv 3b 0.47 -> 0.38 ms, the 5x5 convolutions 0.42 -> 0.36, the
  // 35 x 35 3x3 convolutions 0.20 -> 0.17 / 0.33 -> 0.27, 17 x 17 128-channel 1x7 0.085 -> 0.074; nt > 128 unchanged.
  // (TISE_UMMA_MSUB1P=<largest nt> overrides; 0 disables.)  Each M tile still accumulates its K blocks in the same
  // order, so the output bits do not depend on msub.
  static int msub1p = -1;
  if (msub1p < 0) msub1p = getenv("TISE_UMMA_MSUB1P") ? atoi(getenv("TISE_UMMA_MSUB1P")) : 128;
  if (cs == 2 && passes == 1 && p.nt <= msub1p && 2 * p.nt <= kAccumCols && cdiv(m_tiles, 4) * p.n_ntiles >= ctx->num_sms / 2) msub = 2;
  if (msub_env == 1) msub = 1;
  p.msub = msub;
  p.onebox = (passes == 2 && 2 * p.nt <= 256) ? 1 : 0;
  // one MMA of N = 2 nt per K step where TMEM has room for it (pair: nt <= 128; single CTA: nt <= 64, so that two M
  // tiles fit); decided by nt alone, so that the summation order -- and with it every output bit -- does not depend
  // on the batch size
  p.wide = (p.onebox && (cs == 2 ? 2 : 4) * p.nt <= kAccumCols) ? 1 : 0;
  p.acc_cols = p.wide ? 2 * p.nt : p.nt;
  p.kb_bytes = cs == 2 ? msub * p.a_bytes + passes * (p.nt / 2) * L.kc * 2 : msub * p.a_bytes + passes * p.b_bytes;
  static int debug_env = -1;
  if (debug_env < 0) debug_env = getenv("TISE_UMMA_DEBUG") ? atoi(getenv("TISE_UMMA_DEBUG")) : 0;
  p.debug = debug_env;
  p.dbg_out = nullptr;
  static unsigned long long* d_dbg = nullptr;
  if (p.debug & 64) {  // role cycle counters: one synchronous report per launch on stderr
    if (!d_dbg) TISE_CUDA(ctx, cudaMalloc(&d_dbg, 1024 * 8 * sizeof(unsigned long long)));
    TISE_CUDA(ctx, cudaMemsetAsync(d_dbg, 0, 1024 * 8 * sizeof(unsigned long long), st));
    p.dbg_out = d_dbg;
  }

  static PerDeviceFlag configured_smem;
  const int smem_limit = ctx->smem_optin > 0 ? ctx->smem_optin : 232448;
  if (!configured_smem.is(ctx, smem_limit)) {
    const void* kernels[8] = {(const void*)conv_umma_kernel<false, false, false>, (const void*)conv_umma_kernel<true, false, false>,
                              (const void*)conv_umma_kernel<false, true, false>,  (const void*)conv_umma_kernel<true, true, false>,
                              (const void*)conv_umma_kernel<false, false, true>,  (const void*)conv_umma_kernel<true, false, true>,
                              (const void*)conv_umma_kernel<false, true, true>,   (const void*)conv_umma_kernel<true, true, true>};
    for (int i = 0; i < 8; ++i)
      TISE_CUDA(ctx, cudaFuncSetAttribute(kernels[i], cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit));
    configured_smem.set(ctx, smem_limit);
  }
  const int reserve = 1024 /*alignment*/ + 4 * kEpiBufBytes + 2 * kMaxCout * 4 + 8 * (2 * kMaxStages + 4) + 16 + 16 * kMaxKb;
  const int budget = smem_limit - reserve;
  // K blocks per pipeline stage.  The producer and MMA warps are single instruction streams (~10 cycles per dependent
  // instruction, measured), so a K block with little tensor work (small nt) cannot hide its own barrier wait /
  // expect_tx / commit: two K blocks share a stage when one holds < ~600 tensor cycles and 3 stages still fit.
  const int kb_per_tile = L.R * L.S * p.n_chunks;
  const int kb_cycles = msub * p.ksteps * passes * p.nt / 2;  // tcgen05 floor: M=128 (per SM) x N x K=16 takes N / 2 cycles
  int group = (kb_cycles < 600 && 2 * p.kb_bytes * 3 <= budget) ? 2 : 1;
  if (group > kb_per_tile) group = kb_per_tile;
  const char* genv = getenv("TISE_UMMA_GROUP");
  if (genv && atoi(genv) > 0 && atoi(genv) * p.kb_bytes * 2 <= budget) group = atoi(genv) < kb_per_tile ? atoi(genv) : kb_per_tile;
  p.group = group;
  p.stage_bytes = group * p.kb_bytes;
  int stages = budget / p.stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return fail(ctx, TISE_ERR_INVALID, "umma conv: tile does not fit shared memory");
  p.stages = stages;
  const int num_super = cdiv(m_tiles, cs * msub) * p.n_ntiles;
  // always request the full opt-in size: one CTA per SM, so the 512-column TMEM allocation never contends
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(ctx->num_sms / cs * cs));
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = (size_t)smem_limit;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  p.pdl = umma_pdl_mode() && !ctx->profiling ? 1 : 0;  // (per-launch profiling events would serialise the stream anyway)
  if (p.pdl) {
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  // the schedule is static (persistent CTAs), so the grid must not exceed the clusters that can be co-resident
  static int max_clusters[5] = {0, 0, 0, 0, 0};
  if (!max_clusters[cs]) {
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, conv_umma_kernel<false, true, false>, &cfg) != cudaSuccess || nc < 1) nc = ctx->num_sms / cs;
    max_clusters[cs] = nc;
  }
  int clusters = max_clusters[cs];
  if (clusters > ctx->num_sms / cs) clusters = ctx->num_sms / cs;
  if (clusters > num_super) clusters = num_super;
  cfg.gridDim = dim3((unsigned)(clusters * cs));
  static int verbose = -1;
  if (verbose < 0) verbose = getenv("TISE_UMMA_VERBOSE") ? 1 : 0;
  if (verbose)
    fprintf(stderr, "conv_umma %dx%d cin%d cout%d k%dx%d: nt %d kc %d wide %d msub %d group %d stages %d stage_bytes %d cs %d grid %d tiles %d\n",
            a.h, a.w, L.cin, L.cout, L.R, L.S, p.nt, p.kc, p.wide, p.msub, p.group, p.stages, p.stage_bytes, cs, clusters * cs,
            p.num_tiles);
  const CUtensorMap& tm_bs = plan.tm_bs;
  cudaError_t le;
#define TISE_LAUNCH_UMMA(F32, CL, DBG) \
  le = cudaLaunchKernelEx(&cfg, conv_umma_kernel<F32, CL, DBG>, plan.tm_a, plan.tm_b, plan.tm_b2, tm_bs, p)
  const int variant = (a.out_f32 ? 1 : 0) | (cs > 1 ? 2 : 0) | (p.debug ? 4 : 0);
  switch (variant) {
    case 0: TISE_LAUNCH_UMMA(false, false, false); break;
    case 1: TISE_LAUNCH_UMMA(true, false, false); break;
    case 2: TISE_LAUNCH_UMMA(false, true, false); break;
    case 3: TISE_LAUNCH_UMMA(true, true, false); break;
    case 4: TISE_LAUNCH_UMMA(false, false, true); break;
    case 5: TISE_LAUNCH_UMMA(true, false, true); break;
    case 6: TISE_LAUNCH_UMMA(false, true, true); break;
    default: TISE_LAUNCH_UMMA(true, true, true); break;
  }
#undef TISE_LAUNCH_UMMA
  if (le != cudaSuccess) return fail(ctx, TISE_ERR_CUDA, "conv_umma_kernel launch failed: %s", cudaGetErrorString(le));
  if (p.debug & 64) {
    static unsigned long long h[1024 * 8];
    TISE_CUDA(ctx, cudaStreamSynchronize(st));
    TISE_CUDA(ctx, cudaMemcpy(h, d_dbg, sizeof(h), cudaMemcpyDeviceToHost));
    double sum[8] = {0};
    const int nb = clusters * cs;
    for (int b = 0; b < nb; ++b)
      for (int j = 0; j < 8; ++j) sum[j] += (double)h[b * 8 + j];
    const double kb = sum[7] > 0 ? sum[7] : 1;
    fprintf(stderr,
            "umma_roles %dx%d cin%d cout%d k%dx%d s%d nt%d msub%d: kb/cta %.0f | cycles per K block: producer %.0f (wait empty %.0f) "
            "mma %.0f (wait full %.0f, wait tmem %.0f) epilogue %.0f (wait accum %.0f)\n",
            a.h, a.w, L.cin, L.cout, L.R, L.S, L.stride, p.nt, p.msub, sum[7] / nb, sum[0] / kb, sum[1] / kb, sum[2] / kb,
            sum[3] / kb, sum[4] / kb, sum[5] / kb, sum[6] / kb);
  }
  TISE_LAUNCH_CHECK(ctx);
  return TISE_OK;
}

}  // namespace tise
