// "Flat" tcgen05 convolution for stride-1 layers with few channels (Conv2d_2a_3x3, Conv2d_2b_3x3 of the stem).
//
// conv_umma_kernel feeds the tensor cores one (filter tap, 64-channel chunk) at a time through TMA im2col loads;
// with 32 input channels and 32/64 output channels a K block holds ~100 tensor cycles of work but costs ~500
// cycles of producer / MMA-issue instructions (single instruction streams), so those layers ran at 10-18 %
// tensor-pipe utilisation (profiles/).  Here the convolution is evaluated on the *flattened, padded input grid*:
//
//     out[v] = sum_{r,s} in[v + r * Wp + s] . W[r,s]        v = (n * Hp + y) * Wp + x
//
// so every filter tap is the SAME run of pixels shifted by r * Wp + s rows.  A work item is T consecutive
// 128-pixel tiles (T * 128 flat positions): the producer brings ONE contiguous run of T * 128 + (R - 1) * Wp + S - 1
// pixels into a pipeline stage (a few plain 2D TMA boxes), the MMA warp issues all T * R * S * (C / 16) MMAs of the item
// against shifted views of that run (the shared-memory descriptor's start address moves by whole pixel rows:
// t * 128 + r * Wp + s), and the weights of all taps stay resident in shared memory for the lifetime of the CTA.
// With T = 3 a 3x3 layer loads 232 rows per 128 outputs instead of the 390 of one run per (tile, filter row): the
// rows of the three filter rows overlap, and so do the runs of neighbouring tiles (measured: these 32-channel layers
// were bound by the ~2 cycles the TMA unit needs per row, and Conv2d_2a by HBM at 43 % of its peak).  Outputs that fall outside the valid region (x >= Wo or y >= Ho: the
// wrap-around columns / rows of the padded grid) are written as zeros, so the output tensor is itself a correctly
// zero-padded input grid for the next flat layer (written at a constant flat offset `out_shift`).
//
// Numerics are those of conv_umma_kernel: fp16 activations x (hi | lo) fp16 weights as one N = 2 * cout MMA,
// fp32 TMEM accumulation, hi and lo halves summed in the epilogue, folded-BN scale/shift + ReLU in fp32.
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "common.cuh"
#include "umma_ptx.cuh"

namespace tise {

// The epilogue of these layers is a latency chain per warp (~250 dependent instructions per 128 x 32 box: TMEM load, scale /
// shift, pack, staging, stores) and with two groups it set the pace (ablation: a third of the MMAs and no global stores still
// left 70 % of the time).  Four groups of 4 warps, one per accumulator stage of 128 TMEM columns, double its throughput.
constexpr int kFlatGroups = 4;
constexpr int kFlatAccCols = 512 / kFlatGroups;  // TMEM columns per accumulator stage
constexpr int kFlatThreads = 64 + 128 * kFlatGroups;  // producer warp, MMA warp, kFlatGroups epilogue groups of 4 warps
constexpr int kFlatMaxStages = 8;
constexpr int kFlatEpiBuf = 128 * 64;

struct FlatParams {
  const float* scale;
  const float* shift;
  __half* out;
  long long total;      // virtual output positions: n * Hp * Wp
  long long out_total;  // positions of the output tensor (stores beyond are dropped)
  int out_stride;       // elements per output pixel (= cout)
  int wp, hpwp;         // grid pitch and positions per image
  int ho, wo;           // valid outputs: y < ho && x < wo
  int out_shift;        // output position = v + out_shift
  int R, S, ksteps;     // filter taps, K = 16 steps per tap (cin / 16)
  int row_bytes;        // cin * 2: 64 (SWIZZLE_64B) or 128 (SWIZZLE_128B)
  int swz_type, sbo;
  int nt;               // cout (<= 128)
  int passes;           // 2: hi | lo planes side by side, MMA N = 2 * nt; 1: W_hi only (calibrated shift), MMA N = nt
  int tsub;                 // 128-pixel tiles per work item (their accumulators sit side by side in one TMEM stage)
  int n_boxes, box_rows;    // the run of a work item arrives as n_boxes TMA boxes of box_rows pixels
  int wtile_bytes;          // one tap's weights: 2 * nt rows
  int stages, stage_bytes, num_tiles, relu;
  int pdl;  // launched with programmatic stream serialization
  int debug;  // timing ablations (TISE_FLAT_DEBUG; results are garbage): 1 only the first filter row's MMAs, 2 no global stores
};

// TR, TS, TK: compile-time filter size and K steps (0: taken from the parameters at run time).  The MMA warp is a single
// instruction stream at ~10 cycles per dependent instruction, so for the 3x3x32-channel stem layers its loop is fully
// unrolled with all descriptor offsets folded into constants: one add + one tcgen05.mma per K step.
template <int TR, int TS, int TK>
__global__ void __launch_bounds__(kFlatThreads, 1)
conv_flat_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w, const FlatParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int taps = p.R * p.S;
  // layout: [weights taps x wtile][stages x run of the item][2 staged output boxes][scale | shift][barriers][tmem slot]
  const uint32_t w_base = smem_base;
  const uint32_t st_base = (w_base + (uint32_t)(taps * p.wtile_bytes) + 1023u) & ~1023u;
  const uint32_t epi_base = st_base + (uint32_t)p.stages * (uint32_t)p.stage_bytes;
  const uint32_t ss_base = epi_base + 2u * kFlatGroups * kFlatEpiBuf;
  const uint32_t bar_base = ss_base + 2u * 128u * 4u;
  auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(kFlatMaxStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (uint32_t)(2 * kFlatMaxStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(2 * kFlatMaxStages + kFlatGroups + s); };
  const uint32_t wfull_bar = bar_base + 8u * (2 * kFlatMaxStages + 2 * kFlatGroups);
  const uint32_t tmem_slot = wfull_bar + 8u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < kFlatGroups; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);
    }
    mbar_init(wfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (p.pdl) {
    pdl_wait();
    pdl_trigger();
  }

  if (warp == 0) {
    // ===================================================================== TMA producer
    const int n_stages = p.stages, n_boxes = p.n_boxes, box_rows = p.box_rows;
    const uint32_t stage_bytes = (uint32_t)p.stage_bytes, box_bytes = (uint32_t)(p.box_rows * p.row_bytes);
    const uint32_t tx = (uint32_t)p.n_boxes * box_bytes;
    const long long item_pos = (long long)p.tsub * kBlockM;
    if (elect_one()) {  // all weights, once
      mbar_expect_tx(wfull_bar, (uint32_t)(taps * p.wtile_bytes));
      for (int t = 0; t < taps; ++t) tma_load_2d(&tm_w, w_base + (uint32_t)(t * p.wtile_bytes), wfull_bar, t * (p.ksteps * 16), 0);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < p.num_tiles; item += gridDim.x) {
      const long long v0 = (long long)item * item_pos;
      mbar_wait(empty_bar(stage), phase ^ 1u);
      const uint32_t sa = st_base + (uint32_t)stage * stage_bytes;
      const uint32_t fb = full_bar(stage);
      if (elect_one()) {
        mbar_expect_tx(fb, tx);  // boxes that reach past the end of the tensor are zero-filled and still count in full
        for (int b = 0; b < n_boxes; ++b) tma_load_2d(&tm_a, sa + (uint32_t)b * box_bytes, fb, 0, (int)(v0 + (long long)b * box_rows));
      }
      if (++stage == n_stages) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    const uint32_t idesc = make_instr_desc(p.passes * p.nt);
    const uint64_t dhi = smem_desc_hi(p.swz_type, p.sbo);
    const int R = TR ? TR : p.R, S = TS ? TS : p.S, ksteps = TK ? TK : p.ksteps, n_stages = p.stages, tsub = p.tsub;
    const uint32_t stage_bytes = (uint32_t)p.stage_bytes, row_bytes = (uint32_t)p.row_bytes, wtile_bytes = (uint32_t)p.wtile_bytes;
    // Descriptor start-address fields (>> 4).  The swizzle of both TMA and tcgen05 is a function of the absolute
    // shared-memory address bits, so a view shifted by whole pixel rows needs nothing but the new start address (the
    // descriptor's base-offset field stays 0; setting it to (addr >> 7) & 7 gives wrong results -- measured).
    const uint64_t wdesc0 = make_smem_desc(w_base, dhi);
    const uint32_t tile_step = ((uint32_t)kBlockM * row_bytes) >> 4;  // next 128-pixel tile of the item
    const uint32_t frow_step = ((uint32_t)p.wp * row_bytes) >> 4;     // next filter row: Wp pixels further
    const uint32_t pix_step = row_bytes >> 4;                        // next filter column: one pixel further
    const uint32_t acc_cols = (uint32_t)(p.passes * p.nt);
    mbar_wait(wfull_bar, 0);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int item = blockIdx.x; item < p.num_tiles; item += gridDim.x) {
      mbar_wait(tempty_bar(as), aphase ^ 1u);
      mbar_wait(full_bar(stage), phase);
      tc_fence_after();
      const uint64_t adesc0 = make_smem_desc(st_base + (uint32_t)stage * stage_bytes, dhi);
      if (elect_one()) {
        for (int t = 0; t < tsub; ++t) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(as * kFlatAccCols) + (uint32_t)t * acc_cols;
          const uint64_t at = adesc0 + (uint64_t)((uint32_t)t * tile_step);
          if (TR && TS && TK) {
#pragma unroll
            for (int r = 0; r < ((p.debug & 1) ? 1 : TR); ++r)
#pragma unroll
              for (int s2 = 0; s2 < TS; ++s2)
#pragma unroll
                for (int k = 0; k < TK; ++k) {
                  const uint64_t ad = at + (uint64_t)((uint32_t)r * frow_step + (uint32_t)s2 * pix_step) + (uint64_t)(2 * k);
                  const uint64_t bd = wdesc0 + (uint64_t)(((uint32_t)(r * TS + s2) * wtile_bytes) >> 4) + (uint64_t)(2 * k);
                  umma_f16(d_tmem, ad, bd, idesc, (r | s2 | k) ? 1u : 0u);
                }
          } else {
            uint32_t first = 0;
            uint64_t bd = wdesc0;
            for (int r = 0; r < R; ++r)
              for (int s2 = 0; s2 < S; ++s2) {
                const uint64_t ad = at + (uint64_t)((uint32_t)r * frow_step + (uint32_t)s2 * pix_step);
                for (int k = 0; k < ksteps; ++k) {
                  umma_f16(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, first);
                  first = 1u;
                }
                bd += (uint64_t)(wtile_bytes >> 4);
              }
          }
        }
        umma_commit(empty_bar(stage));
        umma_commit(tfull_bar(as));
      }
      __syncwarp();
      if (++stage == n_stages) {
        stage = 0;
        phase ^= 1u;
      }
      if (++as == kFlatGroups) {
        as = 0;
        aphase ^= 1u;
      }
    }
  } else {
    // ===================================================================== epilogue (warps 2..5 and 6..9)
    // kFlatGroups groups, one per accumulator stage: group g drains this CTA's items number g, g + kFlatGroups, ...
    const int grp = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int etid = threadIdx.x - 64 - grp * 128;
    float* s_scale = reinterpret_cast<float*>(smem_raw + (ss_base - smem_u32(smem_raw)));
    float* s_shift = s_scale + 128;
    for (int c = etid; c < p.nt; c += 128) {
      s_scale[c] = __ldg(p.scale + c);
      s_shift[c] = __ldg(p.shift + c);
    }
    epi_bar(1 + grp);
    const uint32_t my_epi = epi_base + (uint32_t)grp * 2u * kFlatEpiBuf;
    const uint32_t wr_off = (uint32_t)row * 64u;
    const uint32_t wr_swz = (uint32_t)((row >> 1) & 3);
    const int st_chunk = etid & 3, st_row0 = etid >> 2;
    const int nt = p.nt, wp = p.wp, hpwp = p.hpwp, ho = p.ho, wo = p.wo;
    const bool relu = p.relu != 0, two_pass = p.passes == 2;
    const int as = grp;
    uint32_t aphase = 0;
    uint32_t box_it = 0;
    const int tsub = p.tsub;
    const long long item_pos = (long long)tsub * kBlockM;
    for (int item = blockIdx.x + grp * gridDim.x; item < p.num_tiles; item += kFlatGroups * gridDim.x) {
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      for (int t = 0; t < tsub; ++t) {
        const long long v0 = (long long)item * item_pos + (long long)t * kBlockM;
        // is this thread's output position a real output or a wrap-around cell of the padded grid (written as 0)?
        // (32-bit arithmetic: a 64-bit modulo is a ~100-instruction dependent chain in front of every tile's epilogue;
        // the host guarantees total < 2^31)
        const unsigned v = (unsigned)v0 + (unsigned)row;
        const unsigned rem = v % (unsigned)hpwp;
        const int y = (int)(rem / (unsigned)wp), x = (int)(rem - (unsigned)y * (unsigned)wp);
        const bool valid = (long long)v < p.total && y < ho && x < wo;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * kFlatAccCols + t * p.passes * nt);
        for (int c0 = 0; c0 < nt; c0 += 32, ++box_it) {
          const uint32_t buf = my_epi + (box_it & 1u) * kFlatEpiBuf;
          float f[32];
          if (two_pass) {
            uint32_t a[32], b[32];
            tmem_ld32(taddr + (uint32_t)c0, a);
            tmem_ld32(taddr + (uint32_t)(nt + c0), b);  // W_lo products
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(a[j]) + __uint_as_float(b[j]);
          } else {
            uint32_t a[32];
            tmem_ld32(taddr + (uint32_t)c0, a);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(a[j]);
          }
          if (c0 + 32 >= nt && t == tsub - 1) {  // all accumulators of the item read: hand the stage back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(as));
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 sc = *reinterpret_cast<const float4*>(s_scale + c0 + 4 * q);
            const float4 sh = *reinterpret_cast<const float4*>(s_shift + c0 + 4 * q);
            float* yv = f + 4 * q;
            yv[0] = fmaf(yv[0], sc.x, sh.x);
            yv[1] = fmaf(yv[1], sc.y, sh.y);
            yv[2] = fmaf(yv[2], sc.z, sh.z);
            yv[3] = fmaf(yv[3], sc.w, sh.w);
            if (!valid) yv[0] = yv[1] = yv[2] = yv[3] = 0.f;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {  // ReLU and saturation are part of the conversion
            const uint32_t dst = buf + wr_off + (((uint32_t)j ^ wr_swz) << 4);
            if (relu)
              st_shared_v4(dst, pack_h2_sat<true>(f[8 * j], f[8 * j + 1]), pack_h2_sat<true>(f[8 * j + 2], f[8 * j + 3]),
                           pack_h2_sat<true>(f[8 * j + 4], f[8 * j + 5]), pack_h2_sat<true>(f[8 * j + 6], f[8 * j + 7]));
            else
              st_shared_v4(dst, pack_h2_sat<false>(f[8 * j], f[8 * j + 1]), pack_h2_sat<false>(f[8 * j + 2], f[8 * j + 3]),
                           pack_h2_sat<false>(f[8 * j + 4], f[8 * j + 5]), pack_h2_sat<false>(f[8 * j + 6], f[8 * j + 7]));
          }
          epi_bar(1 + grp);
          const int ccol = c0 + st_chunk * 8;
          if (ccol < nt) {
            char* obase = reinterpret_cast<char*>(p.out) + (size_t)ccol * 2;
            const size_t pitch = (size_t)p.out_stride * 2;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = st_row0 + 32 * i;
              const long long mo = v0 + r + p.out_shift;
              if (mo < p.out_total && v0 + r < p.total && !(p.debug & 2)) {
                const uint4 val = ld_shared_v4(buf + (uint32_t)r * 64u + (((uint32_t)st_chunk ^ (uint32_t)((r >> 1) & 3)) << 4));
                *reinterpret_cast<uint4*>(obase + (size_t)mo * pitch) = val;
              }
            }
          }
        }
      }
      aphase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ----------------------------------------------------------------------------- host side
static PFN_cuTensorMapEncodeTiled_v12000 g_flat_encode = nullptr;

bool conv_flat_supported(const ConvLayer& L) {  // (two passes of nt columns must fit one accumulator stage)
  return L.stride == 1 && (L.cin == 32 || L.cin == 64) && 2 * L.cout <= kFlatAccCols && L.cout % 32 == 0 && L.kc == L.cin &&
         L.nt == L.cout && L.S <= 7 && L.R <= 7;
}

// Work-item geometry: T tiles of 128 positions share one contiguous run of T * 128 + (R - 1) * Wp + S - 1 pixels, which
// arrives as n_boxes TMA boxes (<= 256 rows each, a multiple of 8 so that every box starts on a swizzle atom).  T is
// limited by the 256 TMEM columns of an accumulator stage (2 * nt per tile) and by shared memory (>= 3 pipeline stages
// next to the resident weights).
struct FlatGeom {
  int tsub, n_boxes, box_rows, stage_bytes, stages;
};
static FlatGeom flat_geometry(const ConvLayer& L, int wp, int smem_limit, int passes) {
  const int row_bytes = L.cin * 2;
  const int fixed = 1024 + (int)align_up((size_t)L.R * L.S * passes * L.nt * row_bytes, 1024) + 2 * kFlatGroups * kFlatEpiBuf + 2 * 128 * 4 +
                    8 * (2 * kFlatMaxStages + 2 * kFlatGroups + 2) + 16;
  static int t_env = -1;
  if (t_env < 0) t_env = getenv("TISE_FLAT_T") ? atoi(getenv("TISE_FLAT_T")) : 0;
  // Sharing one run among T tiles cuts the TMA rows per output, but these kernels are bound by the shared-memory reads
  // of the MMA operands (A: 4 KB per M = 128, K = 16 step whatever N is), not by the fills: measured, T > 1 gains 7 % at
  // N = 64 (two-pass cout = 32) and loses 7 % at N = 128 (larger work items, fewer stages) -- so only narrow tiles share.
  int tmax = (passes * L.nt <= 64) ? kFlatAccCols / (passes * L.nt) : 1;
  if (tmax > 4) tmax = 4;
  if (t_env > 0) tmax = t_env < kFlatAccCols / (passes * L.nt) ? t_env : kFlatAccCols / (passes * L.nt);
  if (tmax < 1) tmax = 1;
  FlatGeom g{1, 1, 0, 0, 0};
  for (int t = tmax; t >= 1; --t) {
    const int run_rows = t * kBlockM + (L.R - 1) * wp + L.S - 1;
    const int n_boxes = (run_rows + 255) / 256;
    const int box_rows = ((run_rows + n_boxes - 1) / n_boxes + 7) / 8 * 8;
    const int stage_bytes = (int)align_up((size_t)n_boxes * box_rows * row_bytes, 1024);
    int stages = (smem_limit - fixed) / stage_bytes;
    if (stages > kFlatMaxStages) stages = kFlatMaxStages;
    g = FlatGeom{t, n_boxes, box_rows, stage_bytes, stages};
    if (stages >= 3 || (t == 1 && stages >= 2)) break;
  }
  return g;
}

int build_flat_plan(tise_ctx* ctx, const ConvLayer& L, const FlatArgs& a, FlatPlan* plan) {
  if (!conv_flat_supported(L)) return fail(ctx, TISE_ERR_INVALID, "flat conv: unsupported layer");
  if (!g_flat_encode) {
    cudaDriverEntryPointQueryResult q;
    void* fn = nullptr;
    TISE_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) return fail(ctx, TISE_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    g_flat_encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  }
  const CUtensorMapSwizzle swz = L.kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  const long long total = (long long)a.n * a.hp * a.wp;
  for (int passes = 1; passes <= 2; ++passes) {  // one pair of maps per operand policy (the policy can change per call)
    const FlatGeom g = flat_geometry(L, a.wp, ctx->smem_optin > 0 ? ctx->smem_optin : 232448, passes);
    if (g.stages < 2) return fail(ctx, TISE_ERR_INVALID, "flat conv: does not fit shared memory");
    {
      cuuint64_t dims[2] = {(cuuint64_t)L.cin, (cuuint64_t)total};
      cuuint64_t strides[1] = {(cuuint64_t)L.cin * 2};
      cuuint32_t box[2] = {(cuuint32_t)L.kc, (cuuint32_t)g.box_rows};
      cuuint32_t estr[2] = {1, 1};
      CUresult r = g_flat_encode(&plan->tm_a[passes - 1], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)a.in, dims, strides, box,
                                 estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(ctx, TISE_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for flat input", (int)r);
    }
    {  // rows [0, nt) of the packed weights are W_hi, [nt, 2 nt) W_lo (one output tile: nt = cout)
      cuuint64_t dims[2] = {(cuuint64_t)L.ktot, (cuuint64_t)2 * L.cout};
      cuuint64_t strides[1] = {(cuuint64_t)L.ktot * 2};
      cuuint32_t box[2] = {(cuuint32_t)L.kc, (cuuint32_t)(passes * L.nt)};
      cuuint32_t estr[2] = {1, 1};
      CUresult r = g_flat_encode(&plan->tm_w[passes - 1], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)L.d_w, dims, strides, box,
                                 estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(ctx, TISE_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for flat weights", (int)r);
    }
  }
  plan->valid = true;
  return TISE_OK;
}

int launch_conv_flat(tise_ctx* ctx, const ConvLayer& L, const FlatArgs& a, const FlatPlan& plan, int passes, cudaStream_t st) {
  if (!plan.valid) return fail(ctx, TISE_ERR_STATE, "flat conv: plan not built");
  FlatParams p;
  p.scale = L.d_scale;
  p.shift = (passes == 1 && L.d_shift1) ? L.d_shift1 : L.d_shift;
  p.passes = passes;
  p.out = a.out;
  p.total = (long long)a.n * a.hp * a.wp;
  if (p.total + 4 * kBlockM >= (1ll << 31)) return fail(ctx, TISE_ERR_INVALID, "flat conv: batch too large (%lld positions)", p.total);
  p.out_total = p.total;
  p.out_stride = L.cout;
  p.wp = a.wp;
  p.hpwp = a.hp * a.wp;
  p.ho = a.ho;
  p.wo = a.wo;
  p.out_shift = a.out_shift;
  p.R = L.R;
  p.S = L.S;
  p.ksteps = L.cin / 16;
  p.row_bytes = L.cin * 2;
  p.swz_type = L.kc == 64 ? 2 : 4;
  p.sbo = 8 * p.row_bytes;
  p.nt = L.cout;
  p.wtile_bytes = passes * L.nt * p.row_bytes;
  p.relu = a.relu;
  static int dbg_env = -1;
  if (dbg_env < 0) dbg_env = getenv("TISE_FLAT_DEBUG") ? atoi(getenv("TISE_FLAT_DEBUG")) : 0;
  p.debug = dbg_env;

  static PerDeviceFlag configured;
  const int smem_limit = ctx->smem_optin > 0 ? ctx->smem_optin : 232448;
  if (!configured.is(ctx, smem_limit)) {
    TISE_CUDA(ctx, cudaFuncSetAttribute(conv_flat_kernel<3, 3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit));
    TISE_CUDA(ctx, cudaFuncSetAttribute(conv_flat_kernel<0, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit));
    configured.set(ctx, smem_limit);
  }
  const FlatGeom g = flat_geometry(L, a.wp, smem_limit, passes);  // the same geometry the plan's input box was built for
  if (g.stages < 2) return fail(ctx, TISE_ERR_INVALID, "flat conv: does not fit shared memory");
  p.tsub = g.tsub;
  p.n_boxes = g.n_boxes;
  p.box_rows = g.box_rows;
  p.stage_bytes = g.stage_bytes;
  p.stages = g.stages;
  p.num_tiles = cdiv(p.total, (long long)g.tsub * kBlockM);  // work items
  const int grid = p.num_tiles < ctx->num_sms ? p.num_tiles : ctx->num_sms;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kFlatThreads);
  cfg.dynamicSmemBytes = (size_t)smem_limit;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  p.pdl = umma_pdl_mode() && !ctx->profiling ? 1 : 0;
  if (p.pdl) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  cudaError_t le;
  if (L.R == 3 && L.S == 3 && p.ksteps == 2)
    le = cudaLaunchKernelEx(&cfg, conv_flat_kernel<3, 3, 2>, plan.tm_a[passes - 1], plan.tm_w[passes - 1], p);
  else
    le = cudaLaunchKernelEx(&cfg, conv_flat_kernel<0, 0, 0>, plan.tm_a[passes - 1], plan.tm_w[passes - 1], p);
  if (le != cudaSuccess) return fail(ctx, TISE_ERR_CUDA, "conv_flat_kernel launch failed: %s", cudaGetErrorString(le));
  TISE_LAUNCH_CHECK(ctx);
  return TISE_OK;
}

}  // namespace tise
