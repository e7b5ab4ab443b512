// Shared declarations of libtise_b200.so (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/tise_b200.h"

namespace tise {

constexpr int kOutSize = 299;       // network input side
constexpr int kNumSMsDefault = 148; // B200

struct Net;  // net.cu

}  // namespace tise

struct tise_ctx {
  int device = 0;
  int num_sms = tise::kNumSMsDefault;
  int smem_optin = 0;
  std::string err;
  long long launches = 0;
  int conv_impl = TISE_CONV_UMMA;
  int passes = 2;
  int pass1_max_hw = 0;  // layers whose input map is at most this many pixels high run one MMA pass (0: none)
  cudaEvent_t progress_event = nullptr;  // tise_set_progress_event: recorded once per forward pass ...
  int progress_max_hw = 0;               // ... in front of the first convolution whose input map is at most this high
  float* d_lut = nullptr;  // 768 floats
  bool lut_set = false;
  tise::Net* net = nullptr;
  // optional per-step timing of tise_inception_features (tise_profile_enable / tise_profile_read)
  bool profiling = false;
  struct ProfRec {
    int kind;
    cudaEvent_t a, b;
    double flops;
    char tag[96];
  };
  std::vector<ProfRec> prof;
  // Pinned staging ring for the small host -> device uploads of per-call descriptors.  A cudaMemcpyAsync from PAGEABLE
  // memory synchronises the stream before it copies, i.e. the host stalls behind everything queued so far and can
  // never run ahead of the GPU (measured: the H2D copy of the next batch then no longer overlaps the trunk).
  static constexpr int kStageSlots = 16;
  struct StageSlot {
    void* host = nullptr;
    size_t bytes = 0;
    cudaEvent_t done = nullptr;
    bool pending = false;
  };
  StageSlot stage[kStageSlots];
  int stage_next = 0;
};

namespace tise {

int fail(tise_ctx* ctx, int status, const char* fmt, ...);

#define TISE_CUDA(ctx, call)                                                                       \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return tise::fail(ctx, TISE_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,   \
                        cudaGetErrorString(e__));                                                  \
  } while (0)

#define TISE_LAUNCH_CHECK(ctx)                                                                     \
  do {                                                                                             \
    (ctx)->launches++;                                                                             \
    cudaError_t e__ = cudaGetLastError();                                                          \
    if (e__ != cudaSuccess)                                                                        \
      return tise::fail(ctx, TISE_ERR_CUDA, "kernel launch failed at %s:%d: %s", __FILE__,         \
                        __LINE__, cudaGetErrorString(e__));                                        \
  } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE property of a kernel: a process that drives several
// GPUs (one tise_ctx each) must opt in on every one of them.  Each launch site keeps one of these instead of a bool.
struct PerDeviceFlag {
  int value[64] = {0};
  bool is(const tise_ctx* ctx, int want) const { return ctx->device >= 0 && ctx->device < 64 && value[ctx->device] == want; }
  void set(const tise_ctx* ctx, int v) {
    if (ctx->device >= 0 && ctx->device < 64) value[ctx->device] = v;
  }
};

// Saves the caller's current device and restores it on scope exit (tise_ctx_create / tise_ctx_destroy select the
// context's device; the caller -- e.g. torch -- must not see its current device change behind its back).
struct DeviceRestore {
  int prev = -1;
  DeviceRestore() {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
  }
  ~DeviceRestore() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// host buffer -> device, truly asynchronous on `st` (see tise_ctx::stage): the bytes are copied into a pinned slot first
int stage_upload(tise_ctx* ctx, void* d_dst, const void* h_src, size_t bytes, cudaStream_t st);

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// One convolution layer, packed for the device (net.cu / conv_*.cu).
// K is walked as (filter tap, channel chunk of kc = 64 | 32 | 16 channels): kc is the largest of these that
// divides cin (so no zero padding is ever moved), except cin = 48 which stays one padded 64-chunk.
// Weights: fp16, K-major rows of ktot = taps * cin_pad elements with k = (r * S + s) * cin_pad + ci, split into
// hi = fp16(w * 2^e[co]) and lo = fp16(w * 2^e[co] - hi) (scale already contains 2^-e[co]).  Rows are grouped
// per output tile of nt channels as [tile][hi nt rows | lo nt rows], so the B operand of one tile is one
// contiguous block of 2 * nt rows: row(co, plane) = ((co / nt) * 2 + plane) * nt + co % nt.
// ---------------------------------------------------------------------------------------------
struct ConvLayer {
  int cin = 0, cout = 0, R = 1, S = 1, stride = 1, pad_h = 0, pad_w = 0;
  int kc = 64;              // channels per K chunk
  int cin_pad = 0;          // cin rounded up to kc
  int ktot = 0;             // R * S * cin_pad
  int nt = 0;               // output channels per tile (divides cout, multiple of 16, <= 256)
  __half* d_w = nullptr;    // [2 * cout][ktot], see above
  float* d_scale = nullptr;  // [cout]  gamma / sqrt(var + eps) * 2^-e
  float* d_shift = nullptr;  // [cout]  beta - mean * gamma / sqrt(var + eps)
  // [cout] shift of the SINGLE-PASS form of the layer: d_shift + d_scale * (W_lo . mean input activation).  The dropped
  // W_lo pass is a systematic (image-independent to first order) offset of the output; its mean-field value is added
  // back here (net.cu: net_calibrate_lo).  Equal to d_shift until calibrated.
  float* d_shift1 = nullptr;
  int fold = 1;              // filter taps folded into the channel axis (cin = fold * physical channels), see net.cu
};
int conv_chunk_of(int cin);
int conv_ntile_of(int cout);
bool umma_pdl_mode();   // conv_umma.cu: programmatic dependent launch unless TISE_PDL=0
bool umma_pair_mode();  // conv_umma.cu: CTA pairs (cta_group::2) unless TISE_UMMA_PAIR=0
static inline size_t conv_weight_row(const ConvLayer& L, int co, int plane) {
  return ((size_t)(co / L.nt) * 2 + plane) * L.nt + co % L.nt;
}

// A conv launch: NHWC fp16 input view -> up to kMaxSegs NHWC destinations.  The layer's output channels
// are split into consecutive segments, each written into its own tensor / channel slice (this is how
// sibling 1x1 convolutions that read the same input run as one GEMM, and how torch.cat costs nothing).
constexpr int kMaxSegs = 4;
struct ConvSeg {
  void* out = nullptr;  // base of the destination tensor with the channel offset already applied
  int out_stride = 0;   // elements per destination pixel
  int c = 0;            // channels in this segment
  int relu = 1;         // 0: raw scale/shift only (branch_pool convs, pooled + BN + ReLU afterwards)
};
struct ConvArgs {
  const __half* in = nullptr;  // base of the input tensor (channel 0 of the view)
  int n = 0, h = 0, w = 0;     // input spatial dims
  int in_stride = 0;           // elements per input pixel (>= cin)
  int in_row_pitch = 0;        // pixels per physical input row (0: w).  Row-folded layers (net.cu: conv_folded) read a
                               // window of S pixels as ONE pixel of S * cin channels, so their logical width is wo
  ConvSeg segs[kMaxSegs];
  int n_segs = 0;
  int out_f32 = 0;  // all segments fp32 (Mixed_7c) or all fp16
  int ho = 0, wo = 0;
};

int conv_out_dim(int in, int k, int stride, int pad);

// conv_simt.cu
int launch_conv_simt(tise_ctx* ctx, const ConvLayer& L, const ConvArgs& a, cudaStream_t st);
// conv_umma.cu
struct UmmaPlan {  // tensor maps are tied to pointers, so a plan is built once per (layer, buffers)
  alignas(64) CUtensorMap tm_a;    // im2col map of the input view, box kc channels x 128 pixels
  alignas(64) CUtensorMap tm_b;    // weights, box kc x nt rows
  alignas(64) CUtensorMap tm_b2;   // weights, box kc x 2 nt rows (hi + lo in one load; only if 2 nt <= 256)
  alignas(64) CUtensorMap tm_bs;   // weights, box kc x nt / 2 rows: the slices of the cluster multicast
  bool valid = false;
};
int build_umma_plan(tise_ctx* ctx, const ConvLayer& L, const ConvArgs& a, UmmaPlan* plan);
int launch_conv_umma(tise_ctx* ctx, const ConvLayer& L, const ConvArgs& a, const UmmaPlan& plan, int passes,
                     cudaStream_t st);

// conv_flat.cu: stride-1 convolutions with 32 / 64 input channels evaluated on the flattened (padded) input grid
struct FlatArgs {
  const __half* in = nullptr;  // [n][hp][wp][cin], the consumer's padding already materialised as zeros
  __half* out = nullptr;       // [n][hp][wp][cout] on the same grid
  int n = 0, hp = 0, wp = 0;   // physical grid
  int ho = 0, wo = 0;          // valid outputs (top-left aligned): y < ho && x < wo; the rest is written as 0
  int out_shift = 0;           // output position = input position + out_shift (e.g. wp + 1 to centre for a pad-1 consumer)
  int relu = 1;
};
struct FlatPlan {  // index 0: single-pass form (W_hi rows only), 1: two-pass form (hi | lo rows)
  alignas(64) CUtensorMap tm_a[2];
  alignas(64) CUtensorMap tm_w[2];
  bool valid = false;
};
bool conv_flat_supported(const ConvLayer& L);
int build_flat_plan(tise_ctx* ctx, const ConvLayer& L, const FlatArgs& a, FlatPlan* plan);
int launch_conv_flat(tise_ctx* ctx, const ConvLayer& L, const FlatArgs& a, const FlatPlan& plan, int passes, cudaStream_t st);

// stem.cu : Conv2d_1a_3x3 (3->32, 3x3, stride 2) fused with the uint8 LUT / float input
int launch_stem(tise_ctx* ctx, const uint8_t* d_u8, const float* d_f32, int n, const float* d_w /*[27][32]*/,
                const float* d_scale, const float* d_shift, __half* d_out, cudaStream_t st);

// stem_umma.cu : the same layer on the tensor cores (uint8 input only): thread-built A tiles, resident weights
void pack_stem_umma(const float* w_k_major /*[27][32]*/, float* scale /*[32], in/out*/, __half* out /*[64][64]*/);
int launch_stem_umma(tise_ctx* ctx, const uint8_t* d_u8, int n, const __half* d_w16, const float* d_scale,
                     const float* d_shift, __half* d_out, cudaStream_t st);

// pool.cu
// in_rows / in_pitch: physical rows per image and pixels per row of the input (>= h, w: flat-conv outputs keep the
// padded grid); 0 = dense
int launch_maxpool3s2(tise_ctx* ctx, const __half* in, int n, int h, int w, int c, int in_stride, __half* out,
                      int out_stride, cudaStream_t st, int in_rows = 0, int in_pitch = 0);
int launch_avgpool3s1p1(tise_ctx* ctx, const __half* in, int n, int h, int w, int c, int in_stride, __half* out,
                        int out_stride, cudaStream_t st);
// avg_pool2d(3, 1, 1) of a raw (pre-BN) fp16 tensor followed by scale/shift + ReLU into a channel slice
int launch_avgpool_bn(tise_ctx* ctx, const __half* in, int n, int h, int w, int c, const float* scale,
                      const float* shift, void* out, int out_stride, int out_f32, cudaStream_t st);
int launch_global_avgpool(tise_ctx* ctx, const void* in, int in_f32, int n, int hw, int c, int in_stride,
                          float* out, cudaStream_t st);

// resize.cu / stats.cu / frechet.cu entry points are the C ABI functions themselves.

}  // namespace tise
