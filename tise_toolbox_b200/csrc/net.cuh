// Host-side description of torchvision's Inception3 trunk as the hot path uses it
// (image_realism/FID/inception.py:59-95): packed layers, workspace plan, forward.
#pragma once
#include <memory>

#include "common.cuh"

namespace tise {

enum StepKind { STEP_STEM = 0, STEP_CONV, STEP_MAXPOOL, STEP_AVGPOOL, STEP_GAP, STEP_AVGPOOL_BN, STEP_FLATCONV, STEP_ZERO };

struct Step {
  StepKind kind;
  const ConvLayer* layer = nullptr;
  ConvArgs conv;    // STEP_CONV
  UmmaPlan umma;    // STEP_CONV, built lazily
  // pools
  const void* in = nullptr;
  void* out = nullptr;
  int n = 0, h = 0, w = 0, c = 0, in_stride = 0, out_stride = 0, in_f32 = 0, out_f32 = 0;
  FlatArgs flat;    // STEP_FLATCONV
  FlatPlan flatplan;
  int in_rows = 0, in_pitch = 0;  // STEP_MAXPOOL: physical input grid if not dense
  size_t zero_bytes = 0;          // STEP_ZERO: cudaMemsetAsync(out, 0, zero_bytes)
  const float* pool_scale = nullptr;  // STEP_AVGPOOL_BN
  const float* pool_shift = nullptr;
};

struct Plan {
  void* ws = nullptr;
  int n = 0;
  int tap = 3;
  int impl = 0;
  std::vector<Step> steps;
  float* feat_dst = nullptr;  // patched per call (GAP output)
};

struct PoolBN {
  float* d_scale = nullptr;
  float* d_shift = nullptr;
};

struct Net {
  std::map<std::string, std::vector<float>> raw;  // host copy of the state_dict (fp32)
  std::map<std::string, ConvLayer> convs;         // by torchvision module prefix (or "<block>.fused1x1")
  std::map<std::string, PoolBN> pool_bn;          // BN of the branch_pool convs, applied after the pooling
  float* d_stem_w = nullptr;                      // [27][32]
  float* d_stem_scale = nullptr;
  float* d_stem_shift = nullptr;
  __half* d_stem_w16 = nullptr;                   // [64][64] tensor-core operand of the stem (stem_umma.cu)
  float* d_stem_scale16 = nullptr;                // scale with the weight pre-scaling folded out
  float* d_fc_w = nullptr;  // [num_classes][2048]
  float* d_fc_b = nullptr;
  int num_classes = 0;
  size_t bytes_per_image[4] = {0, 0, 0, 0};  // workspace bytes per image for each tap
  std::vector<std::unique_ptr<Plan>> plans;
  std::vector<void*> device_allocs;
};

int net_load(tise_ctx* ctx, const tise_weight_desc* tensors, int n_tensors, int num_classes);
void net_free(tise_ctx* ctx);
int net_calibrate_lo(tise_ctx* ctx, int n, void* d_ws, size_t ws_bytes, cudaStream_t st);
int net_forward(tise_ctx* ctx, const uint8_t* d_u8, const float* d_f32, int n, int tap, float* d_feat, void* d_ws,
                size_t ws_bytes, cudaStream_t st);
int pack_conv_layer(tise_ctx* ctx, ConvLayer* L, const float* w /*[cout][cin][R][S]*/, const float* scale,
                    const float* shift, std::vector<void*>* allocs, int force_kc = 0);

}  // namespace tise
