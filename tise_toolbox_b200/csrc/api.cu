// C ABI glue of libtise_b200.so (see include/tise_b200.h).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "net.cuh"

using namespace tise;

#define TISE_VERSION_INT 100  // 0.1.0

extern "C" int tise_version(void) { return TISE_VERSION_INT; }

extern "C" const char* tise_status_string(int status) {
  switch (status) {
    case TISE_OK: return "ok";
    case TISE_ERR_INVALID: return "invalid argument";
    case TISE_ERR_CUDA: return "CUDA error";
    case TISE_ERR_WORKSPACE: return "workspace too small";
    case TISE_ERR_STATE: return "invalid call order";
    case TISE_ERR_NUMERIC: return "non-finite intermediate";
    default: return "unknown status";
  }
}

extern "C" int tise_ctx_create(int device, tise_ctx** out) {
  if (!out) return TISE_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return TISE_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return TISE_ERR_CUDA;
  if (prop.major != 10) return TISE_ERR_CUDA;  // sm_100a only: no other code path exists in this library
  DeviceRestore restore;  // the caller's current device is left as it was
  if (cudaSetDevice(device) != cudaSuccess) return TISE_ERR_CUDA;
  tise_ctx* ctx = new tise_ctx();
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  ctx->smem_optin = (int)prop.sharedMemPerBlockOptin;
  if (cudaMalloc((void**)&ctx->d_lut, 768 * sizeof(float)) != cudaSuccess) {
    delete ctx;
    return TISE_ERR_CUDA;
  }
  *out = ctx;
  return TISE_OK;
}

int tise::stage_upload(tise_ctx* ctx, void* d_dst, const void* h_src, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return TISE_OK;
  tise_ctx::StageSlot& sl = ctx->stage[ctx->stage_next];
  ctx->stage_next = (ctx->stage_next + 1) % tise_ctx::kStageSlots;
  if (sl.pending) {  // only blocks when the host is kStageSlots uploads ahead of the GPU
    TISE_CUDA(ctx, cudaEventSynchronize(sl.done));
    sl.pending = false;
  }
  if (sl.bytes < bytes) {
    if (sl.host) TISE_CUDA(ctx, cudaFreeHost(sl.host));
    sl.host = nullptr;
    sl.bytes = 0;
    const size_t want = bytes < 65536 ? 65536 : bytes + bytes / 4;
    TISE_CUDA(ctx, cudaMallocHost(&sl.host, want));
    sl.bytes = want;
  }
  if (!sl.done) TISE_CUDA(ctx, cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
  memcpy(sl.host, h_src, bytes);
  TISE_CUDA(ctx, cudaMemcpyAsync(d_dst, sl.host, bytes, cudaMemcpyHostToDevice, st));
  TISE_CUDA(ctx, cudaEventRecord(sl.done, st));
  sl.pending = true;
  return TISE_OK;
}

extern "C" void tise_ctx_destroy(tise_ctx* ctx) {
  if (!ctx) return;
  DeviceRestore restore;
  cudaSetDevice(ctx->device);
  for (auto& sl : ctx->stage) {
    if (sl.pending) cudaEventSynchronize(sl.done);
    if (sl.done) cudaEventDestroy(sl.done);
    if (sl.host) cudaFreeHost(sl.host);
  }
  net_free(ctx);
  if (ctx->d_lut) cudaFree(ctx->d_lut);
  delete ctx;
}

extern "C" const char* tise_last_error(const tise_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" long long tise_launch_count(const tise_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int tise_set_conv_mode(tise_ctx* ctx, int impl, int passes) {
  if (!ctx) return TISE_ERR_INVALID;
  if ((impl != TISE_CONV_UMMA && impl != TISE_CONV_SIMT) || (passes != 1 && passes != 2))
    return fail(ctx, TISE_ERR_INVALID, "tise_set_conv_mode: impl %d passes %d", impl, passes);
  ctx->conv_impl = impl;
  ctx->passes = passes;
  return TISE_OK;
}

extern "C" int tise_set_pass_policy(tise_ctx* ctx, int single_pass_max_hw) {
  if (!ctx) return TISE_ERR_INVALID;
  if (single_pass_max_hw < 0) return fail(ctx, TISE_ERR_INVALID, "tise_set_pass_policy: %d", single_pass_max_hw);
  ctx->pass1_max_hw = single_pass_max_hw;
  return TISE_OK;
}

extern "C" int tise_set_progress_event(tise_ctx* ctx, void* cuda_event, int max_hw) {
  if (!ctx) return TISE_ERR_INVALID;
  if (cuda_event && max_hw <= 0) return fail(ctx, TISE_ERR_INVALID, "tise_set_progress_event: max_hw = %d", max_hw);
  ctx->progress_event = (cudaEvent_t)cuda_event;
  ctx->progress_max_hw = cuda_event ? max_hw : 0;
  return TISE_OK;
}

extern "C" int tise_set_input_lut(tise_ctx* ctx, const float* h_lut768) {
  if (!ctx) return TISE_ERR_INVALID;
  if (!h_lut768) return fail(ctx, TISE_ERR_INVALID, "tise_set_input_lut: null table");
  DeviceRestore restore;  // the caller's current device is left as it was
  TISE_CUDA(ctx, cudaSetDevice(ctx->device));
  TISE_CUDA(ctx, cudaMemcpy(ctx->d_lut, h_lut768, 768 * sizeof(float), cudaMemcpyHostToDevice));
  ctx->lut_set = true;
  return TISE_OK;
}

extern "C" int tise_load_inception(tise_ctx* ctx, const tise_weight_desc* tensors, int n_tensors, int num_classes) {
  if (!ctx) return TISE_ERR_INVALID;
  if (!tensors || n_tensors <= 0 || num_classes < 0)
    return fail(ctx, TISE_ERR_INVALID, "tise_load_inception: no tensors");
  DeviceRestore restore;  // the caller's current device is left as it was
  TISE_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = net_load(ctx, tensors, n_tensors, num_classes);
  if (rc) {
    std::string keep = ctx->err;
    net_free(ctx);
    ctx->err = keep;
  }
  return rc;
}

extern "C" int tise_calibrate_single_pass(tise_ctx* ctx, int n_images, void* d_ws, size_t ws_bytes, tise_stream stream) {
  if (!ctx) return TISE_ERR_INVALID;
  if (!d_ws) return fail(ctx, TISE_ERR_INVALID, "tise_calibrate_single_pass: null workspace");
  DeviceRestore restore;  // the caller's current device is left as it was
  TISE_CUDA(ctx, cudaSetDevice(ctx->device));
  return net_calibrate_lo(ctx, n_images, d_ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" size_t tise_inception_workspace_bytes(const tise_ctx* ctx, int n) {
  if (!ctx || !ctx->net || n <= 0) return 0;
  size_t m = 0;
  for (int t = 0; t < 4; ++t) m = ctx->net->bytes_per_image[t] > m ? ctx->net->bytes_per_image[t] : m;
  return m * (size_t)n;
}

extern "C" int tise_inception_features(tise_ctx* ctx, const uint8_t* d_u8, const float* d_f32, int n, int tap,
                                       float* d_feat, void* d_ws, size_t ws_bytes, tise_stream stream) {
  if (!ctx) return TISE_ERR_INVALID;
  return net_forward(ctx, d_u8, d_f32, n, tap, d_feat, d_ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int tise_debug_conv(tise_ctx* ctx, int impl, int passes, const void* d_x, int n, int h, int w, int cin,
                               const float* h_w, const float* h_scale, const float* h_shift, int cout, int kh, int kw,
                               int stride, int pad_h, int pad_w, int relu, void* d_y) {
  if (!ctx) return TISE_ERR_INVALID;
  if (!d_x || !h_w || !h_scale || !h_shift || !d_y) return fail(ctx, TISE_ERR_INVALID, "tise_debug_conv: null argument");
  if (passes != 1 && passes != 2) return fail(ctx, TISE_ERR_INVALID, "tise_debug_conv: passes %d", passes);
  DeviceRestore restore;  // the caller's current device is left as it was
  TISE_CUDA(ctx, cudaSetDevice(ctx->device));
  ConvLayer L;
  L.cin = cin;
  L.cout = cout;
  L.R = kh;
  L.S = kw;
  L.stride = stride;
  L.pad_h = pad_h;
  L.pad_w = pad_w;
  std::vector<void*> allocs;
  int rc = pack_conv_layer(ctx, &L, h_w, h_scale, h_shift, &allocs);
  if (!rc) {
    ConvArgs a;
    a.in = (const __half*)d_x;
    a.n = n;
    a.h = h;
    a.w = w;
    a.in_stride = cin;
    a.n_segs = 1;
    a.segs[0].out = d_y;
    a.segs[0].out_stride = cout;
    a.segs[0].c = cout;
    a.segs[0].relu = relu;
    a.out_f32 = 0;
    a.ho = conv_out_dim(h, kh, stride, pad_h);
    a.wo = conv_out_dim(w, kw, stride, pad_w);
    if (impl == TISE_CONV_UMMA) {
      UmmaPlan plan;
      rc = build_umma_plan(ctx, L, a, &plan);
      if (!rc) rc = launch_conv_umma(ctx, L, a, plan, passes, 0);
    } else {
      const int keep = ctx->passes;
      ctx->passes = passes;
      rc = launch_conv_simt(ctx, L, a, 0);
      ctx->passes = keep;
    }
    if (!rc) {
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) rc = fail(ctx, TISE_ERR_CUDA, "tise_debug_conv: %s", cudaGetErrorString(e));
    }
  }
  for (void* p : allocs) cudaFree(p);
  return rc;
}

extern "C" int tise_profile_enable(tise_ctx* ctx, int on) {
  if (!ctx) return TISE_ERR_INVALID;
  for (auto& r : ctx->prof) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  ctx->prof.clear();
  ctx->profiling = on != 0;
  return TISE_OK;
}

extern "C" int tise_profile_read(tise_ctx* ctx, double* ms_by_kind, long long* launches_by_kind, double* flops_by_kind) {
  if (!ctx || !ms_by_kind || !launches_by_kind || !flops_by_kind) return TISE_ERR_INVALID;
  for (int k = 0; k < 8; ++k) {
    ms_by_kind[k] = 0.0;
    launches_by_kind[k] = 0;
    flops_by_kind[k] = 0.0;
  }
  const char* dump = getenv("TISE_PROFILE_DUMP");  // optional per-launch listing: "<ms> <GFLOP> <tag>"
  FILE* df = dump ? fopen(dump, "w") : nullptr;
  for (auto& r : ctx->prof) {
    TISE_CUDA(ctx, cudaEventSynchronize(r.b));
    float ms = 0.f;
    TISE_CUDA(ctx, cudaEventElapsedTime(&ms, r.a, r.b));
    if (df) fprintf(df, "%.4f %.4f %s\n", ms, r.flops * 1e-9, r.tag);
    if (r.kind >= 0 && r.kind < 8) {
      ms_by_kind[r.kind] += ms;
      launches_by_kind[r.kind] += 1;
      flops_by_kind[r.kind] += r.flops;
    }
  }
  if (df) fclose(df);
  return TISE_OK;
}
