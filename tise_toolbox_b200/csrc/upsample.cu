// Float bilinear up/down-sampling to 299 x 299 for callers that hand the reference-facing modules float tensors of
// another size: InceptionV3.forward's F.upsample(size=(299, 299), mode="bilinear",
// align_corners=True) (image_realism/FID/inception.py:117-118, an identity on the Dataset's 299 x 299 output) and the O-IS
// script's nn.Upsample(size=(299, 299), mode="bilinear") (object_centric_inception_score.py:49, align_corners=False).
// Same source-index arithmetic as ATen's upsample_bilinear2d (area_pixel_compute_scale / compute_source_index, float).
#include "common.cuh"

namespace tise {

__global__ void __launch_bounds__(256) upsample_bilinear_kernel(const float* __restrict__ in, int planes, int h, int w,
                                                                int align_corners, float* __restrict__ out) {
  const long long total = (long long)planes * kOutSize * kOutSize;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ox = (int)(idx % kOutSize), oy = (int)((idx / kOutSize) % kOutSize);
  const long long plane = idx / ((long long)kOutSize * kOutSize);
  float sy, sx;
  if (align_corners) {
    const float ry = kOutSize > 1 ? (float)(h - 1) / (float)(kOutSize - 1) : 0.f;
    const float rx = kOutSize > 1 ? (float)(w - 1) / (float)(kOutSize - 1) : 0.f;
    sy = __fmul_rn(ry, (float)oy);
    sx = __fmul_rn(rx, (float)ox);
  } else {
    const float ry = (float)h / (float)kOutSize, rx = (float)w / (float)kOutSize;
    sy = __fsub_rn(__fmul_rn(ry, __fadd_rn((float)oy, 0.5f)), 0.5f);
    sx = __fsub_rn(__fmul_rn(rx, __fadd_rn((float)ox, 0.5f)), 0.5f);
    sy = sy < 0.f ? 0.f : sy;
    sx = sx < 0.f ? 0.f : sx;
  }
  int y0 = (int)sy, x0 = (int)sx;
  y0 = y0 > h - 1 ? h - 1 : y0;
  x0 = x0 > w - 1 ? w - 1 : x0;
  const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
  const float ly1 = __fsub_rn(sy, (float)y0), lx1 = __fsub_rn(sx, (float)x0);
  const float ly0 = __fsub_rn(1.f, ly1), lx0 = __fsub_rn(1.f, lx1);
  const float* p = in + plane * (long long)h * w;
  const float v00 = p[(long long)y0 * w + x0], v01 = p[(long long)y0 * w + x1];
  const float v10 = p[(long long)y1 * w + x0], v11 = p[(long long)y1 * w + x1];
  const float top = __fadd_rn(__fmul_rn(lx0, v00), __fmul_rn(lx1, v01));
  const float bot = __fadd_rn(__fmul_rn(lx0, v10), __fmul_rn(lx1, v11));
  out[idx] = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
}

// y[n][c][i] = x[n][c][i] * scale[c] + bias[c] with a rounded multiply followed by a rounded add (what the eager
// `x[:, c] = x[:, c] * s + b` of image_realism/FID/inception.py:120-124 computes -- no FMA contraction)
struct Affine4 {
  float scale[4], bias[4];
};
__global__ void channel_affine_kernel(const float* __restrict__ in, long long total, int c, int hw, Affine4 a, float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)((i / hw) % c);
    out[i] = __fadd_rn(__fmul_rn(in[i], a.scale[ch]), a.bias[ch]);
  }
}

}  // namespace tise

using namespace tise;

extern "C" int tise_channel_affine_f32(tise_ctx* ctx, const float* d_in, int n, int c, int hw, const float* h_scale,
                                       const float* h_bias, float* d_out, tise_stream stream) {
  if (!ctx) return TISE_ERR_INVALID;
  if (n == 0) return TISE_OK;
  if (!d_in || !d_out || !h_scale || !h_bias || n < 0 || c <= 0 || c > 4 || hw <= 0)
    return fail(ctx, TISE_ERR_INVALID, "tise_channel_affine_f32: bad argument");
  Affine4 a;
  for (int k = 0; k < 4; ++k) {
    a.scale[k] = k < c ? h_scale[k] : 1.f;
    a.bias[k] = k < c ? h_bias[k] : 0.f;
  }
  const long long total = (long long)n * c * hw;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)ctx->num_sms * 32) blocks = (long long)ctx->num_sms * 32;
  channel_affine_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_in, total, c, hw, a, d_out);
  TISE_LAUNCH_CHECK(ctx);
  return TISE_OK;
}

extern "C" int tise_upsample_bilinear_f32(tise_ctx* ctx, const float* d_in, int n, int c, int h, int w, int align_corners,
                                          float* d_out, tise_stream stream) {
  if (!ctx) return TISE_ERR_INVALID;
  if (n == 0) return TISE_OK;
  if (!d_in || !d_out || n < 0 || c <= 0 || h <= 0 || w <= 0)
    return fail(ctx, TISE_ERR_INVALID, "tise_upsample_bilinear_f32: bad argument");
  const long long total = (long long)n * c * kOutSize * kOutSize;
  upsample_bilinear_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_in, n * c, h, w, align_corners,
                                                                                               d_out);
  TISE_LAUNCH_CHECK(ctx);
  return TISE_OK;
}
