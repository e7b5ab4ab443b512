// Conv2d_1a_3x3 (3 -> 32, 3x3, stride 2, no padding; 299 -> 149) fused with the input conversion
// (A3/A4: ToTensor + affine, see tise_set_input_lut) and with BN + ReLU.  K = 27 is too thin for the
// tensor cores (0.34 % of the trunk's MACs), so this is a direct fp32 convolution on the CUDA
// cores; it reads the resized uint8 image (or the caller's float CHW batch) and writes the fp16
// NHWC tensor the tcgen05 layers consume.
#include <stdlib.h>

#include "common.cuh"

namespace tise {

constexpr int kStemOut = 149;
constexpr int kStemC = 32;

// kCh = output channels per thread: 32, or 16 (two neighbouring threads share a pixel pair; fewer registers, more
// resident warps to hide the input-load latency).
template <bool kU8, int kCh>
__global__ void __launch_bounds__(128) stem_kernel(const uint8_t* __restrict__ u8, const float* __restrict__ f32,
                                                   const float* __restrict__ lut, const float* __restrict__ w,
                                                   const float* __restrict__ scale, const float* __restrict__ shift,
                                                   __half* __restrict__ out, int n) {
  __shared__ __align__(16) float s_w[27 * kStemC];
  __shared__ float s_lut[768];
  __shared__ float s_scale[kStemC], s_shift[kStemC];
  for (int i = threadIdx.x; i < 27 * kStemC; i += blockDim.x) s_w[i] = w[i];
  if (kU8)
    for (int i = threadIdx.x; i < 768; i += blockDim.x) s_lut[i] = lut[i];
  if (threadIdx.x < kStemC) {
    s_scale[threadIdx.x] = scale[threadIdx.x];
    s_shift[threadIdx.x] = shift[threadIdx.x];
  }
  __syncthreads();

  // Two horizontally adjacent output pixels per thread: every weight fetched from shared memory (one LDS.128 per
  // 4 output channels) feeds 8 FMAs instead of 4 -- with one pixel per thread the kernel was bound by the 243
  // shared-memory loads per 864 FMAs, not by the FMA pipe.
  constexpr int kPairs = (kStemOut + 1) / 2;  // 75 pixel pairs per output row (the last one is a single pixel)
  constexpr int kSplit = kStemC / kCh;
  const long long total = (long long)n * kStemOut * kPairs * kSplit;
  for (long long tt = (long long)blockIdx.x * blockDim.x + threadIdx.x; tt < total; tt += (long long)gridDim.x * blockDim.x) {
    const long long t = tt / kSplit;
    const int coff = (int)(tt - t * kSplit) * kCh;  // first output channel of this thread
    const int img = (int)(t / (kStemOut * kPairs));
    const int rem = (int)(t - (long long)img * kStemOut * kPairs);
    const int oy = rem / kPairs, ox = (rem - oy * kPairs) * 2;
    const bool two = ox + 1 < kStemOut;

    // input window: 3 rows x 5 columns x 3 channels (columns 2*ox .. 2*ox+4); x[r][col*3 + c]
    float x[3][15];
    if (kU8) {
      const uint8_t* base = u8 + ((size_t)img * kOutSize + 2 * oy) * kOutSize * 3 + 2 * ox * 3;
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int j = 0; j < 15; ++j) {
          const bool ok = j < 9 || two;  // columns 3, 4 exist only if the second pixel does
          const int v = ok ? (int)__ldg(base + (size_t)r * kOutSize * 3 + j) : 0;
          x[r][j] = s_lut[(j % 3) * 256 + v];
        }
    } else {
      const float* base = f32 + (size_t)img * 3 * kOutSize * kOutSize;
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int col = 0; col < 5; ++col)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const bool ok = col < 3 || two;
            x[r][col * 3 + c] = ok ? __ldg(base + ((size_t)c * kOutSize + 2 * oy + r) * kOutSize + 2 * ox + col) : 0.f;
          }
    }

    float acc0[kCh], acc1[kCh];
#pragma unroll
    for (int co = 0; co < kCh; ++co) acc0[co] = acc1[co] = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int sc = 0; sc < 9; ++sc) {  // sc = s * 3 + c; weights k = (r * 3 + s) * 3 + c = r * 9 + sc
        const float4* wr = reinterpret_cast<const float4*>(s_w + (r * 9 + sc) * kStemC + coff);
        const float a = x[r][sc], b = x[r][sc + 6];  // second pixel: two columns (6 values) to the right
#pragma unroll
        for (int q = 0; q < kCh / 4; ++q) {
          const float4 wv = wr[q];
          acc0[4 * q + 0] = fmaf(a, wv.x, acc0[4 * q + 0]);
          acc0[4 * q + 1] = fmaf(a, wv.y, acc0[4 * q + 1]);
          acc0[4 * q + 2] = fmaf(a, wv.z, acc0[4 * q + 2]);
          acc0[4 * q + 3] = fmaf(a, wv.w, acc0[4 * q + 3]);
          acc1[4 * q + 0] = fmaf(b, wv.x, acc1[4 * q + 0]);
          acc1[4 * q + 1] = fmaf(b, wv.y, acc1[4 * q + 1]);
          acc1[4 * q + 2] = fmaf(b, wv.z, acc1[4 * q + 2]);
          acc1[4 * q + 3] = fmaf(b, wv.w, acc1[4 * q + 3]);
        }
      }
    const size_t m = ((size_t)img * kStemOut + oy) * kStemOut + ox;
#pragma unroll
    for (int px = 0; px < 2; ++px) {
      if (px == 1 && !two) break;
      const float* acc = px ? acc1 : acc0;
      uint4* dst = reinterpret_cast<uint4*>(out + (m + px) * kStemC + coff);
#pragma unroll
      for (int q = 0; q < kCh / 8; ++q) {
        __half2 h[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int co = 8 * q + 2 * j;
          const float ya = fmaxf(fmaf(acc[co], s_scale[coff + co], s_shift[coff + co]), 0.f);
          const float yb = fmaxf(fmaf(acc[co + 1], s_scale[coff + co + 1], s_shift[coff + co + 1]), 0.f);
          h[j] = __floats2half2_rn(fminf(ya, 65504.f), fminf(yb, 65504.f));
        }
        dst[q] = *reinterpret_cast<uint4*>(h);
      }
    }
  }
}

int launch_stem(tise_ctx* ctx, const uint8_t* d_u8, const float* d_f32, int n, const float* d_w,
                const float* d_scale, const float* d_shift, __half* d_out, cudaStream_t st) {
  static int split = -1;
  if (split < 0) split = getenv("TISE_STEM_SPLIT") ? atoi(getenv("TISE_STEM_SPLIT")) : 1;
  const long long total = (long long)n * kStemOut * ((kStemOut + 1) / 2) * (split ? 2 : 1);
  const long long want = (total + 127) / 128, cap = (long long)ctx->num_sms * 32;
  const int blocks = (int)(want < cap ? want : cap);
  if (d_u8) {
    if (split) stem_kernel<true, 16><<<blocks, 128, 0, st>>>(d_u8, nullptr, ctx->d_lut, d_w, d_scale, d_shift, d_out, n);
    else stem_kernel<true, 32><<<blocks, 128, 0, st>>>(d_u8, nullptr, ctx->d_lut, d_w, d_scale, d_shift, d_out, n);
  } else {
    if (split) stem_kernel<false, 16><<<blocks, 128, 0, st>>>(nullptr, d_f32, nullptr, d_w, d_scale, d_shift, d_out, n);
    else stem_kernel<false, 32><<<blocks, 128, 0, st>>>(nullptr, d_f32, nullptr, d_w, d_scale, d_shift, d_out, n);
  }
  TISE_LAUNCH_CHECK(ctx);
  return TISE_OK;
}

}  // namespace tise
