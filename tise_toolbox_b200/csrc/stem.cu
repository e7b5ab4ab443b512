// Conv2d_1a_3x3 (3 -> 32, 3x3, stride 2, no padding; 299 -> 149) fused with the input conversion
// (A3/A4: ToTensor + affine, see tise_set_input_lut) and with BN + ReLU.  K = 27 is too thin for the
// tensor cores (0.34 % of the trunk's MACs), so this is a direct fp32 convolution on the CUDA
// cores; it reads the resized uint8 image (or the caller's float CHW batch) and writes the fp16
// NHWC tensor the tcgen05 layers consume.
#include "common.cuh"

namespace tise {

constexpr int kStemOut = 149;
constexpr int kStemC = 32;

template <bool kU8>
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

  const long long total = (long long)n * kStemOut * kStemOut;
  for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < total;
       m += (long long)gridDim.x * blockDim.x) {
    const int img = (int)(m / (kStemOut * kStemOut));
    const int rem = (int)(m - (long long)img * kStemOut * kStemOut);
    const int oy = rem / kStemOut, ox = rem - oy * kStemOut;

    float x[27];  // k = (r * 3 + s) * 3 + c
    if (kU8) {
      const uint8_t* base = u8 + ((size_t)img * kOutSize + 2 * oy) * kOutSize * 3 + 2 * ox * 3;
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int j = 0; j < 9; ++j) {  // j = s * 3 + c, 9 contiguous bytes of the row
          const int v = __ldg(base + (size_t)r * kOutSize * 3 + j);
          x[r * 9 + j] = s_lut[(j % 3) * 256 + v];
        }
    } else {
      const float* base = f32 + (size_t)img * 3 * kOutSize * kOutSize;
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s)
#pragma unroll
          for (int c = 0; c < 3; ++c)
            x[(r * 3 + s) * 3 + c] = __ldg(base + ((size_t)c * kOutSize + 2 * oy + r) * kOutSize + 2 * ox + s);
    }

    float acc[kStemC];
#pragma unroll
    for (int co = 0; co < kStemC; ++co) acc[co] = 0.f;
#pragma unroll
    for (int k = 0; k < 27; ++k) {
      const float4* wr = reinterpret_cast<const float4*>(s_w + k * kStemC);
#pragma unroll
      for (int q = 0; q < kStemC / 4; ++q) {
        const float4 wv = wr[q];
        acc[4 * q + 0] = fmaf(x[k], wv.x, acc[4 * q + 0]);
        acc[4 * q + 1] = fmaf(x[k], wv.y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(x[k], wv.z, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(x[k], wv.w, acc[4 * q + 3]);
      }
    }
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)m * kStemC);
#pragma unroll
    for (int q = 0; q < kStemC / 8; ++q) {
      __half2 h[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int co = 8 * q + 2 * j;
        float a = fmaxf(fmaf(acc[co], s_scale[co], s_shift[co]), 0.f);
        float b = fmaxf(fmaf(acc[co + 1], s_scale[co + 1], s_shift[co + 1]), 0.f);
        h[j] = __floats2half2_rn(fminf(a, 65504.f), fminf(b, 65504.f));
      }
      dst[q] = *reinterpret_cast<uint4*>(h);
    }
  }
}

int launch_stem(tise_ctx* ctx, const uint8_t* d_u8, const float* d_f32, int n, const float* d_w,
                const float* d_scale, const float* d_shift, __half* d_out, cudaStream_t st) {
  const long long total = (long long)n * kStemOut * kStemOut;
  const int blocks = (int)((total + 127) / 128 < (long long)ctx->num_sms * 16 ? (total + 127) / 128
                                                                              : (long long)ctx->num_sms * 16);
  if (d_u8)
    stem_kernel<true><<<blocks, 128, 0, st>>>(d_u8, nullptr, ctx->d_lut, d_w, d_scale, d_shift, d_out, n);
  else
    stem_kernel<false><<<blocks, 128, 0, st>>>(nullptr, d_f32, nullptr, d_w, d_scale, d_shift, d_out, n);
  TISE_LAUNCH_CHECK(ctx);
  return TISE_OK;
}

}  // namespace tise
