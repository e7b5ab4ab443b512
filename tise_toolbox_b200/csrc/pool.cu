// Pooling layers of Inception-v3 on NHWC fp16 tensors (torchvision Inception3: max_pool2d(3, 2),
// avg_pool2d(3, 1, 1) with count_include_pad=True, adaptive_avg_pool2d(1)).  HBM/L2-bound
// elementwise work: one thread per (pixel, 8-channel group), 16-byte loads and stores.
#include "common.cuh"

namespace tise {

__global__ void maxpool3s2_kernel(const __half* __restrict__ in, int n, int h, int w, int c8, int in_stride,
                                  __half* __restrict__ out, int out_stride, int ho, int wo) {
  // h, w here are the PHYSICAL rows per image / pixels per row of the input
  const long long total = (long long)n * ho * wo * c8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % c8);
    long long p = i / c8;
    const int ox = (int)(p % wo);
    p /= wo;
    const int oy = (int)(p % ho);
    const int img = (int)(p / ho);
    const __half* base = in + (((size_t)img * h + 2 * oy) * w + 2 * ox) * in_stride + g * 8;
    __half2 m[4];
    {
      const uint4 v = *reinterpret_cast<const uint4*>(base);
      const __half2* hv = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) m[j] = hv[j];
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        if (r == 0 && s == 0) continue;
        const uint4 v = *reinterpret_cast<const uint4*>(base + ((size_t)r * w + s) * in_stride);
        const __half2* hv = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) m[j] = __hmax2(m[j], hv[j]);
      }
    *reinterpret_cast<uint4*>(out + (((size_t)img * ho + oy) * wo + ox) * out_stride + g * 8) =
        *reinterpret_cast<uint4*>(m);
  }
}

__global__ void avgpool3s1p1_kernel(const __half* __restrict__ in, int n, int h, int w, int c8, int in_stride,
                                    __half* __restrict__ out, int out_stride) {
  const long long total = (long long)n * h * w * c8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % c8);
    long long p = i / c8;
    const int ox = (int)(p % w);
    p /= w;
    const int oy = (int)(p % h);
    const int img = (int)(p / h);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int r = -1; r <= 1; ++r)
#pragma unroll
      for (int s = -1; s <= 1; ++s) {
        const int y = oy + r, x = ox + s;
        if (y < 0 || y >= h || x < 0 || x >= w) continue;  // zero padding, still divided by 9
        const uint4 v = *reinterpret_cast<const uint4*>(in + (((size_t)img * h + y) * w + x) * in_stride + g * 8);
        const __half2* hv = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(hv[j]);
          acc[2 * j] += f.x;
          acc[2 * j + 1] += f.y;
        }
      }
    __half2 o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = __floats2half2_rn(acc[2 * j] / 9.0f, acc[2 * j + 1] / 9.0f);
    *reinterpret_cast<uint4*>(out + (((size_t)img * h + oy) * w + ox) * out_stride + g * 8) =
        *reinterpret_cast<uint4*>(o);
  }
}

// branch_pool of InceptionA/C/E with the 1x1 convolution commuted in front of the pooling: the input is the
// RAW (pre-BN) fp16 output of that convolution (pf = 32..192 channels instead of the block's 192..2048), so
// avg_pool2d(3,1,1)(count_include_pad) + BN scale/shift + ReLU is applied here.  Exact: conv1x1 and the
// zero-padded average are both linear and the convolution has no bias.
// One thread owns kXs horizontally adjacent output columns x 8 channels of one image and walks DOWN the rows with a
// three-row window of the kXs + 2 input columns in registers: every input value is loaded once per column strip
// ((kXs + 2) / kXs = 1.5 loads per output instead of the 4.5 of the per-row version, which ran at 7 TB/s of L2 -> SM
// traffic, i.e. L2-bound, for 2.1-2.7 TB/s of HBM traffic -- profiles/launches_r02_summary.txt).  The loads of row y + 2
// are issued before row y is evaluated, so a thread never waits for the row it needs next.
// 32-bit index arithmetic (the host checks the range); consecutive threads take consecutive channel groups (coalesced).
template <bool kF32, int kXs>
__global__ void __launch_bounds__(128) avgpool_bn_kernel(const __half* __restrict__ in, int n, int h, int w, int c8, int segs,
                                                         const float* __restrict__ scale, const float* __restrict__ shift,
                                                         void* __restrict__ out, int out_stride) {
  const unsigned total = (unsigned)n * segs * c8;
  const int c = c8 * 8;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int g = (int)(i % (unsigned)c8);
    unsigned p = i / (unsigned)c8;
    const int sg = (int)(p % (unsigned)segs);
    const int img = (int)(p / (unsigned)segs);
    const int x0 = sg * kXs;
    const __half* base = in + (size_t)img * h * w * c + g * 8;
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = __ldg(scale + g * 8 + j) * (1.0f / 9.0f);
      sh[j] = __ldg(shift + g * 8 + j);
    }
    auto load_row = [&](int y, uint4 (&r)[kXs + 2]) {
#pragma unroll
      for (int t = 0; t < kXs + 2; ++t) {
        const int x = x0 - 1 + t;
        const bool ok = y >= 0 && y < h && x >= 0 && x < w;
        r[t] = ok ? __ldg(reinterpret_cast<const uint4*>(base + ((size_t)y * w + x) * c)) : make_uint4(0, 0, 0, 0);
      }
    };
    uint4 r0[kXs + 2], r1[kXs + 2], r2[kXs + 2], r3[kXs + 2];  // rows y - 1, y, y + 1 and the prefetched y + 2
    load_row(-1, r0);
    load_row(0, r1);
    load_row(1, r2);
    for (int y = 0; y < h; ++y) {
      load_row(y + 2, r3);
      // column sums of the three rows
      float cs[kXs + 2][8];
#pragma unroll
      for (int t = 0; t < kXs + 2; ++t)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 a = __half22float2(reinterpret_cast<const __half2*>(&r0[t])[j]);
          const float2 b = __half22float2(reinterpret_cast<const __half2*>(&r1[t])[j]);
          const float2 d = __half22float2(reinterpret_cast<const __half2*>(&r2[t])[j]);
          cs[t][2 * j] = a.x + b.x + d.x;
          cs[t][2 * j + 1] = a.y + b.y + d.y;
        }
#pragma unroll
      for (int t = 0; t < kXs; ++t) {
        const int x = x0 + t;
        if (x < w) {
          float y8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) y8[j] = fmaxf(fmaf(cs[t][j] + cs[t + 1][j] + cs[t + 2][j], sc[j], sh[j]), 0.f);
          const size_t o = (((size_t)img * h + y) * w + x) * out_stride + g * 8;
          if (kF32) {
            float4* dst = reinterpret_cast<float4*>((float*)out + o);
            dst[0] = make_float4(y8[0], y8[1], y8[2], y8[3]);
            dst[1] = make_float4(y8[4], y8[5], y8[6], y8[7]);
          } else {
            __half2 hh[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) hh[j] = __floats2half2_rn(fminf(y8[2 * j], 65504.f), fminf(y8[2 * j + 1], 65504.f));
            *reinterpret_cast<uint4*>((__half*)out + o) = *reinterpret_cast<uint4*>(hh);
          }
        }
      }
#pragma unroll
      for (int t = 0; t < kXs + 2; ++t) {
        r0[t] = r1[t];
        r1[t] = r2[t];
        r2[t] = r3[t];
      }
    }
  }
}

// out[img][ch] = mean over hw positions.  Consecutive threads -> consecutive channels (coalesced).
template <typename T>
__global__ void global_avgpool_kernel(const T* __restrict__ in, int n, int hw, int c, int in_stride,
                                      float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * c) return;
  const int img = i / c, ch = i - img * c;
  const T* p = in + (size_t)img * hw * in_stride + ch;
  float acc = 0.f;
  for (int k = 0; k < hw; ++k) acc += (float)p[(size_t)k * in_stride];
  out[i] = acc / (float)hw;
}

static int grid_for(long long total, int threads, int num_sms) {
  long long b = (total + threads - 1) / threads;
  const long long cap = (long long)num_sms * 32;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

int launch_maxpool3s2(tise_ctx* ctx, const __half* in, int n, int h, int w, int c, int in_stride, __half* out,
                      int out_stride, cudaStream_t st, int in_rows, int in_pitch) {
  const int ho = (h - 3) / 2 + 1, wo = (w - 3) / 2 + 1;
  const long long total = (long long)n * ho * wo * (c / 8);
  maxpool3s2_kernel<<<grid_for(total, 256, ctx->num_sms), 256, 0, st>>>(in, n, in_rows ? in_rows : h, in_pitch ? in_pitch : w,
                                                                       c / 8, in_stride, out, out_stride, ho, wo);
  TISE_LAUNCH_CHECK(ctx);
  return TISE_OK;
}

int launch_avgpool3s1p1(tise_ctx* ctx, const __half* in, int n, int h, int w, int c, int in_stride, __half* out,
                        int out_stride, cudaStream_t st) {
  const long long total = (long long)n * h * w * (c / 8);
  avgpool3s1p1_kernel<<<grid_for(total, 256, ctx->num_sms), 256, 0, st>>>(in, n, h, w, c / 8, in_stride, out,
                                                                         out_stride);
  TISE_LAUNCH_CHECK(ctx);
  return TISE_OK;
}

int launch_avgpool_bn(tise_ctx* ctx, const __half* in, int n, int h, int w, int c, const float* scale,
                      const float* shift, void* out, int out_stride, int out_f32, cudaStream_t st) {
  constexpr int kXs = 4;  // output columns per thread
  const int segs = (w + kXs - 1) / kXs;
  const long long total = (long long)n * segs * (c / 8);
  if ((long long)n * h * w * c >= (1ll << 31) * 8 || total >= (1ll << 31))
    return fail(ctx, TISE_ERR_INVALID, "avgpool_bn: batch too large (%lld work items)", total);
  const int grid = grid_for(total, 128, ctx->num_sms);
  if (out_f32)
    avgpool_bn_kernel<true, kXs><<<grid, 128, 0, st>>>(in, n, h, w, c / 8, segs, scale, shift, out, out_stride);
  else
    avgpool_bn_kernel<false, kXs><<<grid, 128, 0, st>>>(in, n, h, w, c / 8, segs, scale, shift, out, out_stride);
  TISE_LAUNCH_CHECK(ctx);
  return TISE_OK;
}

int launch_global_avgpool(tise_ctx* ctx, const void* in, int in_f32, int n, int hw, int c, int in_stride,
                          float* out, cudaStream_t st) {
  const int total = n * c;
  if (in_f32)
    global_avgpool_kernel<float><<<(total + 255) / 256, 256, 0, st>>>((const float*)in, n, hw, c, in_stride, out);
  else
    global_avgpool_kernel<__half><<<(total + 255) / 256, 256, 0, st>>>((const __half*)in, n, hw, c, in_stride, out);
  TISE_LAUNCH_CHECK(ctx);
  return TISE_OK;
}

}  // namespace tise
