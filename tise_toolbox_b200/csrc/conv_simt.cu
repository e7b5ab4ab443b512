// TISE_CONV_SIMT: CUDA-core implicit-GEMM convolution with the same numeric model as the tcgen05
// path (fp16 activations, hi + lo fp16 weights, fp32 accumulation, fp32 scale/shift/ReLU epilogue).
// It exists so that the tests can cross-check the UMMA kernels on the device at sizes the CPU
// oracle cannot reach; it is not the product path and is ~20x slower.
#include "common.cuh"

namespace tise {

constexpr int SBM = 64, SBN = 64, SBK = 16;

struct SimtParams {
  const __half* in;
  const __half* wt;  // packed rows, see ConvLayer
  const float* scale;
  const float* shift;
  void* seg_out[kMaxSegs];
  int seg_end[kMaxSegs], seg_stride[kMaxSegs], seg_relu[kMaxSegs];
  int n, h, w, pitch_w, in_stride, cin, cin_pad, ktot, nt;
  int cout, R, S, stride, pad_h, pad_w, ho, wo, out_f32, passes;
  long long M;
};

__global__ void __launch_bounds__(256) conv_simt_kernel(SimtParams p) {
  __shared__ float sA[SBK][SBM + 4];
  __shared__ float sB[SBK][SBN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;  // 16 x 16 threads, 4 x 4 outputs each
  const long long m0 = (long long)blockIdx.x * SBM;
  const int n0 = blockIdx.y * SBN;

  // loader mapping: each thread loads one (row, k-quad): 64 rows x 4 quads of 4 channels
  const int lrow = tid / 4, lq = tid % 4;
  const long long lm = m0 + lrow;
  int limg = 0, loy = 0, lox = 0;
  const bool lvalid = lm < p.M;
  if (lvalid) {
    limg = (int)(lm / (p.ho * p.wo));
    const int rem = (int)(lm - (long long)limg * p.ho * p.wo);
    loy = rem / p.wo;
    lox = rem - loy * p.wo;
  }
  const int lco = n0 + lrow;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int r = 0; r < p.R; ++r)
    for (int s = 0; s < p.S; ++s) {
      const int iy = loy * p.stride - p.pad_h + r, ix = lox * p.stride - p.pad_w + s;
      const bool inb = lvalid && iy >= 0 && iy < p.h && ix >= 0 && ix < p.w;
      const __half* arow = inb ? p.in + (((size_t)limg * p.h + iy) * p.pitch_w + ix) * p.in_stride : p.in;
      const size_t whi = ((size_t)(lco / p.nt) * 2 * p.nt + lco % p.nt) * p.ktot + (size_t)(r * p.S + s) * p.cin_pad;
      const size_t wlo = whi + (size_t)p.nt * p.ktot;
      for (int c0 = 0; c0 < p.cin; c0 += SBK) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int ci = c0 + lq * 4 + j;
          float a = 0.f, b = 0.f;
          if (ci < p.cin) {
            if (inb) a = __half2float(arow[ci]);
            if (lco < p.cout) {
              b = __half2float(p.wt[whi + ci]);
              if (p.passes == 2) b += __half2float(p.wt[wlo + ci]);
            }
          }
          sA[lq * 4 + j][lrow] = a;
          sB[lq * 4 + j][lrow] = b;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SBK; ++k) {
          float av[4], bv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) av[i] = sA[k][ty * 4 + i];
#pragma unroll
          for (int j = 0; j < 4; ++j) bv[j] = sB[k][tx * 4 + j];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
      }
    }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + tx * 4 + j;
      if (co >= p.cout) continue;
      float v = fmaf(acc[i][j], p.scale[co], p.shift[co]);
      int sg = 0;
      while (co >= p.seg_end[sg]) ++sg;
      const int c = co - (sg ? p.seg_end[sg - 1] : 0);
      if (p.seg_relu[sg]) v = fmaxf(v, 0.f);
      if (p.out_f32)
        ((float*)p.seg_out[sg])[(size_t)m * p.seg_stride[sg] + c] = v;
      else
        ((__half*)p.seg_out[sg])[(size_t)m * p.seg_stride[sg] + c] = __float2half_rn(fminf(v, 65504.f));
    }
  }
}

int launch_conv_simt(tise_ctx* ctx, const ConvLayer& L, const ConvArgs& a, cudaStream_t st) {
  SimtParams p;
  p.in = a.in;
  p.wt = L.d_w;
  p.nt = L.nt;
  p.scale = L.d_scale;
  p.shift = L.d_shift;
  int end = 0;
  for (int i = 0; i < kMaxSegs; ++i) {
    if (i < a.n_segs) end += a.segs[i].c;
    p.seg_out[i] = a.segs[i < a.n_segs ? i : 0].out;
    p.seg_stride[i] = a.segs[i < a.n_segs ? i : 0].out_stride;
    p.seg_relu[i] = i < a.n_segs ? a.segs[i].relu : 0;
    p.seg_end[i] = i < a.n_segs ? end : 0x7fffffff;
  }
  p.n = a.n;
  p.h = a.h;
  p.w = a.w;
  p.pitch_w = a.in_row_pitch ? a.in_row_pitch : a.w;
  p.in_stride = a.in_stride;
  p.cin = L.cin;
  p.cin_pad = L.cin_pad;
  p.ktot = L.ktot;
  p.cout = L.cout;
  p.R = L.R;
  p.S = L.S;
  p.stride = L.stride;
  p.pad_h = L.pad_h;
  p.pad_w = L.pad_w;
  p.ho = a.ho;
  p.wo = a.wo;
  p.out_f32 = a.out_f32;
  p.passes = ctx->passes;
  p.M = (long long)a.n * a.ho * a.wo;
  dim3 grid((unsigned)((p.M + SBM - 1) / SBM), (unsigned)((L.cout + SBN - 1) / SBN));
  conv_simt_kernel<<<grid, 256, 0, st>>>(p);
  TISE_LAUNCH_CHECK(ctx);
  return TISE_OK;
}

}  // namespace tise
