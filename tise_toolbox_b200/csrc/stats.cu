// Statistics path: fc head, FID moments (A7) and inception-score partial sums (A11-A13).
// All of it is a few MFLOP per image next to the trunk's 11.4 GFLOP, so these kernels favour
// precision (fp64 accumulation of exact fp32 products): the covariance sums run as an fp64 SYRK on the tensor pipe
// (DMMA), the fc head and the softmax / KL sums as plain coalesced kernels.
#include "net.cuh"

namespace tise {

// ------------------------------------------------------------------------------------ fc head
// logits[n][C] = feat[n][2048] . W[C][2048]^T (+ b), fp32 FMA, one output per thread accumulated over k = 0 .. K - 1 in
// order.  8 x 32 output tiles: a batch of 1024 rows x 51 classes is 256 CTAs (the 32 x 32 tiles of round 1 gave 64 CTAs
// on 148 SMs and 170 us per 1024 images for 0.2 GFLOP).
constexpr int LBM = 8, LBN = 32, LBK = 64;
__global__ void __launch_bounds__(256) logits_kernel(const float* __restrict__ feat, const float* __restrict__ w,
                                                     const float* __restrict__ bias, int n, int C, int K,
                                                     float* __restrict__ out) {
  __shared__ float sA[LBM][LBK];
  __shared__ float sB[LBK][LBN + 1];
  const int tx = threadIdx.x % LBN, ty = threadIdx.x / LBN;
  const int m0 = blockIdx.x * LBM, n0 = blockIdx.y * LBN;
  float acc = 0.f;
  for (int k0 = 0; k0 < K; k0 += LBK) {
    for (int i = threadIdx.x; i < LBM * LBK; i += 256) {
      const int r = i / LBK, k = i % LBK;
      sA[r][k] = (m0 + r < n && k0 + k < K) ? feat[(size_t)(m0 + r) * K + k0 + k] : 0.f;
    }
    for (int i = threadIdx.x; i < LBN * LBK; i += 256) {
      const int r = i / LBK, k = i % LBK;
      sB[k][r] = (n0 + r < C && k0 + k < K) ? w[(size_t)(n0 + r) * K + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < LBK; ++k) acc = fmaf(sA[ty][k], sB[k][tx], acc);
    __syncthreads();
  }
  const int r = m0 + ty, c = n0 + tx;
  if (r < n && c < C) out[(size_t)r * C + c] = acc + (bias ? bias[c] : 0.f);
}

// ------------------------------------------------------------------------------------ moments
// sum[j] += sum_i x[i][j]
__global__ void moments_sum_kernel(const float* __restrict__ x, int n, int d, double* __restrict__ sum) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= d) return;
  const int i0 = blockIdx.y * 256, i1 = min(n, i0 + 256);
  double acc = 0.0;
  for (int i = i0; i < i1; ++i) acc += (double)x[(size_t)i * d + j];
  atomicAdd(sum + j, acc);
}

// sxx[a][b] += sum_i x[i][a] x[i][b] for the lower triangle: a SYRK on the fp64 TENSOR pipe
// (mma.sync.m8n8k4.f64 -> DMMA), exact to fp64 rounding like the FMA kernel it replaces (fid_score.py:194-195's
// np.cov is an fp64 GEMM too).  One CTA owns a 128 x 128 tile (ta >= tb) for a chunk of up to 1024 images: 16 warps,
// each a 32 x 32 sub-tile = 4 x 4 DMMA tiles (32 fp64 accumulators per thread); the images stream through shared
// memory 16 at a time as doubles (fp32 -> fp64 is exact), register-prefetched one chunk ahead.  Fragment layout of
// m8n8k4 (PTX ISA): lane T holds A[T / 4][T % 4], B[T % 4][T / 4] and C[T / 4][2 (T % 4) + {0, 1}]; with A = X^T and
// B = X both fragments are the same kind of read, xs[k0 + T % 4][col0 + T / 4] -- a row pitch of 132 doubles makes it
// conflict-free.  One red.add.f64 per accumulator at the end (the K split over CTAs is the only sharing).
constexpr int ST = 128, SKC = 16, SPITCH = ST + 4, SCHUNK = 1024;
constexpr int MT = 64;  // d must be a multiple of 64 (a last half tile is masked)

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(512) moments_sxx_kernel(const float* __restrict__ x, int n, int d,
                                                          double* __restrict__ sxx) {
  // blockIdx.x enumerates lower-triangular tile pairs
  int t = blockIdx.x, ta = 0;
  while (t > ta) {
    t -= ta + 1;
    ++ta;
  }
  const int tb = t;
  __shared__ double sa[SKC][SPITCH];
  __shared__ double sb[SKC][SPITCH];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wa = (warp >> 2) * 32, wb = (warp & 3) * 32;  // this warp's 32 x 32 sub-tile
  const int fr = lane >> 2, fk = lane & 3;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const int i0 = blockIdx.y * SCHUNK, i1 = min(n, i0 + SCHUNK);
  // loader: thread -> (image k = tid / 32, 4 consecutive columns c = 4 (tid % 32)) of both column blocks
  const int lk = threadIdx.x >> 5, lc = (threadIdx.x & 31) * 4;
  const int ca = ta * ST + lc, cb = tb * ST + lc;
  float4 ra, rb;
  auto fetch = [&](int k0) {
    const int img = k0 + lk;
    const bool ok = img < i1;
    ra = (ok && ca < d) ? __ldg(reinterpret_cast<const float4*>(x + (size_t)img * d + ca)) : make_float4(0.f, 0.f, 0.f, 0.f);
    rb = (ok && cb < d) ? __ldg(reinterpret_cast<const float4*>(x + (size_t)img * d + cb)) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  fetch(i0);
  for (int k0 = i0; k0 < i1; k0 += SKC) {
    sa[lk][lc] = (double)ra.x; sa[lk][lc + 1] = (double)ra.y; sa[lk][lc + 2] = (double)ra.z; sa[lk][lc + 3] = (double)ra.w;
    sb[lk][lc] = (double)rb.x; sb[lk][lc + 1] = (double)rb.y; sb[lk][lc + 2] = (double)rb.z; sb[lk][lc + 3] = (double)rb.w;
    __syncthreads();
    if (k0 + SKC < i1) fetch(k0 + SKC);  // next chunk in flight during the DMMAs
#pragma unroll
    for (int kk = 0; kk < SKC; kk += 4) {
      double af[4], bf[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = sa[kk + fk][wa + 8 * i + fr];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = sb[kk + fk][wb + 8 * j + fr];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int a = ta * ST + wa + 8 * i + fr, b = tb * ST + wb + 8 * j + 2 * fk;
      if (a < d && b < d) {  // (tiles on the diagonal also fill their upper half: the finalize kernel reads a >= b only)
        atomicAdd(sxx + (size_t)a * d + b, acc[i][j][0]);
        atomicAdd(sxx + (size_t)a * d + b + 1, acc[i][j][1]);
      }
    }
}

// mu = sum / count; sigma[a][b] = (sxx[max][min] - count mu_a mu_b) / (count - 1)
__global__ void moments_finalize_kernel(const double* __restrict__ sum, const double* __restrict__ sxx, double count,
                                        int d, double* __restrict__ mu, double* __restrict__ sigma) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)d * d) return;
  const int a = (int)(idx / d), b = (int)(idx % d);
  const double ma = sum[a] / count, mb = sum[b] / count;
  const double s = a >= b ? sxx[(size_t)a * d + b] : sxx[(size_t)b * d + a];
  sigma[idx] = (s - count * (ma * mb)) / (count - 1.0);  // (ma * mb) first: exactly symmetric
  if (b == 0) mu[a] = ma;
}

// ------------------------------------------------------------------------------------ inception score
// One warp per image: p = softmax(logits[class_offset:] / T) in fp32 (as torch / TF compute it),
// then fp64 partial sums per split.  A CTA accumulates the marginal of the split of its first
// image in shared memory (the common case: splits are contiguous index ranges) and falls back to
// global atomics for images of any other split.
constexpr int IS_IMGS_PER_CTA = 64;
__global__ void __launch_bounds__(256) is_update_kernel(const float* __restrict__ logits, int n, int C, int off,
                                                        float temperature, const int32_t* __restrict__ split_id,
                                                        int n_splits, double* __restrict__ splogp,
                                                        double* __restrict__ marg, long long* __restrict__ cnt,
                                                        long long* __restrict__ zeros) {
  extern __shared__ double s_marg[];  // Cs
  const int Cs = C - off;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.x * IS_IMGS_PER_CTA, i1 = min(n, i0 + IS_IMGS_PER_CTA);
  int cta_split = -1;
  for (int i = i0; i < i1 && cta_split < 0; ++i) cta_split = split_id[i];  // first kept image's split
  for (int c = threadIdx.x; c < Cs; c += blockDim.x) s_marg[c] = 0.0;
  __syncthreads();

  for (int i = i0 + warp; i < i1; i += 8) {
    const int sp = split_id[i];
    if (sp < 0 || sp >= n_splits) continue;
    const float* row = logits + (size_t)i * C + off;
    float mx = -INFINITY;
    for (int c = lane; c < Cs; c += 32) mx = fmaxf(mx, row[c] / temperature);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int c = lane; c < Cs; c += 32) sum += expf(row[c] / temperature - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    double plogp = 0.0;
    long long nz = 0;
    for (int c = lane; c < Cs; c += 32) {
      const float p = expf(row[c] / temperature - mx) / sum;
      const double pd = (double)p;
      if (p > 0.f)
        plogp += pd * log(pd);
      else
        ++nz;
      if (sp == cta_split)
        atomicAdd(&s_marg[c], pd);
      else
        atomicAdd(marg + (size_t)sp * Cs + c, pd);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      plogp += __shfl_xor_sync(0xffffffffu, plogp, o);
      nz += __shfl_xor_sync(0xffffffffu, nz, o);
    }
    if (lane == 0) {
      atomicAdd(splogp + sp, plogp);
      atomicAdd((unsigned long long*)(cnt + sp), 1ull);
      if (nz) atomicAdd((unsigned long long*)zeros, (unsigned long long)nz);
    }
  }
  __syncthreads();
  if (cta_split >= 0 && cta_split < n_splits)
    for (int c = threadIdx.x; c < Cs; c += blockDim.x)
      if (s_marg[c] != 0.0) atomicAdd(marg + (size_t)cta_split * Cs + c, s_marg[c]);
}

}  // namespace tise

using namespace tise;

extern "C" int tise_logits(tise_ctx* ctx, const float* d_feat, int n, int use_bias, float* d_logits,
                           tise_stream stream) {
  if (!ctx) return TISE_ERR_INVALID;
  if (!ctx->net || ctx->net->num_classes <= 0) return fail(ctx, TISE_ERR_STATE, "tise_logits: no classifier head loaded");
  if (n == 0) return TISE_OK;
  if (!d_feat || !d_logits || n < 0) return fail(ctx, TISE_ERR_INVALID, "tise_logits: null argument");
  const int C = ctx->net->num_classes;
  dim3 grid((n + LBM - 1) / LBM, (C + LBN - 1) / LBN);
  logits_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_feat, ctx->net->d_fc_w, use_bias ? ctx->net->d_fc_b : nullptr,
                                                        n, C, 2048, d_logits);
  TISE_LAUNCH_CHECK(ctx);
  return TISE_OK;
}

extern "C" int tise_moments_update(tise_ctx* ctx, const float* d_feat, int n, int d, double* d_sum, double* d_sxx,
                                   tise_stream stream) {
  if (!ctx) return TISE_ERR_INVALID;
  if (n == 0) return TISE_OK;
  if (!d_feat || !d_sum || !d_sxx || n < 0) return fail(ctx, TISE_ERR_INVALID, "tise_moments_update: null argument");
  if (d <= 0 || d % MT) return fail(ctx, TISE_ERR_INVALID, "tise_moments_update: d = %d must be a multiple of %d", d, MT);
  cudaStream_t st = (cudaStream_t)stream;
  moments_sum_kernel<<<dim3((d + 255) / 256, (n + 255) / 256), 256, 0, st>>>(d_feat, n, d, d_sum);
  TISE_LAUNCH_CHECK(ctx);
  const int tiles = (d + ST - 1) / ST;
  moments_sxx_kernel<<<dim3(tiles * (tiles + 1) / 2, (n + SCHUNK - 1) / SCHUNK), 512, 0, st>>>(d_feat, n, d, d_sxx);
  TISE_LAUNCH_CHECK(ctx);
  return TISE_OK;
}

extern "C" int tise_moments_finalize(tise_ctx* ctx, const double* d_sum, const double* d_sxx, int64_t count, int d,
                                     double* d_mu, double* d_sigma, tise_stream stream) {
  if (!ctx) return TISE_ERR_INVALID;
  if (!d_sum || !d_sxx || !d_mu || !d_sigma || d <= 0)
    return fail(ctx, TISE_ERR_INVALID, "tise_moments_finalize: null argument");
  if (count < 1) return fail(ctx, TISE_ERR_INVALID, "tise_moments_finalize: count = %lld", (long long)count);
  const size_t total = (size_t)d * d;
  moments_finalize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_sum, d_sxx, (double)count,
                                                                                          d, d_mu, d_sigma);
  TISE_LAUNCH_CHECK(ctx);
  return TISE_OK;
}

extern "C" int tise_is_update(tise_ctx* ctx, const float* d_logits, int n, int C, int class_offset, float temperature,
                              const int32_t* d_split_id, int n_splits, double* d_splogp, double* d_marg,
                              long long* d_cnt, long long* d_zeros, tise_stream stream) {
  if (!ctx) return TISE_ERR_INVALID;
  if (n == 0) return TISE_OK;
  if (!d_logits || !d_split_id || !d_splogp || !d_marg || !d_cnt || !d_zeros || n < 0)
    return fail(ctx, TISE_ERR_INVALID, "tise_is_update: null argument");
  if (C <= 0 || class_offset < 0 || class_offset >= C || n_splits <= 0 || !(temperature > 0.f))
    return fail(ctx, TISE_ERR_INVALID, "tise_is_update: bad C/class_offset/n_splits/temperature");
  const int Cs = C - class_offset;
  if ((size_t)Cs * sizeof(double) > 40000) return fail(ctx, TISE_ERR_INVALID, "tise_is_update: too many classes (%d)", Cs);
  is_update_kernel<<<(n + IS_IMGS_PER_CTA - 1) / IS_IMGS_PER_CTA, 256, Cs * sizeof(double), (cudaStream_t)stream>>>(
      d_logits, n, C, class_offset, temperature, d_split_id, n_splits, d_splogp, d_marg, d_cnt, d_zeros);
  TISE_LAUNCH_CHECK(ctx);
  return TISE_OK;
}
