// Temperature calibration of the classifier behind IS* / O-IS:
// classifier_calibration/temperature_scaling.py:34-77 optimises one scalar T on the validation logits with LBFGS on
// nn.CrossEntropyLoss(logits / T, labels), and :80-119 reports the expected calibration error.  The optimiser stays
// on the host (torch.optim.LBFGS over one scalar); each closure evaluation is ONE launch of the kernel below, which
// returns the loss, its derivative dNLL/dT in closed form, and the ECE histogram of the same logits / T:
//   nll_i      = log sum_c exp(l_c / T) - l_y / T
//   d nll_i/dT = (l_y - sum_c p_c l_c) / T^2,   p = softmax(l / T)
//   confidence = max_c p_c = 1 / sum_c exp(l_c / T - max),  prediction = argmax_c, bin b: lower_b < conf <= upper_b
#include "common.cuh"

namespace tise {

constexpr int kMaxBins = 64;
struct BinBounds {
  float v[kMaxBins + 1];
};

// out: [0] sum nll, [1] sum dnll/dT, [2] rows, [3] correct predictions, then count[n_bins], conf[n_bins], acc[n_bins]
__global__ void __launch_bounds__(256) calibration_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                                                          int n, int C, float T, int n_bins, BinBounds bounds,
                                                          double* __restrict__ out, int* __restrict__ bad_label) {
  __shared__ double s_acc[4 + 3 * kMaxBins];
  const int n_acc = 4 + 3 * n_bins;
  for (int i = threadIdx.x; i < n_acc; i += blockDim.x) s_acc[i] = 0.0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row < n) {
    const float* l = logits + (size_t)row * C;
    float mx = -INFINITY;
    int am = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {
      const float v = l[c] / T;
      if (v > mx) {  // strict: the first maximum wins inside a lane
        mx = v;
        am = c;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float omx = __shfl_xor_sync(0xffffffffu, mx, o);
      const int oam = __shfl_xor_sync(0xffffffffu, am, o);
      if (omx > mx || (omx == mx && oam < am)) {
        mx = omx;
        am = oam;
      }
    }
    float sum = 0.f;
    double sl = 0.0;
    for (int c = lane; c < C; c += 32) {
      const float e = expf(l[c] / T - mx);
      sum += e;
      sl += (double)e * (double)l[c];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sum += __shfl_xor_sync(0xffffffffu, sum, o);
      sl += __shfl_xor_sync(0xffffffffu, sl, o);
    }
    if (lane == 0) {
      const long long y = labels[row];
      if (y < 0 || y >= C) {
        atomicExch(bad_label, 1);
      } else {
        const double ly = (double)l[y];
        const double nll = (double)logf(sum) + (double)mx - (double)(l[y] / T);
        const double dnll = (ly - sl / (double)sum) / ((double)T * (double)T);
        const float conf = 1.0f / sum;
        const int hit = am == (int)y;
        atomicAdd(&s_acc[0], nll);
        atomicAdd(&s_acc[1], dnll);
        atomicAdd(&s_acc[2], 1.0);
        atomicAdd(&s_acc[3], (double)hit);
        for (int b = 0; b < n_bins; ++b)
          if (conf > bounds.v[b] && conf <= bounds.v[b + 1]) {
            atomicAdd(&s_acc[4 + b], 1.0);
            atomicAdd(&s_acc[4 + n_bins + b], (double)conf);
            atomicAdd(&s_acc[4 + 2 * n_bins + b], (double)hit);
          }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_acc; i += blockDim.x)
    if (s_acc[i] != 0.0) atomicAdd(out + i, s_acc[i]);
}

}  // namespace tise

using namespace tise;

extern "C" int tise_calibration_eval(tise_ctx* ctx, const float* d_logits, const int64_t* d_labels, int n, int C,
                                     float temperature, const float* h_bin_bounds, int n_bins, double* d_out,
                                     tise_stream stream) {
  if (!ctx) return TISE_ERR_INVALID;
  if (n == 0) return TISE_OK;
  if (!d_logits || !d_labels || !d_out || n < 0 || C <= 0) return fail(ctx, TISE_ERR_INVALID, "tise_calibration_eval: null argument");
  if (!(temperature > 0.f) && !(temperature < 0.f))
    return fail(ctx, TISE_ERR_INVALID, "tise_calibration_eval: temperature must be non-zero and finite");
  if (n_bins < 0 || n_bins > kMaxBins || (n_bins > 0 && !h_bin_bounds))
    return fail(ctx, TISE_ERR_INVALID, "tise_calibration_eval: n_bins = %d (max %d)", n_bins, kMaxBins);
  BinBounds bb;
  for (int i = 0; i <= n_bins; ++i) bb.v[i] = h_bin_bounds[i];
  // the last slot of d_out's caller-zeroed tail doubles as the bad-label flag: keep it separate and explicit instead
  int* d_flag = reinterpret_cast<int*>(d_out + 4 + 3 * n_bins);
  calibration_kernel<<<cdiv(n, 8), 256, 0, (cudaStream_t)stream>>>(d_logits, (const long long*)d_labels, n, C, temperature,
                                                                  n_bins, bb, d_out, d_flag);
  TISE_LAUNCH_CHECK(ctx);
  return TISE_OK;
}
