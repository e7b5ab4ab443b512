// A9: Frechet distance, fp64 on the device (see include/tise_b200.h: tise_frechet).
//
// The reference evaluates tr sqrtm(s1 s2) with scipy's Schur-based sqrtm of the non-symmetric
// product (fid_score.py:155), an O(25 d^3) poorly parallel CPU routine (8-10 s at d = 2048).
// The eigenvalues of s1 s2 are those of the symmetric matrix M = L^T s2 L with s1 = L L^T, so
//     tr sqrt(s1 s2) = sum_i sqrt(lambda_i(M)),  lambda_i >= 0,
// which needs no eigenvectors and no matrix square root:
//   1. A = s1 + delta I (delta = 1e-11 * mean eigenvalue keeps rank-deficient s1 factorisable; its
//      effect on the trace is O(sqrt(delta)) on the null space only), blocked Cholesky A = L L^T;
//   2. M = L^T (s2 L): two fp64 GEMMs, then symmetrise;
//   3. Householder tridiagonalisation of M (two kernels per column: reflector + symmetric
//      matrix-vector product, then the rank-2 update), the matrix stays L2-resident;
//   4. Sturm-sequence bisection, one thread per eigenvalue, then sum sqrt(max(lambda, 0)).
// Everything is fp64 FMA on the CUDA cores; the two GEMMs are the only O(d^3) dense steps.
#include <math.h>
#include <stdlib.h>

#include "net.cuh"

namespace tise {

// scalars[] slots
enum { FR_DIFF2 = 0, FR_TR1, FR_TR2, FR_TAU, FR_TRSQRT, FR_NEGMAX, FR_BADPIV, FR_NSLOTS = 16 };

// ---------------------------------------------------------------------------- small helpers
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum, result valid in every thread; red must hold 32 doubles
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
  if (warp == 0) {
    t = warp_sum(t);
    if (lane == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

// |mu1 - mu2|^2, tr s1, tr s2
__global__ void fr_scalars_kernel(const double* mu1, const double* s1, const double* mu2, const double* s2, int d,
                                  double* scal) {
  __shared__ double red[32];
  double a = 0, b = 0, c = 0;
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    const double df = mu1[i] - mu2[i];
    a += df * df;
    b += s1[(size_t)i * d + i];
    c += s2[(size_t)i * d + i];
  }
  a = block_sum(a, red);
  b = block_sum(b, red);
  c = block_sum(c, red);
  if (threadIdx.x == 0) {
    scal[FR_DIFF2] = a;
    scal[FR_TR1] = b;
    scal[FR_TR2] = c;
    scal[FR_TRSQRT] = 0.0;
    scal[FR_NEGMAX] = 0.0;
    scal[FR_BADPIV] = 0.0;
  }
}

// dst = src + (add + rel * max(tr, 0)) I
__global__ void fr_copy_shift_kernel(const double* src, double* dst, int d, double add, double rel, const double* scal,
                                     int tr_slot) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)d * d) return;
  const int r = (int)(i / d), c = (int)(i % d);
  double v = src[i];
  if (r == c) v += add + rel * fmax(scal[tr_slot], 0.0);
  dst[i] = v;
}

// ---------------------------------------------------------------------------- blocked Cholesky (lower, row-major)
constexpr int CNB = 32;

// factor the nb x nb diagonal block at (k, k) in shared memory; non-positive pivots are clamped
__global__ void __launch_bounds__(CNB* CNB) chol_diag_kernel(double* A, int d, int k, int nb, double rel,
                                                             double* scal) {
  __shared__ double s[CNB][CNB + 1];
  const int i = threadIdx.y, j = threadIdx.x;
  if (i < nb && j < nb) s[i][j] = A[(size_t)(k + i) * d + k + j];
  __syncthreads();
  for (int c = 0; c < nb; ++c) {
    if (i == c && j == c) {
      // the jitter delta = rel * tr(s1) should dominate the rounding noise of the Schur complement; a pivot below
      // delta / 4 means it did not: clamp (keeps everything finite) and report, the caller retries with 100 x delta
      const double delta = rel * fmax(scal[FR_TR1], 1e-300);
      double p = s[c][c];
      if (!(p > 0.25 * delta)) {
        p = 0.25 * delta;
        scal[FR_BADPIV] = 1.0;
      }
      s[c][c] = sqrt(p);
    }
    __syncthreads();
    if (j == c && i > c && i < nb) s[i][c] /= s[c][c];
    __syncthreads();
    if (j > c && j < nb && i >= j && i < nb) s[i][j] -= s[i][c] * s[j][c];
    __syncthreads();
  }
  if (i < nb && j < nb) A[(size_t)(k + i) * d + k + j] = (j <= i) ? s[i][j] : 0.0;
}

// rows below the diagonal block: X = A[i, k:k+nb] * Lkk^-T  (one thread per row)
__global__ void chol_trsm_kernel(double* A, int d, int k, int nb) {
  __shared__ double l[CNB][CNB + 1];
  for (int t = threadIdx.x; t < nb * nb; t += blockDim.x) l[t / nb][t % nb] = A[(size_t)(k + t / nb) * d + k + t % nb];
  __syncthreads();
  const int i = k + nb + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d) return;
  double x[CNB];
  double* row = A + (size_t)i * d + k;
#pragma unroll
  for (int j = 0; j < CNB; ++j) x[j] = (j < nb) ? row[j] : 0.0;
#pragma unroll
  for (int j = 0; j < CNB; ++j) {
    if (j < nb) {
      double v = x[j];
#pragma unroll
      for (int c = 0; c < CNB; ++c)
        if (c < j) v -= x[c] * l[j][c];
      x[j] = v / l[j][j];
    }
  }
#pragma unroll
  for (int j = 0; j < CNB; ++j)
    if (j < nb) row[j] = x[j];
}

// trailing update of the lower triangle: A[i][j] -= sum_c P[i][c] P[j][c], i >= j >= k + nb, P = A[:, k:k+nb]
__global__ void __launch_bounds__(256) chol_syrk_kernel(double* A, int d, int k, int nb) {
  const int base = k + nb;
  const int ti = blockIdx.y, tj = blockIdx.x;
  if (tj > ti) return;
  __shared__ double pi[CNB][64 + 1];
  __shared__ double pj[CNB][64 + 1];
  const int i0 = base + ti * 64, j0 = base + tj * 64;
  for (int t = threadIdx.x; t < 64 * nb; t += 256) {
    const int r = t / nb, c = t % nb;
    pi[c][r] = (i0 + r < d) ? A[(size_t)(i0 + r) * d + k + c] : 0.0;
    pj[c][r] = (j0 + r < d) ? A[(size_t)(j0 + r) * d + k + c] : 0.0;
  }
  __syncthreads();
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  double acc[4][4] = {};
  for (int c = 0; c < nb; ++c) {
    double a[4], b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = pi[c][ty * 4 + u];
#pragma unroll
    for (int v = 0; v < 4; ++v) b[v] = pj[c][tx * 4 + v];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[u][v] = fma(a[u], b[v], acc[u][v]);
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int i = i0 + ty * 4 + u, j = j0 + tx * 4 + v;
      if (i < d && j < d && j <= i) A[(size_t)i * d + j] -= acc[u][v];
    }
}

__global__ void zero_upper_kernel(double* A, int d) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)d * d) return;
  if ((int)(i % d) > (int)(i / d)) A[i] = 0.0;
}

// ---------------------------------------------------------------------------- fp64 GEMM  C = op(A) * B, row-major d x d
template <bool TA>
__global__ void __launch_bounds__(256) dgemm_kernel(const double* __restrict__ A, const double* __restrict__ B,
                                                    double* __restrict__ Cm, int d) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ double sa[BK][BM + 1];
  __shared__ double sb[BK][BN + 1];
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  // Both products have a lower-triangular factor (L[k][j] = 0 for k < j, exact zeros from zero_upper_kernel):
  //   TA = false: T = S2 L      -> only k >= n0 contributes to column block n0        (half of the k range on average)
  //   TA = true:  M = L^T T     -> only k >= m0 contributes to row block m0, and M is symmetric: tiles above the
  //                                diagonal are not computed at all (mirror_lower_kernel fills them)   (a quarter)
  if (TA && n0 > m0) return;
  const int k_begin = TA ? m0 : n0;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  double acc[4][4] = {};
  for (int k0 = k_begin; k0 < d; k0 += BK) {
    for (int t = threadIdx.x; t < BM * BK; t += 256) {
      if (TA) {  // op(A)[m][k] = A[k][m]: consecutive threads walk m (contiguous)
        const int kk = t / BM, mm = t % BM;
        sa[kk][mm] = (m0 + mm < d && k0 + kk < d) ? A[(size_t)(k0 + kk) * d + m0 + mm] : 0.0;
      } else {
        const int mm = t / BK, kk = t % BK;
        sa[kk][mm] = (m0 + mm < d && k0 + kk < d) ? A[(size_t)(m0 + mm) * d + k0 + kk] : 0.0;
      }
      const int kb = t / BN, nn = t % BN;
      sb[kb][nn] = (n0 + nn < d && k0 + kb < d) ? B[(size_t)(k0 + kb) * d + n0 + nn] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = sa[kk][ty * 4 + u];
#pragma unroll
      for (int v = 0; v < 4; ++v) b[v] = sb[kk][tx * 4 + v];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fma(a[u], b[v], acc[u][v]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int i = m0 + ty * 4 + u, j = n0 + tx * 4 + v;
      if (i < d && j < d) Cm[(size_t)i * d + j] = acc[u][v];
    }
}

// M[j][i] = M[i][j] for j < i: the strictly upper triangle is a copy of the computed lower one
__global__ void symmetrize_kernel(double* M, int d) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)d * d) return;
  const int i = (int)(idx / d), j = (int)(idx % d);
  if (j >= i) return;
  M[(size_t)j * d + i] = M[(size_t)i * d + j];
}

// ---------------------------------------------------------------------------- Householder tridiagonalisation
// Step k works on the trailing matrix M22 = M[k+1:, k+1:] (size m) with x = M[k, k+1:].
// Kernel 1: every CTA rebuilds the reflector v (v0 = 1) from x in shared memory (cheap, avoids a
//           launch), then its warps compute p_i = tau * sum_j M22[i][j] v_j.  CTA 0 publishes v, tau, a_k, b_k.
constexpr int TRI_THREADS = 256;
__global__ void __launch_bounds__(TRI_THREADS) tri_reflect_symv_kernel(const double* __restrict__ M, int d, int k,
                                                                       double* __restrict__ vbuf,
                                                                       double* __restrict__ pbuf,
                                                                       double* __restrict__ diag,
                                                                       double* __restrict__ off, double* scal) {
  extern __shared__ double sv[];  // m doubles
  __shared__ double red[32];
  const int m = d - k - 1;
  const double* x = M + (size_t)k * d + k + 1;
  double ss = 0.0;
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    const double xi = x[i];
    sv[i] = xi;
    if (i > 0) ss += xi * xi;
  }
  ss = block_sum(ss, red);
  const double alpha = sv[0];
  double tau, beta, scale;
  if (ss == 0.0) {
    tau = 0.0;
    beta = alpha;
    scale = 0.0;
  } else {
    beta = -copysign(sqrt(alpha * alpha + ss), alpha);
    tau = (beta - alpha) / beta;
    scale = 1.0 / (alpha - beta);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < m; i += blockDim.x) sv[i] = (i == 0) ? 1.0 : sv[i] * scale;
  __syncthreads();
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < m; i += blockDim.x) vbuf[i] = sv[i];
    if (threadIdx.x == 0) {
      scal[FR_TAU] = tau;
      diag[k] = M[(size_t)k * d + k];
      off[k] = beta;
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = blockIdx.x * (TRI_THREADS / 32) + warp; i < m; i += gridDim.x * (TRI_THREADS / 32)) {
    const double* row = M + (size_t)(k + 1 + i) * d + k + 1;
    double acc = 0.0;
    for (int j = lane; j < m; j += 32) acc = fma(row[j], sv[j], acc);
    acc = warp_sum(acc);
    if (lane == 0) pbuf[i] = tau * acc;
  }
}

// Kernel 2: w = p - (tau/2)(p.v) v, M22 -= v w^T + w v^T.  Each CTA recomputes the dot product.
__global__ void __launch_bounds__(TRI_THREADS) tri_rank2_kernel(double* __restrict__ M, int d, int k,
                                                                const double* __restrict__ vbuf,
                                                                const double* __restrict__ pbuf, const double* scal) {
  extern __shared__ double sm[];  // v[m], w[m]
  __shared__ double red[32];
  const int m = d - k - 1;
  double* v = sm;
  double* w = sm + m;
  double dot = 0.0;
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    const double vi = vbuf[i], pi = pbuf[i];
    v[i] = vi;
    w[i] = pi;
    dot += vi * pi;
  }
  dot = block_sum(dot, red);
  const double c = 0.5 * scal[FR_TAU] * dot;
  for (int i = threadIdx.x; i < m; i += blockDim.x) w[i] -= c * v[i];
  __syncthreads();
  // rows strided over CTAs, columns over threads (coalesced)
  for (int i = blockIdx.x; i < m; i += gridDim.x) {
    double* row = M + (size_t)(k + 1 + i) * d + k + 1;
    const double vi = v[i], wi = w[i];
    for (int j = threadIdx.x; j < m; j += blockDim.x) row[j] -= vi * w[j] + wi * v[j];
  }
}

// ---------------------------------------------------------------------------- persistent tridiagonalisation
// The two kernels above, run per column, cost 2 x 2046 launches whose ~10 us launch + drain latency was 60 % of
// tise_frechet at d = 2048 (profiles/frechet_launches_r01.txt).  This kernel walks all columns in ONE cooperative
// launch: every CTA rebuilds the reflector of column k from row k of M (so nothing but p is exchanged), the CTAs
// split the rows of the symmetric matrix-vector product and of the rank-2 update, and a grid-wide barrier (monotonic
// atomic counter, acquire / release at gpu scope) separates the two phases.  Data written by other CTAs is read
// with ld.global.cg (L2): L1 is not coherent across SMs inside one launch.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while (v < target);
    __threadfence();
  }
  __syncthreads();
}

// Reflector of x (length n, in shared memory, x[0] = alpha): returns tau and beta, writes v (v[0] = 1) to out.
__device__ __forceinline__ void make_reflector(const double* x, int n, double* out, double* red, double& tau, double& beta) {
  double ss = 0.0;
  for (int i = 1 + threadIdx.x; i < n; i += TRI_THREADS) ss += x[i] * x[i];
  ss = block_sum(ss, red);
  const double alpha = x[0];
  double scale;
  if (ss == 0.0) {
    tau = 0.0;
    beta = alpha;
    scale = 0.0;
  } else {
    beta = -copysign(sqrt(alpha * alpha + ss), alpha);
    tau = (beta - alpha) / beta;
    scale = 1.0 / (alpha - beta);
  }
  for (int i = threadIdx.x; i < n; i += TRI_THREADS) out[i] = (i == 0) ? 1.0 : x[i] * scale;
  __syncthreads();
}

// Column k's rank-2 update and column k+1's matrix-vector product in ONE pass over the trailing matrix: once
// w_k is known every CTA can form the updated row k+1 on its own (O(m) work), hence the next reflector v_{k+1};
// the warp that updates a piece of row i then accumulates its part of p_{k+1}[i] = tau' sum_j Mnew[i][j] v_{k+1}[j]
// from the values it has in registers.  One grid barrier and one read + write of the trailing matrix per column (the
// two-kernel form: two launches, three passes).  The matrix is L2-resident and every step is a round trip to L2
// (~1 us), so the work is cut into (row, 512-column chunk) items with 16 independent loads in flight per lane, and
// the chunk partials of p are summed by the readers (no atomics: the result does not depend on the schedule).
constexpr int TRI_CHUNK = 512;
constexpr int TRI_MAX_CHUNKS = 8;  // d <= 4096

__device__ __forceinline__ void tri_pass(double* M, int d, int row0, int col0, int mn, const double* v, const double* w,
                                         const double* vn, double tau_n, bool update, double* p_out, int debug = 0) {
  // rows row0 + t, columns col0 + u, t, u in [0, mn); v / w are indexed from the trailing index of (row0, col0)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kWarps = TRI_THREADS / 32;
  const int nch = (mn + TRI_CHUNK - 1) / TRI_CHUNK;
  const int items = mn * nch;
  for (int it = blockIdx.x * kWarps + warp; it < items; it += gridDim.x * kWarps) {
    const int t = it / nch, ch = it - t * nch;
    double* row = M + (size_t)(row0 + t) * d + col0;
    const double vi = update ? v[t] : 0.0, wi = update ? w[t] : 0.0;
    double acc = 0.0;
    {
      double val[16];
      const int u0 = ch * TRI_CHUNK + lane;
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int u = u0 + e * 32;
        val[e] = u < mn ? __ldcg(row + u) : 0.0;
      }
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int u = u0 + e * 32;
        if (u < mn) {
          if (update) {
            val[e] -= vi * w[u] + wi * v[u];
            if (!(debug & 4)) row[u] = val[e];
          }
          acc = fma(val[e], vn[u], acc);
        }
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) p_out[(size_t)ch * d + t] = tau_n * acc;
  }
}

__global__ void __launch_bounds__(TRI_THREADS) tri_persistent_kernel(double* M, int d, double* pbuf0, double* pbuf1,
                                                                     double* diag, double* off, unsigned* bar, int debug) {
  extern __shared__ double sm[];  // v[d], w[d], r[d], vn[d]
  __shared__ double red[32];
  double* v = sm;
  double* w = sm + d;
  double* r = sm + 2 * (size_t)d;
  double* vn = sm + 3 * (size_t)d;
  const unsigned nb = gridDim.x;
  unsigned phase = 0;
  double tau, beta;
  // prologue: reflector of column 0 and p_0 = tau M22 v
  {
    const int m = d - 1;
    for (int i = threadIdx.x; i < m; i += TRI_THREADS) r[i] = __ldcg(M + 1 + i);
    __syncthreads();
    make_reflector(r, m, v, red, tau, beta);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      diag[0] = __ldcg(M);
      off[0] = beta;
    }
    tri_pass(M, d, 1, 1, m, nullptr, nullptr, v, tau, false, pbuf0);
    grid_barrier(bar, ++phase * nb);
  }
  for (int k = 0; k + 2 < d; ++k) {
    const int m = d - k - 1;  // trailing matrix of column k: rows / columns k+1 .. d-1
    // p_k is read from one buffer while p_{k+1} is written to the other: a CTA that is ahead must not overwrite the
    // vector a slower CTA is still reading
    const double* p_in = (k & 1) ? pbuf1 : pbuf0;
    double* p_out = (k & 1) ? pbuf0 : pbuf1;
    const int nch = (m + TRI_CHUNK - 1) / TRI_CHUNK;
    // w = p - (tau / 2)(p . v) v, and the old row k+1 (index 0 of the trailing matrix; nobody writes it during this
    // column) -- all loads of a thread issued before the first use
    const double* row0 = M + (size_t)(k + 1) * d + k + 1;
    double dot = 0.0;
    for (int i0 = threadIdx.x; i0 < m; i0 += 4 * TRI_THREADS) {
      double pv[4][TRI_MAX_CHUNKS], rv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = i0 + e * TRI_THREADS;
#pragma unroll
        for (int c = 0; c < TRI_MAX_CHUNKS; ++c) pv[e][c] = (i < m && c < nch) ? __ldcg(p_in + (size_t)c * d + i) : 0.0;
        rv[e] = i < m ? __ldcg(row0 + i) : 0.0;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = i0 + e * TRI_THREADS;
        if (i < m) {
          double pi = pv[e][0];
#pragma unroll
          for (int c = 1; c < TRI_MAX_CHUNKS; ++c) pi += pv[e][c];
          w[i] = pi;
          r[i] = rv[e];
          dot += v[i] * pi;
        }
      }
    }
    dot = block_sum(dot, red);
    const double c = 0.5 * tau * dot;
    for (int i = threadIdx.x; i < m; i += TRI_THREADS) w[i] -= c * v[i];
    __syncthreads();
    const double v0 = v[0], w0 = w[0];
    for (int j = threadIdx.x; j < m; j += TRI_THREADS) r[j] -= v0 * w[j] + w0 * v[j];  // updated row k+1
    __syncthreads();
    if (m == 2) {  // last column: the 2 x 2 tail is the end of the tridiagonal
      if (blockIdx.x == 0 && threadIdx.x == 0) {
        diag[d - 2] = r[0];
        off[d - 2] = r[1];
        diag[d - 1] = __ldcg(M + (size_t)(d - 1) * d + d - 1) - 2.0 * v[1] * w[1];
      }
      break;
    }
    double tau_n, beta_n;
    make_reflector(r + 1, m - 1, vn, red, tau_n, beta_n);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      diag[k + 1] = r[0];
      off[k + 1] = beta_n;
    }
    // fused pass over rows / columns k+2 .. d-1
    if (!(debug & 1)) tri_pass(M, d, k + 2, k + 2, m - 1, v + 1, w + 1, vn, tau_n, !(debug & 8), p_out, debug);
    if (!(debug & 2)) grid_barrier(bar, ++phase * nb);
    double* tmp = v;
    v = vn;
    vn = tmp;
    tau = tau_n;
  }
}

__global__ void tri_tail_kernel(const double* M, int d, double* diag, double* off) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (d >= 2) {
      diag[d - 2] = M[(size_t)(d - 2) * d + d - 2];
      off[d - 2] = M[(size_t)(d - 1) * d + d - 2];
    }
    diag[d - 1] = M[(size_t)(d - 1) * d + d - 1];
  }
}

// ---------------------------------------------------------------------------- Sturm bisection
// Thread i finds the i-th smallest eigenvalue of the symmetric tridiagonal (diag, off) and the block
// adds sqrt(max(lambda, 0)) into scal[FR_TRSQRT]; the most negative eigenvalue goes to FR_NEGMAX.
__global__ void __launch_bounds__(256) sturm_sqrt_trace_kernel(const double* __restrict__ diag,
                                                               const double* __restrict__ off, int d, double* scal) {
  extern __shared__ double sm[];  // a[d], b2[d]
  __shared__ double red[32];
  __shared__ double s_lo, s_hi;
  double* a = sm;
  double* b2 = sm + d;
  double lo = INFINITY, hi = -INFINITY;
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    const double ai = diag[i];
    const double bl = i > 0 ? fabs(off[i - 1]) : 0.0, br = i < d - 1 ? fabs(off[i]) : 0.0;
    a[i] = ai;
    b2[i] = i > 0 ? off[i - 1] * off[i - 1] : 0.0;
    lo = fmin(lo, ai - bl - br);
    hi = fmax(hi, ai + bl + br);
  }
  // block min / max via negated sums is awkward: use shared atomics-free two-pass reduce
  __syncthreads();
  {
    // reduce lo
    double v = lo;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = red[0];
      for (int q = 1; q < (int)(blockDim.x >> 5); ++q) t = fmin(t, red[q]);
      s_lo = t;
    }
    __syncthreads();
    v = hi;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = red[0];
      for (int q = 1; q < (int)(blockDim.x >> 5); ++q) t = fmax(t, red[q]);
      s_hi = t;
    }
    __syncthreads();
  }
  const double glo = s_lo, ghi = s_hi;
  const double norm = fmax(fabs(glo), fabs(ghi));
  const double tiny = fmax(norm * 1e-300, 1e-300);
  double part = 0.0, negmax = 0.0;
  // One WARP per eigenvalue, 32-way multisection: lane j counts the eigenvalues below l + (j + 1)(h - l) / 33, so the
  // bracket shrinks 33x per sweep of the d-long (serial, division-bound) Sturm recurrence: ~11 sweeps instead of the
  // ~55 of plain bisection.
  const int lane = threadIdx.x & 31, warp_in_grid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int warps_in_grid = (gridDim.x * blockDim.x) >> 5;
  for (int idx = warp_in_grid; idx < d; idx += warps_in_grid) {
    double l = glo, h = ghi;
    for (int it = 0; it < 16; ++it) {  // 33^11 > 2^53: eleven sweeps resolve a bracket of any width
      const double width = h - l;
      if (!(width > 4.0 * 2.220446049250313e-16 * fmax(fabs(l), fabs(h)))) break;  // at the resolution of fp64
      const double xq = l + width * (1.0 / 33.0) * (double)(lane + 1);
      const bool inside = xq > l && xq < h;
      // Sturm count = sign changes of the leading principal minors p_j(xq) = (a_j - xq) p_{j-1} - b_j^2 p_{j-2}:
      // the multiplicative form of the recurrence (no fp64 division on the d-long dependent chain; the quotient form
      // q_j = p_j / p_{j-1} costs ~300 cycles per step), both minors rescaled by a power of two whenever they leave
      // [2^-500, 2^500].  An exact zero counts like a negative quotient (sign opposite to its predecessor), and
      // b_j^2 == 0 splits the matrix: the recurrence restarts and the counts of the blocks add up.
      int cnt = 0;  // number of eigenvalues < xq
      double pp = 1.0, pc = a[0] - xq;
      bool neg = !(pc > 0.0);  // sign of pc relative to "+": a zero takes the sign opposite to its predecessor (+)
      cnt += neg;
      for (int j = 1; j < d; ++j) {
        const double bj = b2[j];
        double pn;
        bool nneg;
        if (bj == 0.0) {
          pn = a[j] - xq;
          pc = 1.0;
          nneg = !(pn > 0.0);
          cnt += nneg;
        } else {
          pn = fma(a[j] - xq, pc, -(bj * pp));
          nneg = pn == 0.0 ? !neg : pn < 0.0;
          cnt += nneg != neg;
          const double m = fabs(pn);
          if (m > 0x1p500) {
            pn *= 0x1p-500;
            pc *= 0x1p-500;
          } else if (m < 0x1p-500 && m != 0.0) {
            pn *= 0x1p500;
            pc *= 0x1p500;
          }
        }
        pp = pc;
        pc = pn;
        neg = nneg;
      }
      const unsigned above = __ballot_sync(0xffffffffu, inside && cnt > idx);   // candidates for the new upper end
      const unsigned below = __ballot_sync(0xffffffffu, inside && cnt <= idx);  // candidates for the new lower end
      const int first = above ? __ffs(above) - 1 : 0, last = below ? 31 - __clz(below) : 0;
      const double x_hi = __shfl_sync(0xffffffffu, xq, first), x_lo = __shfl_sync(0xffffffffu, xq, last);
      const double nh = above ? x_hi : h, nl = below ? x_lo : l;
      if (!(nl < nh) || !(nh - nl < width)) break;  // (a non-monotonic count or no progress: keep the current bracket)
      l = nl;
      h = nh;
    }
    const double lam = 0.5 * (l + h);
    if (lane == 0) {
      if (lam > 0.0) part += sqrt(lam);
      else negmax = fmax(negmax, -lam);
    }
  }
  part = block_sum(part, red);
  if (threadIdx.x == 0 && part != 0.0) atomicAdd(scal + FR_TRSQRT, part);
  // max of non-negative doubles == max of their bit patterns
  negmax = fmax(negmax, 0.0);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) negmax = fmax(negmax, __shfl_xor_sync(0xffffffffu, negmax, o));
  if ((threadIdx.x & 31) == 0 && negmax > 0.0)
    atomicMax((unsigned long long*)(scal + FR_NEGMAX), (unsigned long long)__double_as_longlong(negmax));
}

struct FrWs {
  size_t off_L, off_S2, off_T, off_M, off_v, off_p, off_diag, off_off, off_scal, total;
};
static FrWs fr_layout(int d) {
  FrWs w;
  size_t o = 0;
  const size_t mat = align_up((size_t)d * d * 8, 256), vec = align_up((size_t)d * 8, 256);
  w.off_L = o; o += mat;
  w.off_S2 = o; o += mat;
  w.off_T = o; o += mat;
  w.off_M = o; o += mat;
  w.off_v = o; o += 8 * vec;  // (TRI_MAX_CHUNKS chunk partials of p, double-buffered: v and p)
  w.off_p = o; o += 8 * vec;
  w.off_diag = o; o += vec;
  w.off_off = o; o += vec;
  w.off_scal = o; o += 256;
  w.total = o;
  return w;
}

static int frechet_once(tise_ctx* ctx, const double* mu1, const double* s1, const double* mu2, const double* s2, int d,
                        double eps_add, double delta_rel, char* ws, cudaStream_t st, double* h_scal) {
  const FrWs w = fr_layout(d);
  double* L = (double*)(ws + w.off_L);
  double* S2 = (double*)(ws + w.off_S2);
  double* T = (double*)(ws + w.off_T);
  double* M = (double*)(ws + w.off_M);
  double* vbuf = (double*)(ws + w.off_v);
  double* pbuf = (double*)(ws + w.off_p);
  double* diag = (double*)(ws + w.off_diag);
  double* off = (double*)(ws + w.off_off);
  double* scal = (double*)(ws + w.off_scal);
  const unsigned nel_blocks = (unsigned)(((size_t)d * d + 255) / 256);

  fr_scalars_kernel<<<1, 1024, 0, st>>>(mu1, s1, mu2, s2, d, scal);
  TISE_LAUNCH_CHECK(ctx);
  fr_copy_shift_kernel<<<nel_blocks, 256, 0, st>>>(s1, L, d, eps_add, delta_rel, scal, FR_TR1);
  TISE_LAUNCH_CHECK(ctx);
  fr_copy_shift_kernel<<<nel_blocks, 256, 0, st>>>(s2, S2, d, eps_add, 0.0, scal, FR_TR2);
  TISE_LAUNCH_CHECK(ctx);

  // 1. Cholesky
  for (int k = 0; k < d; k += CNB) {
    const int nb = d - k < CNB ? d - k : CNB;
    chol_diag_kernel<<<1, dim3(CNB, CNB), 0, st>>>(L, d, k, nb, delta_rel, scal);
    TISE_LAUNCH_CHECK(ctx);
    const int rest = d - k - nb;
    if (rest > 0) {
      chol_trsm_kernel<<<(rest + 127) / 128, 128, 0, st>>>(L, d, k, nb);
      TISE_LAUNCH_CHECK(ctx);
      const int tiles = (rest + 63) / 64;
      chol_syrk_kernel<<<dim3(tiles, tiles), 256, 0, st>>>(L, d, k, nb);
      TISE_LAUNCH_CHECK(ctx);
    }
  }
  zero_upper_kernel<<<nel_blocks, 256, 0, st>>>(L, d);
  TISE_LAUNCH_CHECK(ctx);

  // 2. M = L^T (S2 L)
  const dim3 gg((d + 63) / 64, (d + 63) / 64);
  dgemm_kernel<false><<<gg, 256, 0, st>>>(S2, L, T, d);
  TISE_LAUNCH_CHECK(ctx);
  dgemm_kernel<true><<<gg, 256, 0, st>>>(L, T, M, d);
  TISE_LAUNCH_CHECK(ctx);
  symmetrize_kernel<<<nel_blocks, 256, 0, st>>>(M, d);
  TISE_LAUNCH_CHECK(ctx);

  // 3. tridiagonalise
  static int legacy = -1;
  if (legacy < 0) legacy = getenv("TISE_FRECHET_LEGACY") ? atoi(getenv("TISE_FRECHET_LEGACY")) : 0;
  if (!legacy && d > 2) {  // one cooperative launch for all columns
    unsigned* bar = (unsigned*)(scal + FR_NSLOTS);
    TISE_CUDA(ctx, cudaMemsetAsync(bar, 0, sizeof(unsigned), st));
    const size_t smem = (size_t)d * 32;
    int per_sm = 0;
    TISE_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tri_persistent_kernel, TRI_THREADS, smem));
    if (per_sm < 1) return fail(ctx, TISE_ERR_CUDA, "tise_frechet: tri_persistent_kernel does not fit on an SM");
    static int want = -1;
    if (want < 0) want = getenv("TISE_FRECHET_CTAS_PER_SM") ? atoi(getenv("TISE_FRECHET_CTAS_PER_SM")) : 2;
    int grid = (per_sm > want ? (want > 0 ? want : 1) : per_sm) * ctx->num_sms;
    int dd = d;
    static int dbg = -1;
    if (dbg < 0) dbg = getenv("TISE_FRECHET_DEBUG") ? atoi(getenv("TISE_FRECHET_DEBUG")) : 0;
    void* args[] = {(void*)&M, (void*)&dd, (void*)&pbuf, (void*)&vbuf, (void*)&diag, (void*)&off, (void*)&bar, (void*)&dbg};
    TISE_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)tri_persistent_kernel, dim3(grid), dim3(TRI_THREADS), args, smem, st));
    ctx->launches++;
  } else {
    for (int k = 0; k + 2 < d; ++k) {
      const int m = d - k - 1;
      int g1 = (m + 7) / 8;
      if (g1 > 2 * ctx->num_sms) g1 = 2 * ctx->num_sms;
      tri_reflect_symv_kernel<<<g1, TRI_THREADS, (size_t)m * 8, st>>>(M, d, k, vbuf, pbuf, diag, off, scal);
      TISE_LAUNCH_CHECK(ctx);
      int g2 = m < 2 * ctx->num_sms ? m : 2 * ctx->num_sms;
      tri_rank2_kernel<<<g2, TRI_THREADS, (size_t)m * 16, st>>>(M, d, k, vbuf, pbuf, scal);
      TISE_LAUNCH_CHECK(ctx);
    }
  }
  if (legacy || d <= 2) {  // (the persistent kernel publishes the 2 x 2 tail itself)
    tri_tail_kernel<<<1, 32, 0, st>>>(M, d, diag, off);
    TISE_LAUNCH_CHECK(ctx);
  }

  // 4. eigenvalues -> sum of square roots
  sturm_sqrt_trace_kernel<<<(d + 7) / 8, 256, (size_t)d * 16, st>>>(diag, off, d, scal);  // one warp per eigenvalue
  TISE_LAUNCH_CHECK(ctx);

  TISE_CUDA(ctx, cudaMemcpyAsync(h_scal, scal, FR_NSLOTS * sizeof(double), cudaMemcpyDeviceToHost, st));
  TISE_CUDA(ctx, cudaStreamSynchronize(st));
  return TISE_OK;
}

}  // namespace tise

using namespace tise;

extern "C" size_t tise_frechet_workspace_bytes(int d) { return d > 0 ? fr_layout(d).total : 0; }

extern "C" int tise_frechet(tise_ctx* ctx, const double* d_mu1, const double* d_sigma1, const double* d_mu2,
                            const double* d_sigma2, int d, double eps, double* h_fid, int* h_flags, void* d_ws,
                            size_t ws_bytes, tise_stream stream) {
  if (!ctx) return TISE_ERR_INVALID;
  if (!d_mu1 || !d_sigma1 || !d_mu2 || !d_sigma2 || !h_fid || !d_ws || d <= 0)
    return fail(ctx, TISE_ERR_INVALID, "tise_frechet: null argument");
  if (d > 4096) return fail(ctx, TISE_ERR_INVALID, "tise_frechet: d = %d > 4096", d);
  if (ws_bytes < fr_layout(d).total)
    return fail(ctx, TISE_ERR_WORKSPACE, "tise_frechet: workspace %zu < %zu bytes", ws_bytes, fr_layout(d).total);
  cudaStream_t st = (cudaStream_t)stream;
  static PerDeviceFlag configured;
  if (!configured.is(ctx, 1)) {
    TISE_CUDA(ctx, cudaFuncSetAttribute(tri_rank2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 16));
    TISE_CUDA(ctx, cudaFuncSetAttribute(tri_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 32));
    TISE_CUDA(ctx, cudaFuncSetAttribute(sturm_sqrt_trace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 16));
    configured.set(ctx, 1);
  }
  double scal[FR_NSLOTS];
  int flags = 0;
  double fid = 0.0;
  // jitter delta = delta_rel * tr(s1): start tiny (bias of FID(X, X) ~ d * delta) and grow only if a Cholesky pivot
  // drowned in rounding noise (heavily rank-deficient s1, e.g. 16 images in 2048 dimensions)
  double eps_add = 0.0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    for (double delta_rel = 1e-14; delta_rel <= 1e-7; delta_rel *= 100.0) {
      int rc = frechet_once(ctx, d_mu1, d_sigma1, d_mu2, d_sigma2, d, eps_add, delta_rel, (char*)d_ws, st, scal);
      if (rc) return rc;
      if (scal[FR_BADPIV] == 0.0) break;
    }
    fid = scal[FR_DIFF2] + scal[FR_TR1] + scal[FR_TR2] - 2.0 * scal[FR_TRSQRT];
    if (isfinite(fid)) break;
    // fid_score.py:156-160: singular product -> add eps to the diagonal of both covariances and redo
    flags |= 1;
    eps_add = eps;
  }
  // fid_score.py:163-166: an imaginary part above 1e-3 is an error; here that is sqrt(-lambda_min) of M
  if (scal[FR_NEGMAX] > 0.0 && sqrt(scal[FR_NEGMAX]) > 1e-3) flags |= 2;
  *h_fid = fid;
  if (h_flags) *h_flags = flags;
  if (!isfinite(fid)) return fail(ctx, TISE_ERR_NUMERIC, "tise_frechet: non-finite result");
  return TISE_OK;
}
