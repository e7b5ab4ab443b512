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
#include <stdio.h>
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
// On the fp64 TENSOR pipe (mma.sync.m8n8k4.f64 -> DMMA): 11.8 TFLOP/s against 7.6 for the FMA version of this kernel.  One CTA owns a 128 x 128 tile of C: 16 warps, each a 32 x 32 sub-tile
// = 4 x 4 DMMA tiles; op(A) and B stream through shared memory 16 k at a time, register-prefetched one chunk ahead.
// Fragments of m8n8k4: lane T holds A[T / 4][T % 4], B[T % 4][T / 4], C[T / 4][2 (T % 4) + {0, 1}]; both operands are
// kept k-major in shared memory ([k][m] and [k][n], row pitch 132 doubles: conflict-free fragment reads).
constexpr int GT = 128, GKC = 16, GPITCH = GT + 4;

__device__ __forceinline__ void fr_dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

template <bool TA>
__global__ void __launch_bounds__(512) dgemm_kernel(const double* __restrict__ A, const double* __restrict__ B,
                                                    double* __restrict__ Cm, int d) {
  __shared__ double sa[GKC][GPITCH];
  __shared__ double sb[GKC][GPITCH];
  const int m0 = blockIdx.y * GT, n0 = blockIdx.x * GT;
  // Both products have a lower-triangular factor (L[k][j] = 0 for k < j, exact zeros from zero_upper_kernel):
  //   TA = false: T = S2 L      -> only k >= n0 contributes to column block n0        (half of the k range on average)
  //   TA = true:  M = L^T T     -> only k >= m0 contributes to row block m0, and M is symmetric: tiles above the
  //                                diagonal are not computed at all (symmetrize_kernel fills them)   (a quarter)
  if (TA && n0 > m0) return;
  const int k_begin = TA ? m0 : n0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wa = (warp >> 2) * 32, wb = (warp & 3) * 32;
  const int fr = lane >> 2, fk = lane & 3;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  // loaders.  [k][x] operands (B always, op(A) = A^T when TA): thread -> (k = tid / 32, 4 consecutive x = 4 (tid % 32)).
  // A not transposed ([m][k] in memory): thread -> (m = tid % 128, 4 consecutive k = 4 (tid / 128)), stored transposed.
  const int lk = threadIdx.x >> 5, lx = (threadIdx.x & 31) * 4;
  const int tm = threadIdx.x & 127, tk = (threadIdx.x >> 7) * 4;
  double ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int u = 0; u < 4; ++u) ra[u] = rb[u] = 0.0;
    if (k0 + lk < d) {
      if (n0 + lx + 3 < d) {
        const double2 v0 = __ldg(reinterpret_cast<const double2*>(B + (size_t)(k0 + lk) * d + n0 + lx));
        const double2 v1 = __ldg(reinterpret_cast<const double2*>(B + (size_t)(k0 + lk) * d + n0 + lx + 2));
        rb[0] = v0.x; rb[1] = v0.y; rb[2] = v1.x; rb[3] = v1.y;
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (n0 + lx + u < d) rb[u] = __ldg(B + (size_t)(k0 + lk) * d + n0 + lx + u);
      }
    }
    if (TA) {
      if (k0 + lk < d) {
        if (m0 + lx + 3 < d) {
          const double2 v0 = __ldg(reinterpret_cast<const double2*>(A + (size_t)(k0 + lk) * d + m0 + lx));
          const double2 v1 = __ldg(reinterpret_cast<const double2*>(A + (size_t)(k0 + lk) * d + m0 + lx + 2));
          ra[0] = v0.x; ra[1] = v0.y; ra[2] = v1.x; ra[3] = v1.y;
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (m0 + lx + u < d) ra[u] = __ldg(A + (size_t)(k0 + lk) * d + m0 + lx + u);
        }
      }
    } else if (m0 + tm < d) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (k0 + tk + u < d) ra[u] = __ldg(A + (size_t)(m0 + tm) * d + k0 + tk + u);
    }
  };
  fetch(k_begin);
  for (int k0 = k_begin; k0 < d; k0 += GKC) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      sb[lk][lx + u] = rb[u];
      if (TA) sa[lk][lx + u] = ra[u];
      else sa[tk + u][tm] = ra[u];
    }
    __syncthreads();
    if (k0 + GKC < d) fetch(k0 + GKC);
#pragma unroll
    for (int kk = 0; kk < GKC; kk += 4) {
      double af[4], bf[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = sa[kk + fk][wa + 8 * i + fr];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = sb[kk + fk][wb + 8 * j + fr];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) fr_dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = m0 + wa + 8 * i + fr, c = n0 + wb + 8 * j + 2 * fk;
      if (r < d && c < d) Cm[(size_t)r * d + c] = acc[i][j][0];
      if (r < d && c + 1 < d) Cm[(size_t)r * d + c + 1] = acc[i][j][1];
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

// ---------------------------------------------------------------------------- on-chip tridiagonalisation
// The persistent kernel above is bound by synchronisation and L2 round trips (~9.7 us per column at d = 2048: grid
// barrier 1.9 us + ~4 dependent L2 round trips for the pass + the preamble, profiles/frechet_launches_r01.txt), not
// by bandwidth or arithmetic.  This kernel keeps the WHOLE symmetric matrix on chip for the whole factorisation:
//   * row i lives in CTA i % G (cyclic, so the shrinking trailing matrix stays balanced), as local row i / G: the
//     first TR_RS local rows in shared memory, the last TR_RR (the longest-lived ones) in registers; thread t owns the
//     columns t + 512 c of every local row.  148 CTAs x 14 rows x 2048 columns x 8 B = 32 MB of shared memory + registers;
//   * v, w and the next reflector are column-distributed the same way (thread t holds entries t + 512 c in registers);
//     the few row-indexed entries a CTA needs (its own <= 14 rows) are passed through tiny shared arrays;
//   * per column the CTAs exchange exactly two vectors through L2: y = A v_next and column k + 2 of the updated matrix
//     (= the next row to finalise, by symmetry), each CTA contributing the entries of its own rows.  There is no grid
//     barrier and no fence: every 8-byte value travels as two (32-bit half, epoch) pairs in one 16-byte volatile
//     access ("LL" words), so the data carries its own flag and a reader spins on the values it needs, requesting
//     only what has not arrived yet.  Two alternating buffers suffice (a CTA can be at most one column ahead of the
//     slowest reader of its data).
// Per column: one L2 write -> read latency, two block reductions (p.v and the reflector norm), the fused rank-2 update +
// matrix-vector pass over on-chip data, and a 14-value transpose reduction.  Measured on B200, d = 2048
// (TISE_FRECHET_DEBUG=64, cycles per column): exchange ~7 k, reflector 1.9 k, pass 3.3 k (fp64 FMA rate), y reduction 1 k.
constexpr int TR_THREADS = 512;
constexpr int TR_CS = 4;                        // column slots per thread: j = t + 512 c
constexpr int TR_PITCH = TR_THREADS * TR_CS;    // d <= 2048
constexpr int TR_RS = 10;                       // local rows in shared memory
constexpr int TR_RR = 4;                        // local rows in registers
constexpr int TR_ROWS = TR_RS + TR_RR;

__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// "LL" words: an 8-byte value travels as two (32-bit half, epoch) pairs in one 16-byte access, so the data carries its own
// flag -- no fence, no separate flag round trip; a reader spins on the value itself.
__device__ __forceinline__ void ll_store(uint4* p, double v, unsigned epoch) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned)b), "r"(epoch),
               "r"((unsigned)(b >> 32)), "r"(epoch)
               : "memory");
}
__device__ __forceinline__ bool ll_load(const uint4* p, unsigned epoch, double& v) {
  unsigned a, fa, b, fb;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(fa), "=r"(b), "=r"(fb) : "l"(p) : "memory");
  v = __longlong_as_double((long long)(((unsigned long long)b << 32) | (unsigned long long)a));
  return fa == epoch && fb == epoch;
}
// sum over the 512 threads of the CTA with ONE __syncthreads; every thread gets the same bits
__device__ __forceinline__ double block_sum512(double v, double* red16) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red16[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = red16[threadIdx.x & 15];
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}

// The two passes over a CTA's rows, specialised on the first live column slot CF (= r0 / 512, uniform): no per-slot
// predicates inside, ~12 instructions per row -- the passes were bound by instruction issue, not by the fp64 pipe.
template <int CF>
__device__ __forceinline__ void tr_matvec(const double* __restrict__ S, const double (&areg)[TR_RR][TR_CS],
                                          const double (&vj)[TR_CS], const double (&wj)[TR_CS], const double (&vnj)[TR_CS],
                                          double (&acc)[16], const double* s_v, const double* s_w, double* s_col, int t, int G,
                                          int cta, int r0, int d) {
  const bool own = t == r0 - CF * TR_THREADS;  // this thread holds column r0 (slot CF)
#pragma unroll
  for (int l = 0; l < TR_ROWS; ++l) {
    const int i = l * G + cta;
    if (i >= r0 && i < d) {  // (uniform across the CTA)
      double s0 = 0.0, s1 = 0.0, a_cf = 0.0;
#pragma unroll
      for (int c = CF; c < TR_CS; ++c) {
        const double a = l < TR_RS ? S[l * TR_PITCH + t + c * TR_THREADS] : areg[l < TR_RS ? 0 : l - TR_RS][c];
        if (c == CF) a_cf = a;
        if ((c - CF) & 1) s1 = fma(a, vnj[c], s1);
        else s0 = fma(a, vnj[c], s0);
      }
      acc[l] = s0 + s1;
      // column r0 of this row after the update (= entry i of the next row to finalise): travels with y_i
      if (own) s_col[l] = fma(-s_w[l], vj[CF], fma(-s_v[l], wj[CF], a_cf));
    }
  }
}
template <int CF>
__device__ __forceinline__ void tr_update(double* __restrict__ S, double (&areg)[TR_RR][TR_CS], const double (&vj)[TR_CS],
                                          const double (&wj)[TR_CS], const double* s_v, const double* s_w, int t, int G, int cta,
                                          int r0, int d) {
#pragma unroll
  for (int l = 0; l < TR_ROWS; ++l) {
    const int i = l * G + cta;
    if (i >= r0 && i < d) {
      const double vi = -s_v[l], wi = -s_w[l];
#pragma unroll
      for (int c = CF; c < TR_CS; ++c) {
        if (l < TR_RS) {
          double* ap = S + l * TR_PITCH + t + c * TR_THREADS;
          *ap = fma(wi, vj[c], fma(vi, wj[c], *ap));
        } else {
          double& ar = areg[l < TR_RS ? 0 : l - TR_RS][c];
          ar = fma(wi, vj[c], fma(vi, wj[c], ar));
        }
      }
    }
  }
}

template <bool kDbg>
__global__ void __launch_bounds__(TR_THREADS, 1)
tri_resident_kernel(const double* __restrict__ M, int d, uint4* xbuf,
                    double* __restrict__ diag, double* __restrict__ off, unsigned long long* dbg, int abl) {
  // abl (kDbg only; results are garbage, timing ablations): 1 no pass, 2 no flag waits, 4 no loads of y / row,
  // 8 relaxed flag stores, 16 no y reduction
  // kDbg: thread 0 of every CTA accumulates the cycles of the five phases of a column (TISE_FRECHET_DEBUG=64)
  unsigned long long ph[6] = {0, 0, 0, 0, 0, 0};
  long long tp = 0;
#define TR_MARK(idx)                                   \
  if (kDbg && threadIdx.x == 0) {                      \
    const long long now = clock64();                   \
    ph[idx] += (unsigned long long)(now - tp);         \
    tp = now;                                          \
  }
  extern __shared__ double S[];  // [TR_RS][TR_PITCH]
  __shared__ double red_a[16], red_b[16], red_y[16][17];
  __shared__ double s_v[2][16], s_w[16], s_bc[4], s_col[16];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int G = gridDim.x, cta = blockIdx.x;
  int own_l[TR_CS];  // local row whose index equals this thread's column (the row-indexed copy of v / w it must publish)
#pragma unroll
  for (int c = 0; c < TR_CS; ++c) {
    const int j = t + c * TR_THREADS;
    own_l[c] = (j < d && j % G == cta) ? j / G : -1;
  }
  // ---- the matrix goes on chip
#pragma unroll
  for (int l = 0; l < TR_RS; ++l) {
    const int i = l * G + cta;
#pragma unroll
    for (int c = 0; c < TR_CS; ++c) {
      const int j = t + c * TR_THREADS;
      S[l * TR_PITCH + j] = (i < d && j < d) ? M[(size_t)i * d + j] : 0.0;
    }
  }
  double areg[TR_RR][TR_CS];
#pragma unroll
  for (int l = 0; l < TR_RR; ++l) {
    const int i = (TR_RS + l) * G + cta;
#pragma unroll
    for (int c = 0; c < TR_CS; ++c) {
      const int j = t + c * TR_THREADS;
      areg[l][c] = (i < d && j < d) ? M[(size_t)i * d + j] : 0.0;
    }
  }
  double vj[TR_CS], wj[TR_CS], vnj[TR_CS];
#pragma unroll
  for (int c = 0; c < TR_CS; ++c) vj[c] = wj[c] = vnj[c] = 0.0;
  double tau = 0.0;
  if (t < 16) {
    s_v[0][t] = 0.0;
    s_v[1][t] = 0.0;
    s_w[t] = 0.0;
  }
  if (t < 4) s_bc[t] = 0.0;
  __syncthreads();
  int cur = 0;
  // step k: finalise row q = k + 1 (its diagonal and sub-diagonal), build the reflector of column q, apply the rank-2
  // update of column k to rows / columns >= q + 1 and accumulate y = A v_q there.  k = -1 is the start (no update).
  if (kDbg) tp = clock64();
  for (int k = -1; k <= d - 3; ++k) {
    const int q = k + 1, r0 = q + 1;
    double yj[TR_CS], rj[TR_CS];
    if (k < 0) {
#pragma unroll
      for (int c = 0; c < TR_CS; ++c) {
        const int j = t + c * TR_THREADS;
        yj[c] = 0.0;
        rj[c] = j < d ? M[j] : 0.0;
      }
    } else {
      const unsigned epoch = (unsigned)(k + 1);
      const uint4* xb = xbuf + (size_t)(k & 1) * 2 * d;  // word 2 j: y_j, word 2 j + 1: A[j][q] (= row q by symmetry)
      // spin on the values themselves; a word that has arrived is not requested again
      unsigned need = 0;
#pragma unroll
      for (int c = 0; c < TR_CS; ++c) {
        const int j = t + c * TR_THREADS;
        yj[c] = rj[c] = 0.0;
        if (j >= q && j < d && !(kDbg && (abl & 4))) need |= 3u << (2 * c);
      }
      while (need) {  // (all outstanding words are requested before the first one is examined)
        unsigned got = (kDbg && (abl & 2)) ? need : 0u;
#pragma unroll
        for (int c = 0; c < TR_CS; ++c) {
          const int j = t + c * TR_THREADS;
          if (need & (1u << (2 * c)))
            if (ll_load(xb + 2 * j, epoch, yj[c])) got |= 1u << (2 * c);
          if (need & (2u << (2 * c)))
            if (ll_load(xb + 2 * j + 1, epoch, rj[c])) got |= 2u << (2 * c);
        }
        need &= ~got;
        if (kDbg && (abl & 32) && need) __nanosleep(100);
      }
      TR_MARK(0);
    }
    // w = p - (tau / 2)(p . v) v with p = tau y      (v is zero outside its support)
    double dot = 0.0;
#pragma unroll
    for (int c = 0; c < TR_CS; ++c) {
      yj[c] *= tau;
      dot = fma(vj[c], yj[c], dot);
    }
    dot = block_sum512(dot, red_a);
    TR_MARK(1);
    const double cc = 0.5 * tau * dot;
#pragma unroll
    for (int c = 0; c < TR_CS; ++c) {
      wj[c] = yj[c] - cc * vj[c];
      if (own_l[c] >= 0) s_w[own_l[c]] = wj[c];
      if (t + c * TR_THREADS == q) {
        s_bc[0] = vj[c];
        s_bc[1] = wj[c];
      }
    }
    __syncthreads();
    // row q after the update of column k; its reflector
    const double v0 = s_bc[0], w0 = s_bc[1];
    double ss = 0.0;
#pragma unroll
    for (int c = 0; c < TR_CS; ++c) {
      const int j = t + c * TR_THREADS;
      rj[c] -= v0 * wj[c] + w0 * vj[c];
      if (j >= q + 2 && j < d) ss = fma(rj[c], rj[c], ss);
      if (j == q + 1) s_bc[2] = rj[c];
      if (j == q) s_bc[3] = rj[c];
    }
    ss = block_sum512(ss, red_b);
    const double alpha = s_bc[2];
    double tau_n, beta, scale;
    if (ss == 0.0) {
      tau_n = 0.0;
      beta = alpha;
      scale = 0.0;
    } else {
      beta = -copysign(sqrt(alpha * alpha + ss), alpha);
      tau_n = (beta - alpha) / beta;
      scale = 1.0 / (alpha - beta);
    }
    if (cta == 0 && t == 0) {
      diag[q] = s_bc[3];
      off[q] = beta;
    }
#pragma unroll
    for (int c = 0; c < TR_CS; ++c) {
      const int j = t + c * TR_THREADS;
      vnj[c] = (j == r0) ? 1.0 : ((j > r0 && j < d) ? rj[c] * scale : 0.0);
      if (own_l[c] >= 0) s_v[cur ^ 1][own_l[c]] = vnj[c];
    }
    TR_MARK(2);
    // ---- y = (A - v w^T - w v^T) v_next = A v_next - v (w . v_next) - w (v . v_next): the matrix-vector product runs on
    // the matrix as it stands and y is published at once; the rank-2 update itself (two thirds of the arithmetic) is
    // applied AFTER the publication, while the other CTAs' contributions are in flight (LAPACK dlatrd's delayed update).
    const bool last = k == d - 3;
    const unsigned epoch_n = (unsigned)(k + 2);
    uint4* xb_n = xbuf + (size_t)((k + 1) & 1) * 2 * d;
    double acc[16];
#pragma unroll
    for (int l = 0; l < 14; ++l) acc[l] = 0.0;
    {
      double s14 = 0.0, s15 = 0.0;  // the two correction dot products ride along in the unused reduction slots
#pragma unroll
      for (int c = 0; c < TR_CS; ++c) {
        s14 = fma(wj[c], vnj[c], s14);
        s15 = fma(vj[c], vnj[c], s15);
      }
      acc[14] = s14;
      acc[15] = s15;
    }
    // No per-element predicates: columns left of r0 are dead storage (the finished part of a row is never read again)
    // and v_next is zero there, columns >= d hold zeros; a column slot is skipped once ALL of its columns are dead.
    const int c_first = r0 / TR_THREADS;  // first column slot with live columns (uniform)
    if (!(kDbg && (abl & 1))) {
      switch (c_first) {
        case 0: tr_matvec<0>(S, areg, vj, wj, vnj, acc, s_v[cur], s_w, s_col, t, G, cta, r0, d); break;
        case 1: tr_matvec<1>(S, areg, vj, wj, vnj, acc, s_v[cur], s_w, s_col, t, G, cta, r0, d); break;
        case 2: tr_matvec<2>(S, areg, vj, wj, vnj, acc, s_v[cur], s_w, s_col, t, G, cta, r0, d); break;
        default: tr_matvec<3>(S, areg, vj, wj, vnj, acc, s_v[cur], s_w, s_col, t, G, cta, r0, d); break;
      }
    }
    TR_MARK(3);
    // ---- y: transpose reduction of the 16 sums across the warp (31 exchanges instead of 16 x 5), then across warps
    if (!(kDbg && (abl & 16)))
#pragma unroll
    for (int m = 16, n = 8; m >= 2; m >>= 1, n >>= 1) {
      const bool hi = (lane & m) != 0;
#pragma unroll
      for (int r = 0; r < n; ++r) {
        const double send = hi ? acc[r] : acc[r + n];
        const double keep = hi ? acc[r + n] : acc[r];
        acc[r] = keep + __shfl_xor_sync(0xffffffffu, send, m);
      }
    }
    acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], 1);  // lanes 2 r and 2 r + 1 now hold the warp's sum of local row r
    if ((lane & 1) == 0) red_y[lane >> 1][warp] = acc[0];
    __syncthreads();
    if (warp == 0 && !last) {
      const int row = lane & 15, h8 = (lane >> 4) * 8;
      double s = 0.0;
#pragma unroll
      for (int u = 0; u < 8; ++u) s += red_y[row][h8 + u];
      s += __shfl_xor_sync(0xffffffffu, s, 16);
      const double w_vn = __shfl_sync(0xffffffffu, s, 14), v_vn = __shfl_sync(0xffffffffu, s, 15);
      const int i = row * G + cta;
      if (lane < TR_ROWS && i >= r0 && i < d) {
        s = fma(-s_w[lane], v_vn, fma(-s_v[cur][lane], w_vn, s));
        ll_store(xb_n + 2 * i, s, epoch_n);
        ll_store(xb_n + 2 * i + 1, s_col[lane], epoch_n);
      }
    }
    if (last && t == 0 && (d - 1) % G == cta) diag[d - 1] = s_col[(d - 1) / G];
    // ---- the delayed rank-2 update (off the critical path: the next column's values are still in flight)
    if (!(kDbg && (abl & 1))) {
      switch (c_first) {
        case 0: tr_update<0>(S, areg, vj, wj, s_v[cur], s_w, t, G, cta, r0, d); break;
        case 1: tr_update<1>(S, areg, vj, wj, s_v[cur], s_w, t, G, cta, r0, d); break;
        case 2: tr_update<2>(S, areg, vj, wj, s_v[cur], s_w, t, G, cta, r0, d); break;
        default: tr_update<3>(S, areg, vj, wj, s_v[cur], s_w, t, G, cta, r0, d); break;
      }
    }
    TR_MARK(4);
#pragma unroll
    for (int c = 0; c < TR_CS; ++c) vj[c] = vnj[c];
    tau = tau_n;
    cur ^= 1;
  }
  if (kDbg && threadIdx.x == 0)
    for (int u = 0; u < 5; ++u) dbg[blockIdx.x * 8 + u] = ph[u];
#undef TR_MARK
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
  const dim3 gg((d + GT - 1) / GT, (d + GT - 1) / GT);
  dgemm_kernel<false><<<gg, 512, 0, st>>>(S2, L, T, d);
  TISE_LAUNCH_CHECK(ctx);
  dgemm_kernel<true><<<gg, 512, 0, st>>>(L, T, M, d);
  TISE_LAUNCH_CHECK(ctx);
  symmetrize_kernel<<<nel_blocks, 256, 0, st>>>(M, d);
  TISE_LAUNCH_CHECK(ctx);

  // 3. tridiagonalise
  static int legacy = -1;
  if (legacy < 0) legacy = getenv("TISE_FRECHET_LEGACY") ? atoi(getenv("TISE_FRECHET_LEGACY")) : 0;
  static int resident = -1;
  if (resident < 0) resident = getenv("TISE_FRECHET_RESIDENT") ? atoi(getenv("TISE_FRECHET_RESIDENT")) : 1;
  bool use_resident = !legacy && resident && d >= (resident > 1 ? 3 : 512) && d <= TR_PITCH && d <= TR_ROWS * ctx->num_sms;
  if (use_resident) {  // every CTA of the kernel must be co-resident (it spins on its peers' values): ask the runtime first
    int per_sm = 0;
    const void* probe = (const void*)tri_resident_kernel<false>;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, probe, TR_THREADS, (size_t)TR_RS * TR_PITCH * sizeof(double)) !=
            cudaSuccess || per_sm < 1) {
      cudaGetLastError();
      use_resident = false;  // (a device with less shared memory per SM: the L2-resident kernel below)
    }
  }
  if (use_resident) {
    // the whole matrix in shared memory + registers of one cooperative launch (one CTA per SM), see tri_resident_kernel
    uint4* xbuf = (uint4*)vbuf;    // 2 buffers x d x {y, column} LL words = 64 d bytes (one vector area of the workspace)
    TISE_CUDA(ctx, cudaMemsetAsync(xbuf, 0, (size_t)4 * d * sizeof(uint4), st));  // epochs restart at 1 in every call
    int dd = d;
    static int dbg = -1;
    if (dbg < 0) dbg = getenv("TISE_FRECHET_DEBUG") ? atoi(getenv("TISE_FRECHET_DEBUG")) : 0;
    unsigned long long* d_dbg = nullptr;
    int abl = getenv("TISE_FRECHET_ABL") ? atoi(getenv("TISE_FRECHET_ABL")) : 0;
    if (dbg & 64) TISE_CUDA(ctx, cudaMalloc(&d_dbg, (size_t)ctx->num_sms * 8 * sizeof(unsigned long long)));
    void* args[] = {(void*)&M, (void*)&dd, (void*)&xbuf, (void*)&diag, (void*)&off, (void*)&d_dbg, (void*)&abl};
    const void* fn = (dbg & 64) ? (const void*)tri_resident_kernel<true> : (const void*)tri_resident_kernel<false>;
    const cudaError_t le = cudaLaunchCooperativeKernel(fn, dim3(ctx->num_sms), dim3(TR_THREADS), args,
                                                       (size_t)TR_RS * TR_PITCH * sizeof(double), st);
    if (le != cudaSuccess) {  // e.g. SMs reserved by another client: cooperative launch too large -> L2-resident kernel
      cudaGetLastError();
      if (d_dbg) cudaFree(d_dbg);
      use_resident = false;
    } else if (dbg & 64) {  // cycles per column of the five phases, averaged over the CTAs
      std::vector<unsigned long long> h((size_t)ctx->num_sms * 8);
      TISE_CUDA(ctx, cudaStreamSynchronize(st));
      TISE_CUDA(ctx, cudaMemcpy(h.data(), d_dbg, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
      cudaFree(d_dbg);
      double sum[5] = {0, 0, 0, 0, 0};
      for (int b = 0; b < ctx->num_sms; ++b)
        for (int u = 0; u < 5; ++u) sum[u] += (double)h[(size_t)b * 8 + u];
      const double den = (double)ctx->num_sms * (d - 1);
      fprintf(stderr, "tri_resident d=%d cycles/column: wait flags %.0f | load y,row + p.v %.0f | w, row, reflector %.0f | pass %.0f | "
                      "y reduce + publish %.0f\n", d, sum[0] / den, sum[1] / den, sum[2] / den, sum[3] / den, sum[4] / den);
    }
    if (use_resident) ctx->launches++;
  }
  if (use_resident) {
    // done
  } else if (!legacy && d > 2) {  // one cooperative launch for all columns, matrix in L2
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
  if (legacy || d <= 2) {  // (the persistent kernels publish the 2 x 2 tail themselves)
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
    TISE_CUDA(ctx, cudaFuncSetAttribute(tri_resident_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        TR_RS * TR_PITCH * (int)sizeof(double)));
    TISE_CUDA(ctx, cudaFuncSetAttribute(tri_resident_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        TR_RS * TR_PITCH * (int)sizeof(double)));
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
