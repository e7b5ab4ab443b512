// A2: Pillow-exact bilinear resize of packed uint8 HWC images to 299x299 (see include/tise_b200.h).
//
// Pillow's ImagingResample (src/libImaging/Resample.c) is a separable triangle filter whose
// weights are computed in double, normalised, quantised to 22-bit fixed point, and applied with
// int32 accumulation -- horizontal pass first into a uint8 temporary, then vertical.  The
// coefficient kernel below repeats the double arithmetic operation by operation with explicit
// round-to-nearest intrinsics (no FMA contraction), so the integer tables -- and therefore every
// output byte -- are identical to Pillow's.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"

namespace tise {

constexpr int kPrecisionBits = 32 - 8 - 2;

__host__ __device__ inline int resize_ksize(int in_size) {
  // ksize = ceil(support) * 2 + 1, support = max(in/out, 1)
  int c = (in_size + kOutSize - 1) / kOutSize;  // ceil(in/out) for in > out
  if (c < 1) c = 1;
  return c * 2 + 1;
}

// One block per distinct input size; thread xx computes bounds/weights of output index xx.
// Table layout per size: int bounds[299][2] (xmin, count) followed by int kk[299][kmax].
__global__ void resize_coeffs_kernel(const int* __restrict__ sizes, int kmax, int* __restrict__ tabs) {
  const int in_size = sizes[blockIdx.x];
  int* bounds = tabs + (size_t)blockIdx.x * kOutSize * (2 + kmax);
  int* kk = bounds + kOutSize * 2;
  const int xx = threadIdx.x;
  if (xx >= kOutSize) return;

  const double scale = __ddiv_rn((double)in_size, (double)kOutSize);
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = filterscale;  // 1.0 * filterscale
  const double ss = __ddiv_rn(1.0, filterscale);
  const double center = __dmul_rn(__dadd_rn((double)xx, 0.5), scale);
  int xmin = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;

  double ww = 0.0;
  for (int x = 0; x < xmax; ++x) {
    double a = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
    if (a < 0.0) a = -a;
    const double w = a < 1.0 ? __dsub_rn(1.0, a) : 0.0;
    ww = __dadd_rn(ww, w);
  }
  for (int x = 0; x < kmax; ++x) {
    int q = 0;
    if (x < xmax) {
      double a = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
      if (a < 0.0) a = -a;
      double w = a < 1.0 ? __dsub_rn(1.0, a) : 0.0;
      if (ww != 0.0) w = __ddiv_rn(w, ww);
      const double t = __dmul_rn(w, (double)(1 << kPrecisionBits));
      q = w < 0.0 ? (int)__dadd_rn(-0.5, t) : (int)__dadd_rn(0.5, t);
    }
    kk[xx * kmax + x] = q;
  }
  bounds[xx * 2 + 0] = xmin;
  bounds[xx * 2 + 1] = xmax;
}

struct ImgDesc {
  long long src_off;  // byte offset of the image (of the crop's first pixel) in d_imgs
  long long tmp_off;  // byte offset of its [h][299][3] horizontal-pass temporary
  int h, w;           // size of the image / of the crop box
  int tab_x, tab_y;   // table indices (by input width / height)
  int pitch, pad;     // bytes between consecutive source rows (w * 3 unless this is a crop of a wider image)
};

__device__ __forceinline__ uint8_t clip8(int acc) {
  int v = acc >> kPrecisionBits;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// Horizontal pass: tmp[i][y][xx][c] for y < h.  One thread per output byte, consecutive threads
// walk (xx, c) inside a row so both the gather and the store are contiguous per warp.
// grid = (ceil(897 / 256), max_h, n)
__global__ void resize_h_kernel(const uint8_t* __restrict__ imgs, const ImgDesc* __restrict__ descs,
                                const int* __restrict__ tabs, int kmax, uint8_t* __restrict__ tmp) {
  const ImgDesc d = descs[blockIdx.z];
  const int y = blockIdx.y;
  if (y >= d.h) return;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;  // xx * 3 + c
  if (j >= kOutSize * 3) return;
  const int xx = j / 3, c = j - xx * 3;
  const uint8_t* row = imgs + d.src_off + (size_t)y * d.pitch;
  uint8_t* out = tmp + d.tmp_off + (size_t)y * kOutSize * 3;
  if (d.w == kOutSize) {  // Pillow skips a pass that does not change the size
    out[j] = row[j];
    return;
  }
  const int* bounds = tabs + (size_t)d.tab_x * kOutSize * (2 + kmax);
  const int* kk = bounds + kOutSize * 2 + xx * kmax;
  const int xmin = bounds[xx * 2], cnt = bounds[xx * 2 + 1];
  int acc = 1 << (kPrecisionBits - 1);
  for (int x = 0; x < cnt; ++x) acc += (int)row[(xmin + x) * 3 + c] * kk[x];
  out[j] = clip8(acc);
}

// Vertical pass: out[i][yy][xx][c].  grid = (ceil(897 / 256), 299, n)
__global__ void resize_v_kernel(const uint8_t* __restrict__ tmp, const ImgDesc* __restrict__ descs,
                                const int* __restrict__ tabs, int kmax, uint8_t* __restrict__ out) {
  const ImgDesc d = descs[blockIdx.z];
  const int yy = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= kOutSize * 3) return;
  const uint8_t* src = tmp + d.tmp_off;
  uint8_t* dst = out + ((size_t)blockIdx.z * kOutSize + yy) * kOutSize * 3;
  if (d.h == kOutSize) {
    dst[j] = src[(size_t)yy * kOutSize * 3 + j];
    return;
  }
  const int* bounds = tabs + (size_t)d.tab_y * kOutSize * (2 + kmax);
  const int* kk = bounds + kOutSize * 2 + yy * kmax;
  const int ymin = bounds[yy * 2], cnt = bounds[yy * 2 + 1];
  int acc = 1 << (kPrecisionBits - 1);
  for (int y = 0; y < cnt; ++y) acc += (int)src[(size_t)(ymin + y) * kOutSize * 3 + j] * kk[y];
  dst[j] = clip8(acc);
}

// Fused variant: one CTA = (image, strip of kStripRows output rows).  The input rows the strip depends on are staged in
// shared memory, resampled horizontally into a uint8 shared-memory temporary (Pillow rounds to uint8 between the
// passes) and vertically into the output strip, which is one contiguous byte range of the NHWC output.  HBM traffic is
// the algorithmic minimum (read the image once, write the 299 x 299 result once); used whenever the strip's input
// rows fit in shared memory (always for up-scaling; the two-kernel path above remains for big down-scales).
constexpr int kStripRows = 13;  // 299 = 23 * 13
constexpr int kFusedSmem = 96 * 1024;
constexpr int kTmpPitch = 900;  // bytes per row of the uint8 temporary in shared memory (897 rounded up to words)
__global__ void __launch_bounds__(256) resize_fused_kernel(const uint8_t* __restrict__ imgs, const ImgDesc* __restrict__ descs,
                                                           const int* __restrict__ tabs, int kmax,
                                                           uint8_t* __restrict__ out) {
  extern __shared__ uint8_t sm[];
  const ImgDesc d = descs[blockIdx.y];
  const int r0 = blockIdx.x * kStripRows;
  const int* bounds_y = tabs + (size_t)d.tab_y * kOutSize * (2 + kmax);
  const int* kk_y = bounds_y + kOutSize * 2;
  const int* bounds_x = tabs + (size_t)d.tab_x * kOutSize * (2 + kmax);
  const int* kk_x = bounds_x + kOutSize * 2;
  const bool vpass = d.h != kOutSize, hpass = d.w != kOutSize;
  // input rows needed by this strip
  int y_lo = r0, y_hi = r0 + kStripRows;
  if (vpass) {
    y_lo = bounds_y[r0 * 2];
    y_hi = bounds_y[(r0 + kStripRows - 1) * 2] + bounds_y[(r0 + kStripRows - 1) * 2 + 1];
  }
  const int n_in = y_hi - y_lo;
  const int in_row = d.w * 3, out_row = kOutSize * 3;
  uint8_t* s_tmp = sm;                                          // [n_in][kTmpPitch]   horizontally resampled rows
  uint8_t* s_in = sm + (((size_t)n_in * kTmpPitch + 15) & ~(size_t)15);  // [n_in][w * 3]  raw input rows
  const uint8_t* src = imgs + d.src_off + (size_t)y_lo * d.pitch;
  const int in_bytes = n_in * in_row;
  if (hpass) {
    if (d.pitch != in_row) {  // a crop box: rows are not contiguous
      for (int i = threadIdx.x; i < in_bytes; i += blockDim.x) {
        const int y = i / in_row;
        s_in[i] = __ldg(src + (size_t)y * d.pitch + (i - y * in_row));
      }
    } else if ((((uintptr_t)src | (uintptr_t)in_bytes) & 15) == 0) {
      for (int i = threadIdx.x; i < in_bytes / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(s_in)[i] = __ldg(reinterpret_cast<const uint4*>(src) + i);
    } else {
      for (int i = threadIdx.x; i < in_bytes; i += blockDim.x) s_in[i] = __ldg(src + i);
    }
    __syncthreads();
    // one thread per (row, output pixel): bounds and weights are fetched once for the 3 channels
    for (int i = threadIdx.x; i < n_in * kOutSize; i += blockDim.x) {
      const int y = i / kOutSize, xx = i - y * kOutSize;
      const int xmin = bounds_x[xx * 2], cnt = bounds_x[xx * 2 + 1];
      const int* k = kk_x + xx * kmax;
      const uint8_t* px = s_in + y * in_row + xmin * 3;
      int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
      for (int x = 0; x < cnt; ++x) {
        const int kw = k[x];
        a0 += (int)px[x * 3 + 0] * kw;
        a1 += (int)px[x * 3 + 1] * kw;
        a2 += (int)px[x * 3 + 2] * kw;
      }
      uint8_t* o = s_tmp + y * kTmpPitch + xx * 3;
      o[0] = clip8(a0);
      o[1] = clip8(a1);
      o[2] = clip8(a2);
    }
  } else {  // Pillow skips the horizontal pass: the rows are already 897 bytes wide
    for (int i = threadIdx.x; i < in_bytes; i += blockDim.x) {
      const int y = i / out_row;
      s_tmp[y * kTmpPitch + (i - y * out_row)] = __ldg(src + (size_t)y * d.pitch + (i - y * out_row));
    }
  }
  __syncthreads();
  // one thread per (output row, group of 4 consecutive bytes): one 32-bit shared-memory load per filter tap
  uint8_t* dst = out + ((size_t)blockIdx.y * kOutSize + r0) * out_row;
  constexpr int kGroups = kTmpPitch / 4;
  for (int i = threadIdx.x; i < kStripRows * kGroups; i += blockDim.x) {
    const int yy = i / kGroups, j = (i - yy * kGroups) * 4;
    uint32_t packed;
    if (!vpass) {
      packed = *reinterpret_cast<const uint32_t*>(s_tmp + yy * kTmpPitch + j);
    } else {
      const int ymin = bounds_y[(r0 + yy) * 2], cnt = bounds_y[(r0 + yy) * 2 + 1];
      const int* k = kk_y + (r0 + yy) * kmax;
      const uint8_t* col = s_tmp + (ymin - y_lo) * kTmpPitch + j;
      int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0, a3 = a0;
      for (int y = 0; y < cnt; ++y) {
        const uint32_t v = *reinterpret_cast<const uint32_t*>(col + y * kTmpPitch);
        const int kw = k[y];
        a0 += (int)(v & 0xffu) * kw;
        a1 += (int)((v >> 8) & 0xffu) * kw;
        a2 += (int)((v >> 16) & 0xffu) * kw;
        a3 += (int)(v >> 24) * kw;
      }
      packed = (uint32_t)clip8(a0) | ((uint32_t)clip8(a1) << 8) | ((uint32_t)clip8(a2) << 16) | ((uint32_t)clip8(a3) << 24);
    }
    uint8_t* o = dst + yy * out_row + j;  // rows of 897 bytes: not word-aligned in general
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if (j + b < out_row) o[b] = (uint8_t)(packed >> (8 * b));
  }
}

// Workspace layout: [descs n][sizes 2n ints][tables (2n) * 299 * (2 + kmax) ints][tmp n * max_h * 897]
struct ResizeWs {
  size_t off_desc, off_sizes, off_tabs, off_tmp, total;
};
static ResizeWs resize_ws_layout(int n, int max_h, int max_w) {
  ResizeWs L;
  const int kmax = resize_ksize(max_h > max_w ? max_h : max_w);
  size_t o = 0;
  L.off_desc = o;
  o = align_up(o + sizeof(ImgDesc) * (size_t)n, 256);
  L.off_sizes = o;
  o = align_up(o + sizeof(int) * 2 * (size_t)n, 256);
  L.off_tabs = o;
  o = align_up(o + sizeof(int) * 2 * (size_t)n * kOutSize * (2 + kmax), 256);
  L.off_tmp = o;
  o = align_up(o + (size_t)n * max_h * kOutSize * 3, 256);
  L.total = o;
  return L;
}

}  // namespace tise

using namespace tise;

extern "C" size_t tise_resize_workspace_bytes(int n, int max_h, int max_w) {
  if (n <= 0 || max_h <= 0 || max_w <= 0) return 0;
  return resize_ws_layout(n, max_h, max_w).total;
}

// h_box: NULL (whole images) or n x (x0, y0, x1, y1) crop boxes inside the images
static int resize_impl(tise_ctx* ctx, const char* who, const uint8_t* d_imgs, const int64_t* h_offsets, const int32_t* h_hw,
                       const int32_t* h_box, int uniform, int n, uint8_t* d_out, void* d_ws, size_t ws_bytes,
                       tise_stream stream) {
  if (!ctx) return TISE_ERR_INVALID;
  if (n == 0) return TISE_OK;
  if (!d_imgs || !h_hw || !d_out || !d_ws || n < 0) return fail(ctx, TISE_ERR_INVALID, "%s: null argument", who);
  if (!uniform && !h_offsets) return fail(ctx, TISE_ERR_INVALID, "%s: ragged batch needs h_offsets", who);
  cudaStream_t st = (cudaStream_t)stream;

  int max_h = 0, max_w = 0;
  for (int i = 0; i < n; ++i) {
    int h = h_hw[uniform ? 0 : 2 * i], w = h_hw[uniform ? 1 : 2 * i + 1];
    if (h <= 0 || w <= 0) return fail(ctx, TISE_ERR_INVALID, "%s: image %d has size %dx%d", who, i, h, w);
    if (h_box) {
      const int x0 = h_box[4 * i], y0 = h_box[4 * i + 1], x1 = h_box[4 * i + 2], y1 = h_box[4 * i + 3];
      if (x0 < 0 || y0 < 0 || x1 > w || y1 > h || x1 <= x0 || y1 <= y0)
        return fail(ctx, TISE_ERR_INVALID, "%s: box %d = (%d, %d, %d, %d) is empty or outside its %dx%d image", who, i, x0,
                    y0, x1, y1, h, w);
      h = y1 - y0;
      w = x1 - x0;
    }
    max_h = h > max_h ? h : max_h;
    max_w = w > max_w ? w : max_w;
    if (uniform && !h_box) break;
  }
  const ResizeWs L = resize_ws_layout(n, max_h, max_w);
  if (ws_bytes < L.total) return fail(ctx, TISE_ERR_WORKSPACE, "%s: workspace %zu < %zu bytes", who, ws_bytes, L.total);
  const int kmax = resize_ksize(max_h > max_w ? max_h : max_w);

  // distinct sizes -> table indices
  std::vector<ImgDesc> descs(n);
  std::vector<int> sizes;
  std::map<int, int> tab_of;
  auto table = [&](int s) {
    auto it = tab_of.find(s);
    if (it != tab_of.end()) return it->second;
    const int t = (int)sizes.size();
    sizes.push_back(s);
    tab_of[s] = t;
    return t;
  };
  long long tmp_off = 0;
  for (int i = 0; i < n; ++i) {
    int h = h_hw[uniform ? 0 : 2 * i], w = h_hw[uniform ? 1 : 2 * i + 1];
    ImgDesc& d = descs[i];
    d.src_off = uniform ? (long long)i * h * w * 3 : (long long)h_offsets[i];
    d.pitch = w * 3;
    d.pad = 0;
    if (h_box) {
      d.src_off += (long long)h_box[4 * i + 1] * d.pitch + (long long)h_box[4 * i] * 3;
      h = h_box[4 * i + 3] - h_box[4 * i + 1];
      w = h_box[4 * i + 2] - h_box[4 * i];
    }
    d.h = h;
    d.w = w;
    d.tmp_off = tmp_off;
    tmp_off += (long long)h * kOutSize * 3;
    d.tab_x = table(w);
    d.tab_y = table(h);
  }
  char* ws = (char*)d_ws;
  ImgDesc* d_descs = (ImgDesc*)(ws + L.off_desc);
  int* d_sizes = (int*)(ws + L.off_sizes);
  int* d_tabs = (int*)(ws + L.off_tabs);
  uint8_t* d_tmp = (uint8_t*)(ws + L.off_tmp);
  int urc = stage_upload(ctx, d_descs, descs.data(), sizeof(ImgDesc) * n, st);
  if (!urc) urc = stage_upload(ctx, d_sizes, sizes.data(), sizeof(int) * sizes.size(), st);
  if (urc) return urc;

  resize_coeffs_kernel<<<(unsigned)sizes.size(), 320, 0, st>>>(d_sizes, kmax, d_tabs);
  TISE_LAUNCH_CHECK(ctx);
  // fused single-kernel path if every strip's input rows (+ the uint8 temporary) fit in shared memory
  {
    const double vscale = max_h > kOutSize ? (double)max_h / kOutSize : 1.0;
    const int n_in_max = (int)(kStripRows * ((double)max_h / kOutSize) + 2.0 * vscale + 3.0);
    const size_t need = (((size_t)n_in_max * kTmpPitch + 15) & ~(size_t)15) + (size_t)n_in_max * max_w * 3 + 16;
    static int fused_env = -1;
    if (fused_env < 0) fused_env = getenv("TISE_RESIZE_FUSED") ? atoi(getenv("TISE_RESIZE_FUSED")) : 1;
    if (fused_env && need <= (size_t)kFusedSmem) {
      static PerDeviceFlag configured;
      if (!configured.is(ctx, 1)) {
        TISE_CUDA(ctx, cudaFuncSetAttribute(resize_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmem));
        configured.set(ctx, 1);
      }
      for (int i0 = 0; i0 < n; i0 += 32768) {
        const int cnt = n - i0 < 32768 ? n - i0 : 32768;
        resize_fused_kernel<<<dim3(kOutSize / kStripRows, cnt), 256, need, st>>>(
            d_imgs, d_descs + i0, d_tabs, kmax, d_out + (size_t)i0 * kOutSize * kOutSize * 3);
        TISE_LAUNCH_CHECK(ctx);
      }
      return TISE_OK;
    }
  }
  const dim3 blk(256);
  const unsigned gx = (kOutSize * 3 + 255) / 256;
  // gridDim.z is limited to 65535: chunk the batch
  for (int i0 = 0; i0 < n; i0 += 32768) {
    const int cnt = n - i0 < 32768 ? n - i0 : 32768;
    resize_h_kernel<<<dim3(gx, max_h, cnt), blk, 0, st>>>(d_imgs, d_descs + i0, d_tabs, kmax, d_tmp);
    TISE_LAUNCH_CHECK(ctx);
    resize_v_kernel<<<dim3(gx, kOutSize, cnt), blk, 0, st>>>(d_tmp, d_descs + i0, d_tabs, kmax,
                                                             d_out + (size_t)i0 * kOutSize * kOutSize * 3);
    TISE_LAUNCH_CHECK(ctx);
  }
  return TISE_OK;
}

extern "C" int tise_resize_u8(tise_ctx* ctx, const uint8_t* d_imgs, const int64_t* h_offsets, const int32_t* h_hw,
                              int uniform, int n, uint8_t* d_out, void* d_ws, size_t ws_bytes, tise_stream stream) {
  return resize_impl(ctx, "tise_resize_u8", d_imgs, h_offsets, h_hw, nullptr, uniform, n, d_out, d_ws, ws_bytes, stream);
}

extern "C" int tise_crop_resize_u8(tise_ctx* ctx, const uint8_t* d_imgs, const int64_t* h_offsets, const int32_t* h_hw,
                                   const int32_t* h_box, int n, uint8_t* d_out, void* d_ws, size_t ws_bytes,
                                   tise_stream stream) {
  if (ctx && n > 0 && (!h_box || !h_offsets)) return fail(ctx, TISE_ERR_INVALID, "tise_crop_resize_u8: null argument");
  return resize_impl(ctx, "tise_crop_resize_u8", d_imgs, h_offsets, h_hw, h_box, 0, n, d_out, d_ws, ws_bytes, stream);
}
