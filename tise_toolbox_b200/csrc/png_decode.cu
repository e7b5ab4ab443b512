// A1: PNG files -> packed RGB8 HWC images in HBM, the input format of tise_resize_u8.
// Replaces PIL.Image.open(f).convert("RGB") on DataLoader workers (image_realism/FID/img_data.py:19-25,
// object_fidelity/O-IS/object_centric_inception_score.py:100-102).
//
// Host: walk the chunk list of every file (IHDR, PLTE, IDAT segments) -- a few dozen bytes touched per file.
// Device, three kernels per batch:
//   png_gather_kernel    IDAT payloads -> one contiguous, word-aligned zlib stream per image        (HBM copy)
//   png_inflate_kernel   one warp per image: warp-uniform bit reader + Huffman decoder in registers; literal runs
//                        are decoded 32 candidate bit positions at a time (one table look-up per lane, the chain of
//                        code boundaries followed through a 64-byte shared-memory window, literals stored side by
//                        side); 32 lanes copy LZ77 matches and build the lookup tables  (latency / issue bound)
//   png_unfilter_kernel  one warp per image: scanline rows staged in shared memory, Sub/Avg/Paeth run as one
//                        sequential chain per byte lane, Up/None and the RGB conversion across the warp;
//                        the Adler-32 of the inflated bytes is checked on the way                  (shared memory)
// The algorithms live in png_core.h, which also compiles for the host so that the CPU test-suite can check them.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"
#include "png_core.h"

namespace tise {

struct PngDesc {
  long long stream_off;  // byte offset of the compacted zlib stream in the workspace (16-byte aligned)
  long long raw_off;     // byte offset of the inflated scanlines in the workspace
  long long out_off;     // byte offset of the RGB8 image in d_out
  long long plte_off;    // absolute byte offset of the PLTE data in d_files, -1: none
  int stream_bytes, raw_bytes;
  int width, height, bit_depth, color_type, plte_entries, interlace, pad;
  unsigned adler;        // written by the inflate kernel
};
struct PngSeg {
  long long src;  // absolute byte offset in d_files
  long long dst;  // absolute byte offset in the workspace
  int bytes, pad;
};

__global__ void png_gather_kernel(const uint8_t* __restrict__ files, const PngSeg* __restrict__ segs, uint8_t* __restrict__ ws) {
  const PngSeg s = segs[blockIdx.x];
  const uint8_t* src = files + s.src;
  uint8_t* dst = ws + s.dst;
  // head bytes up to a 4-byte aligned destination, then 32-bit stores assembled from two aligned 32-bit loads
  int head = (int)((4 - (s.dst & 3)) & 3);
  if (head > s.bytes) head = s.bytes;
  if ((int)threadIdx.x < head) dst[threadIdx.x] = src[threadIdx.x];
  const int body = (s.bytes - head) >> 2;
  const uint8_t* bsrc = src + head;
  const int mis = (int)((uintptr_t)bsrc & 3);
  const uint32_t* asrc = reinterpret_cast<const uint32_t*>(bsrc - mis);
  uint32_t* adst = reinterpret_cast<uint32_t*>(dst + head);
  for (int i = threadIdx.x; i < body; i += blockDim.x) {
    uint32_t lo = __ldg(asrc + i);
    if (mis) lo = __funnelshift_r(lo, __ldg(asrc + i + 1), mis * 8);
    adst[i] = lo;
  }
  const int done = head + body * 4;
  if ((int)threadIdx.x < s.bytes - done) dst[done + threadIdx.x] = src[done + threadIdx.x];
}

constexpr int kInflateWarps = 4;
struct InflateSmem {
  png::Huff hl, hd;
  uint8_t lens[320];
  uint16_t win[32];
};

__global__ void __launch_bounds__(kInflateWarps * 32) png_inflate_kernel(PngDesc* __restrict__ descs, int n, uint8_t* ws,
                                                                         int* __restrict__ status) {
  __shared__ InflateSmem sm[kInflateWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int img = blockIdx.x * kInflateWarps + warp;
  if (img >= n) return;
  PngDesc& d = descs[img];
  uint32_t out_len = 0, adler = 0;
  int st = png::inflate_zlib<32>(reinterpret_cast<const uint32_t*>(ws + d.stream_off), (uint32_t)d.stream_bytes,
                                 ws + d.raw_off, (uint32_t)d.raw_bytes, &out_len, &adler, &sm[warp].hl, &sm[warp].hd,
                                 sm[warp].lens, sm[warp].win, lane);
  if (st == png::kOk && out_len != (uint32_t)d.raw_bytes) st = png::kSizeMismatch;
  if (lane == 0) {
    d.adler = adler;
    status[img] = st;
  }
}

__global__ void __launch_bounds__(32) png_unfilter_kernel(const PngDesc* __restrict__ descs, const uint8_t* __restrict__ files,
                                                          const uint8_t* ws, uint8_t* __restrict__ out, int* status) {
  extern __shared__ uint8_t rows[];
  const int img = blockIdx.x, lane = threadIdx.x;
  if (status[img] != png::kOk) return;  // inflate failed: leave its status
  const PngDesc d = descs[img];
  const int st = png::unfilter_to_rgb<32>(ws + d.raw_off, d.width, d.height, d.bit_depth, d.color_type, d.interlace,
                                          d.plte_off >= 0 ? files + d.plte_off : nullptr, d.plte_entries, rows,
                                          out + d.out_off, d.adler, lane);
  if (lane == 0 && st != png::kOk) status[img] = st;
}

struct PngWs {
  size_t off_desc, off_seg, off_stream, off_raw, total;
};
// n images, n_seg IDAT chunks, stream bytes already padded per image, raw bytes already padded per image
static PngWs png_ws_layout(long long n, long long n_seg, long long stream_bytes, long long raw_bytes) {
  PngWs L;
  size_t o = 0;
  L.off_desc = o;
  o = align_up(o + sizeof(PngDesc) * (size_t)n, 256);
  L.off_seg = o;
  o = align_up(o + sizeof(PngSeg) * (size_t)n_seg, 256);
  L.off_stream = o;
  o = align_up(o + (size_t)stream_bytes, 256);
  L.off_raw = o;
  o = align_up(o + (size_t)raw_bytes, 256);
  L.total = o;
  return L;
}
static inline long long pad_stream(long long b) { return (long long)align_up((size_t)b + 16, 16); }
static inline long long pad_raw(long long b) { return (long long)align_up((size_t)b, 16); }

constexpr long long kPngRowSmemBudget = 232448 - 1024;  // two scanlines next to the kernel's static shared memory

static void fill_info(const png::Header& h, tise_png_info* info) {
  info->width = h.width;
  info->height = h.height;
  info->bit_depth = h.bit_depth;
  info->color_type = h.color_type;
  info->interlace = h.interlace;
  info->channels = h.channels;
  info->n_idat = h.n_idat;
  info->supported = h.supported;
  info->idat_bytes = h.idat_bytes;
}

}  // namespace tise

using namespace tise;

extern "C" int tise_png_probe(const uint8_t* h_file, int64_t file_bytes, tise_png_info* info) {
  if (!h_file || !info || file_bytes < 0) return TISE_ERR_INVALID;
  png::Header h;
  const int rc = png::parse(h_file, file_bytes, &h, [](int64_t, int64_t) {});
  fill_info(h, info);
  if (rc) {
    info->supported = 0;
    return TISE_ERR_INVALID;
  }
  // the unfilter kernel keeps two scanlines in shared memory: a file whose scanline does not fit in the smallest opt-in
  // limit this library runs with (227 KB) is reported as unsupported HERE, so callers route it to their host decoder
  // instead of having a whole tise_png_decode call rejected later
  if (2 * (long long)png::stride_of(h.width, h.color_type, h.bit_depth) > kPngRowSmemBudget) info->supported = 0;
  return TISE_OK;
}

extern "C" size_t tise_png_decode_workspace_bytes(const tise_png_info* infos, int n) {
  if (!infos || n <= 0) return 0;
  long long n_seg = 0, sb = 0, rb = 0;
  for (int i = 0; i < n; ++i) {
    n_seg += infos[i].n_idat;
    sb += pad_stream(infos[i].idat_bytes);
    rb += pad_raw(png::raw_bytes_of(infos[i].width, infos[i].height, infos[i].color_type, infos[i].bit_depth, infos[i].interlace));
  }
  return png_ws_layout(n, n_seg, sb, rb).total;
}

extern "C" int tise_png_decode(tise_ctx* ctx, const uint8_t* h_files, const uint8_t* d_files, const int64_t* h_file_off,
                               const int64_t* h_file_bytes, int n, uint8_t* d_out, const int64_t* h_out_off,
                               int32_t* d_status, void* d_ws, size_t ws_bytes, tise_stream stream) {
  if (!ctx) return TISE_ERR_INVALID;
  if (n == 0) return TISE_OK;
  if (!h_files || !d_files || !h_file_off || !h_file_bytes || !d_out || !h_out_off || !d_status || !d_ws || n < 0)
    return fail(ctx, TISE_ERR_INVALID, "tise_png_decode: null argument");
  cudaStream_t st = (cudaStream_t)stream;

  std::vector<PngDesc> descs(n);
  std::vector<PngSeg> segs;
  long long stream_total = 0, raw_total = 0;
  int max_stride = 0;
  for (int i = 0; i < n; ++i) {
    png::Header h;
    const long long seg0 = (long long)segs.size();
    const int64_t base = h_file_off[i];
    long long in_stream = 0;
    const int rc = png::parse(h_files + base, h_file_bytes[i], &h, [&](int64_t off, int64_t bytes) {
      PngSeg s;
      s.src = base + off;
      s.dst = stream_total + in_stream;  // + off_stream below
      s.bytes = (int)bytes;
      s.pad = 0;
      segs.push_back(s);
      in_stream += bytes;
    });
    if (rc || !h.supported)
      return fail(ctx, TISE_ERR_INVALID,
                  "tise_png_decode: file %d is %s (tise_png_probe reports what the device decoder handles: every "
                  "colour type at 8 bits, 16-bit RGB / RGBA / grey+alpha, 1/2/4-bit grey / palette, plain or Adam7)",
                  i, rc ? "not a well-formed PNG" : "a PNG variant the device decoder does not handle");
    (void)seg0;
    PngDesc& d = descs[i];
    d.stream_off = stream_total;
    d.stream_bytes = (int)h.idat_bytes;
    d.raw_off = raw_total;
    const int stride = png::stride_of(h.width, h.color_type, h.bit_depth);
    const long long raw = png::raw_bytes_of(h.width, h.height, h.color_type, h.bit_depth, h.interlace);
    if (raw > 0x7fffffffLL) return fail(ctx, TISE_ERR_INVALID, "tise_png_decode: file %d is too large", i);
    d.raw_bytes = (int)raw;
    d.out_off = h_out_off[i];
    d.plte_off = h.plte_off >= 0 ? base + h.plte_off : -1;
    d.plte_entries = h.plte_entries;
    d.width = h.width;
    d.height = h.height;
    d.bit_depth = h.bit_depth;
    d.interlace = h.interlace;
    d.pad = 0;
    d.color_type = h.color_type;
    d.adler = 0;
    stream_total += pad_stream(h.idat_bytes);
    raw_total += pad_raw(raw);
    max_stride = stride > max_stride ? stride : max_stride;
  }
  const PngWs L = png_ws_layout(n, (long long)segs.size(), stream_total, raw_total);
  if (ws_bytes < L.total)
    return fail(ctx, TISE_ERR_WORKSPACE, "tise_png_decode: workspace %zu < %zu bytes", ws_bytes, L.total);
  const size_t row_smem = 2 * (size_t)max_stride;
  if (row_smem > (size_t)ctx->smem_optin - 1024)
    return fail(ctx, TISE_ERR_INVALID, "tise_png_decode: scanlines of %d bytes do not fit in shared memory", max_stride);
  for (auto& d : descs) {
    d.stream_off += (long long)L.off_stream;
    d.raw_off += (long long)L.off_raw;
  }
  for (auto& s : segs) s.dst += (long long)L.off_stream;

  uint8_t* ws = (uint8_t*)d_ws;
  PngDesc* d_descs = (PngDesc*)(ws + L.off_desc);
  PngSeg* d_segs = (PngSeg*)(ws + L.off_seg);
  int urc = stage_upload(ctx, d_descs, descs.data(), sizeof(PngDesc) * n, st);
  if (!urc) urc = stage_upload(ctx, d_segs, segs.data(), sizeof(PngSeg) * segs.size(), st);
  if (urc) return urc;
  TISE_CUDA(ctx, cudaMemsetAsync(d_status, 0xff, sizeof(int32_t) * n, st));  // -1 until the inflate kernel writes it

  png_gather_kernel<<<(unsigned)segs.size(), 256, 0, st>>>(d_files, d_segs, ws);
  TISE_LAUNCH_CHECK(ctx);
  png_inflate_kernel<<<cdiv(n, kInflateWarps), kInflateWarps * 32, 0, st>>>(d_descs, n, ws, d_status);
  TISE_LAUNCH_CHECK(ctx);
  if (row_smem > 48 * 1024)
    TISE_CUDA(ctx, cudaFuncSetAttribute(png_unfilter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_smem));
  png_unfilter_kernel<<<n, 32, row_smem, st>>>(d_descs, d_files, ws, d_out, d_status);
  TISE_LAUNCH_CHECK(ctx);
  return TISE_OK;
}
