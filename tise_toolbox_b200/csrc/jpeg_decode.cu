// A1: baseline JPEG files -> packed RGB8 HWC images in HBM, bit-identical to Pillow's
// Image.open(f).convert("RGB") (libjpeg-turbo defaults: islow IDCT, fancy upsampling) -- image_realism/FID/img_data.py:21.
// The algorithms live in jpeg_core.h (also compiled for the host by the CPU test-suite).  Three kernels per batch:
//   jpeg_entropy_kernel  one warp per image: the lanes zero the coefficient planes, lane 0 walks the entropy-coded
//                        segments of all scans (byte-stuffed, MSB first, restart intervals; sequential or progressive
//                        with spectral selection / successive approximation), rebuilding the Huffman look-up tables
//                        of each scan in shared memory                                     (latency bound, serial stream)
//   jpeg_idct_kernel     one thread per 8 x 8 block: dequantise + jpeg_idct_islow into uint8 component planes
//   jpeg_color_kernel    one thread per output pixel: h2v1 / h2v2 triangle upsampling of Cb / Cr + YCbCr -> RGB
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"
#include "jpeg_core.h"

namespace tise {

struct JpegDesc {
  jpg::Core h;
  long long scan_base;    // first entry of this image in the scan pool / the Huffman-definition pool
  long long table_base;
  long long file_off;     // byte offset of the file in d_files
  long long coef_off;     // byte offset of the int16 coefficient planes in the workspace
  long long plane_off;    // byte offset of the uint8 component planes in the workspace
  long long out_off;      // byte offset of the RGB8 image in d_out
  long long comp_coef[3];  // int16 index of each component's first coefficient (relative to coef_off)
  long long comp_plane[3];  // byte index of each component's plane (relative to plane_off)
  long long blocks;       // blocks of all components
};

constexpr int kJpegWarps = 4;
struct JpegSmem {
  jpg::DTable work[6];  // 3 DC + 3 AC tables of the scan being decoded
};

__global__ void __launch_bounds__(kJpegWarps * 32) jpeg_entropy_kernel(const JpegDesc* __restrict__ descs, int n,
                                                                       const jpg::Scan* __restrict__ scans,
                                                                       const jpg::HuffDef* __restrict__ tables,
                                                                       const uint8_t* __restrict__ files, uint8_t* ws,
                                                                       int* __restrict__ status) {
  __shared__ JpegSmem sm[kJpegWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int img = blockIdx.x * kJpegWarps + warp;
  if (img >= n) return;
  const JpegDesc& d = descs[img];
  uint32_t* z = reinterpret_cast<uint32_t*>(ws + d.coef_off);
  const long long words = d.blocks * 32;  // 64 int16 per block
  for (long long i = lane; i < words; i += 32) z[i] = 0u;
  __syncwarp();
  int st = 0;
  if (lane == 0)
    st = jpg::decode_image(&d.h, scans + d.scan_base, tables + d.table_base, files + d.file_off, sm[warp].work,
                           reinterpret_cast<int16_t*>(ws + d.coef_off),
                           reinterpret_cast<const int64_t*>(d.comp_coef));
  if (lane == 0) status[img] = st;
}

// grid = (ceil(max blocks / 256), n)
__global__ void __launch_bounds__(256) jpeg_idct_kernel(const JpegDesc* __restrict__ descs, uint8_t* ws, const int* __restrict__ status) {
  const JpegDesc& d = descs[blockIdx.y];
  if (status[blockIdx.y] != 0) return;
  long long b = (long long)blockIdx.x * 256 + threadIdx.x;
  if (b >= d.blocks) return;
  int c = 0;
  for (; c < d.h.ncomp - 1; ++c) {
    const long long nb = (long long)d.h.comp[c].blocks_w * d.h.comp[c].blocks_h;
    if (b < nb) break;
    b -= nb;
  }
  const jpg::Component& k = d.h.comp[c];
  const int by = (int)(b / k.blocks_w), bx = (int)(b - (long long)by * k.blocks_w);
  const int16_t* coef = reinterpret_cast<const int16_t*>(ws + d.coef_off) + d.comp_coef[c] + b * 64;
  uint8_t* plane = ws + d.plane_off + d.comp_plane[c];
  jpg::idct_islow(coef, d.h.qt[k.tq], plane + ((size_t)by * 8 * k.blocks_w * 8 + (size_t)bx * 8), k.blocks_w * 8);
}

// grid = (ceil(max pixels / 256), n)
__global__ void __launch_bounds__(256) jpeg_color_kernel(const JpegDesc* __restrict__ descs, const uint8_t* ws,
                                                         uint8_t* __restrict__ out, const int* __restrict__ status) {
  const JpegDesc& d = descs[blockIdx.y];
  if (status[blockIdx.y] != 0) return;
  const long long p = (long long)blockIdx.x * 256 + threadIdx.x;
  if (p >= (long long)d.h.width * d.h.height) return;
  const int y = (int)(p / d.h.width), x = (int)(p - (long long)y * d.h.width);
  const uint8_t* planes[3];
  for (int c = 0; c < 3; ++c) planes[c] = ws + d.plane_off + d.comp_plane[c < d.h.ncomp ? c : 0];
  uint8_t rgb[3];
  jpg::output_pixel(&d.h, planes, x, y, rgb);
  uint8_t* o = out + d.out_off + p * 3;
  o[0] = rgb[0];
  o[1] = rgb[1];
  o[2] = rgb[2];
}

static long long jpeg_blocks(const jpg::Core& h) {
  long long b = 0;
  for (int c = 0; c < h.ncomp; ++c) b += (long long)h.comp[c].blocks_w * h.comp[c].blocks_h;
  return b;
}

static void fill_jpeg_info(const jpg::Header& h, tise_jpeg_info* info) {
  info->width = h.width;
  info->height = h.height;
  info->components = h.ncomp;
  info->chroma_mode = h.chroma_mode;
  info->restart_interval = h.nscans ? h.scans[0].restart_interval : 0;
  info->progressive = h.progressive;
  info->scans = h.nscans;
  info->huffman_tables = h.ntables;
  info->supported = h.supported;
  info->scan_bytes = h.scan_bytes;
  info->blocks = h.supported ? jpeg_blocks(h) : 0;
}

}  // namespace tise

using namespace tise;

extern "C" int tise_jpeg_probe(const uint8_t* h_file, int64_t file_bytes, tise_jpeg_info* info) {
  if (!h_file || !info || file_bytes < 0) return TISE_ERR_INVALID;
  jpg::Header* h = new jpg::Header;
  const int rc = jpg::parse(h_file, file_bytes, h);
  fill_jpeg_info(*h, info);
  delete h;
  if (rc) {
    info->supported = 0;
    return TISE_ERR_INVALID;
  }
  return TISE_OK;
}

extern "C" size_t tise_jpeg_decode_workspace_bytes(const tise_jpeg_info* infos, int n) {
  if (!infos || n <= 0) return 0;
  size_t total = align_up(sizeof(JpegDesc) * (size_t)n, 256);
  size_t n_scans = 0, n_tables = 0;
  for (int i = 0; i < n; ++i) {
    total += align_up((size_t)infos[i].blocks * 128, 256) + align_up((size_t)infos[i].blocks * 64, 256);
    n_scans += (size_t)infos[i].scans;
    n_tables += (size_t)infos[i].huffman_tables;
  }
  return total + align_up(sizeof(jpg::Scan) * n_scans, 256) + align_up(sizeof(jpg::HuffDef) * n_tables, 256);
}

extern "C" int tise_jpeg_decode(tise_ctx* ctx, const uint8_t* h_files, const uint8_t* d_files, const int64_t* h_file_off,
                                const int64_t* h_file_bytes, int n, uint8_t* d_out, const int64_t* h_out_off,
                                int32_t* d_status, void* d_ws, size_t ws_bytes, tise_stream stream) {
  if (!ctx) return TISE_ERR_INVALID;
  if (n == 0) return TISE_OK;
  if (!h_files || !d_files || !h_file_off || !h_file_bytes || !d_out || !h_out_off || !d_status || !d_ws || n < 0)
    return fail(ctx, TISE_ERR_INVALID, "tise_jpeg_decode: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<JpegDesc> descs(n);
  std::vector<jpg::Scan> scans;
  std::vector<jpg::HuffDef> tables;
  jpg::Header* hdr = new jpg::Header;
  long long max_blocks = 0, max_pixels = 0;
  for (int i = 0; i < n; ++i) {
    JpegDesc& d = descs[i];
    const int rc = jpg::parse(h_files + h_file_off[i], h_file_bytes[i], hdr);
    if (rc || !hdr->supported) {
      delete hdr;
      return fail(ctx, TISE_ERR_INVALID, "tise_jpeg_decode: file %d is %s (tise_jpeg_probe reports what the device decoder handles)",
                  i, rc ? "not a JPEG" : "a JPEG variant the device decoder does not handle");
    }
    d.h = *static_cast<const jpg::Core*>(hdr);
    d.scan_base = (long long)scans.size();
    d.table_base = (long long)tables.size();
    scans.insert(scans.end(), hdr->scans, hdr->scans + hdr->nscans);
    tables.insert(tables.end(), hdr->tables, hdr->tables + hdr->ntables);
    d.file_off = h_file_off[i];
    d.out_off = h_out_off[i];
    d.blocks = jpeg_blocks(d.h);
    long long ci = 0, pi = 0;
    for (int c = 0; c < 3; ++c) {
      d.comp_coef[c] = ci;
      d.comp_plane[c] = pi;
      if (c < d.h.ncomp) {
        const long long nb = (long long)d.h.comp[c].blocks_w * d.h.comp[c].blocks_h;
        ci += nb * 64;
        pi += nb * 64;
      }
    }
    max_blocks = d.blocks > max_blocks ? d.blocks : max_blocks;
    const long long px = (long long)d.h.width * d.h.height;
    max_pixels = px > max_pixels ? px : max_pixels;
  }
  delete hdr;
  // workspace: [descs][scan pool][Huffman-definition pool][per image: coefficients, planes]
  size_t off = align_up(sizeof(JpegDesc) * (size_t)n, 256);
  const size_t off_scans = off;
  off += align_up(sizeof(jpg::Scan) * scans.size(), 256);
  const size_t off_tables = off;
  off += align_up(sizeof(jpg::HuffDef) * tables.size(), 256);
  for (int i = 0; i < n; ++i) {
    JpegDesc& d = descs[i];
    d.coef_off = (long long)off;
    off += align_up((size_t)d.blocks * 128, 256);
    d.plane_off = (long long)off;
    off += align_up((size_t)d.blocks * 64, 256);
  }
  if (ws_bytes < off) return fail(ctx, TISE_ERR_WORKSPACE, "tise_jpeg_decode: workspace %zu < %zu bytes", ws_bytes, off);
  uint8_t* ws = (uint8_t*)d_ws;
  JpegDesc* d_descs = (JpegDesc*)ws;
  jpg::Scan* d_scans = (jpg::Scan*)(ws + off_scans);
  jpg::HuffDef* d_tables = (jpg::HuffDef*)(ws + off_tables);
  // (through the context's pinned staging ring: no stream synchronisation, common.cuh)
  int urc = stage_upload(ctx, d_descs, descs.data(), sizeof(JpegDesc) * n, st);
  if (!urc) urc = stage_upload(ctx, d_scans, scans.data(), sizeof(jpg::Scan) * scans.size(), st);
  if (!urc) urc = stage_upload(ctx, d_tables, tables.data(), sizeof(jpg::HuffDef) * tables.size(), st);
  if (urc) return urc;
  TISE_CUDA(ctx, cudaMemsetAsync(d_status, 0xff, sizeof(int32_t) * n, st));
  for (int i0 = 0; i0 < n; i0 += 32768) {  // gridDim.y <= 65535
    const int cnt = n - i0 < 32768 ? n - i0 : 32768;
    jpeg_entropy_kernel<<<cdiv(cnt, kJpegWarps), kJpegWarps * 32, 0, st>>>(d_descs + i0, cnt, d_scans, d_tables, d_files, ws,
                                                                           d_status + i0);
    TISE_LAUNCH_CHECK(ctx);
    jpeg_idct_kernel<<<dim3((unsigned)cdiv(max_blocks, 256), cnt), 256, 0, st>>>(d_descs + i0, ws, d_status + i0);
    TISE_LAUNCH_CHECK(ctx);
    jpeg_color_kernel<<<dim3((unsigned)cdiv(max_pixels, 256), cnt), 256, 0, st>>>(d_descs + i0, ws, d_out, d_status + i0);
    TISE_LAUNCH_CHECK(ctx);
  }
  return TISE_OK;
}
