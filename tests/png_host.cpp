// TEST HARNESS ONLY (never linked into libtise_b200.so): compiles tise_toolbox_b200/csrc/png_core.h -- the
// source of the device PNG decoder -- for the host with LANES = 1, so that the CPU test-suite can check the
// inflate / unfilter / parser logic against zlib and Pillow on machines without a GPU.
//   g++ -O2 -shared -fPIC tests/png_host.cpp -o tests/_build/png_host.so
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../tise_toolbox_b200/csrc/png_core.h"

using namespace tise::png;

extern "C" {

// zlib stream -> out; returns Status, *out_len bytes written, *adler = trailer
int host_inflate(const uint8_t* stream, uint32_t n, uint8_t* out, uint32_t cap, uint32_t* out_len, uint32_t* adler) {
  std::vector<uint32_t> words((n + 3) / 4 + 4, 0);
  memcpy(words.data(), stream, n);
  Huff* hl = new Huff;
  Huff* hd = new Huff;
  uint8_t lens[320];
  uint16_t win[32];
  const int st = inflate_zlib<1>(words.data(), n, out, cap, out_len, adler, hl, hd, lens, win, 0);
  delete hl;
  delete hd;
  return st;
}

// PNG file -> RGB8 [h][w][3]; returns 0 / Status (> 0) / -1 not a PNG / -2 unsupported; hw[0..1] = height, width
int host_png_decode(const uint8_t* file, int64_t n, uint8_t* out, int64_t out_cap, int32_t* hw) {
  Header h;
  std::vector<uint8_t> stream;
  if (parse(file, n, &h, [&](int64_t off, int64_t bytes) { stream.insert(stream.end(), file + off, file + off + bytes); }))
    return -1;
  hw[0] = h.height;
  hw[1] = h.width;
  if (!h.supported) return -2;
  if ((int64_t)h.width * h.height * 3 > out_cap) return -3;
  const int stride = stride_of(h.width, h.color_type, h.bit_depth);
  const uint32_t raw_bytes = (uint32_t)raw_bytes_of(h.width, h.height, h.color_type, h.bit_depth, h.interlace);
  std::vector<uint8_t> raw(raw_bytes + 16);
  uint32_t out_len = 0, adler = 0;
  int st = host_inflate(stream.data(), (uint32_t)stream.size(), raw.data(), raw_bytes, &out_len, &adler);
  if (st) return st;
  if (out_len != raw_bytes) return kSizeMismatch;
  std::vector<uint8_t> rows(2 * (size_t)stride + 16);
  return unfilter_to_rgb<1>(raw.data(), h.width, h.height, h.bit_depth, h.color_type, h.interlace,
                            h.plte_off >= 0 ? file + h.plte_off : nullptr, h.plte_entries, rows.data(), out, adler, 0);
}

}  // extern "C"
