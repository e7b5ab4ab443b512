// TEST HARNESS ONLY (never linked into libtise_b200.so): compiles tise_toolbox_b200/csrc/jpeg_core.h -- the source of the
// device JPEG decoder -- for the host, so that the CPU test-suite can check it bit for bit against Pillow / libjpeg-turbo.
//   g++ -O2 -shared -fPIC tests/jpeg_host.cpp -o tests/_build/jpeg_host.so
#include <string.h>

#include <vector>

#include "../tise_toolbox_b200/csrc/jpeg_core.h"

using namespace tise::jpg;

extern "C" {

// JPEG file -> RGB8 [h][w][3]; returns 0 / Status (> 0) / -1 not a JPEG / -2 unsupported; hw[0..1] = height, width
int host_jpeg_decode(const uint8_t* file, int64_t n, uint8_t* out, int64_t out_cap, int32_t* hw) {
  Header* h = new Header;
  int rc = parse(file, n, h);
  hw[0] = h->height;
  hw[1] = h->width;
  if (rc) {
    delete h;
    return -1;
  }
  if (!h->supported) {
    delete h;
    return -2;
  }
  if ((int64_t)h->width * h->height * 3 > out_cap) {
    delete h;
    return -3;
  }
  DTable* work = new DTable[6];
  int64_t comp_off[3] = {0, 0, 0}, total = 0;
  for (int c = 0; c < h->ncomp; ++c) {
    comp_off[c] = total;
    total += (int64_t)h->comp[c].blocks_w * h->comp[c].blocks_h * 64;
  }
  std::vector<int16_t> coef(total, 0);
  int st = decode_image(h, h->scans, h->tables, file, work, coef.data(), comp_off);
  if (!st) {
    std::vector<std::vector<uint8_t>> planes(h->ncomp);
    const uint8_t* pl[3] = {nullptr, nullptr, nullptr};
    for (int c = 0; c < h->ncomp; ++c) {
      const Component& k = h->comp[c];
      planes[c].resize((size_t)k.blocks_w * 8 * k.blocks_h * 8);
      for (int by = 0; by < k.blocks_h; ++by)
        for (int bx = 0; bx < k.blocks_w; ++bx)
          idct_islow(coef.data() + comp_off[c] + ((int64_t)by * k.blocks_w + bx) * 64, h->qt[k.tq],
                     planes[c].data() + ((size_t)by * 8 * k.blocks_w * 8 + bx * 8), k.blocks_w * 8);
      pl[c] = planes[c].data();
    }
    for (int y = 0; y < h->height; ++y)
      for (int x = 0; x < h->width; ++x) output_pixel(h, pl, x, y, out + ((int64_t)y * h->width + x) * 3);
  }
  delete[] work;
  delete h;
  return st;
}

}  // extern "C"
