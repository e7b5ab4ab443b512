// A1: baseline JPEG decode building blocks, written once for the device and the host (same scheme as png_core.h: the
// host build exists only so that the CPU test-suite can check the decoder against Pillow / libjpeg-turbo bit for bit).
//
// Replaces PIL.Image.open(f).convert("RGB") (image_realism/FID/img_data.py:21) for JPEG files -- the reference's file
// filter takes "jpg" as well as "png" (img_data.py:31), and real-image sets such as MS-COCO are JPEG.  Pillow decodes
// with libjpeg-turbo's defaults, which are reproduced operation by operation:
//   entropy decoding   jdhuff.c      sequential Huffman, restart intervals
//   inverse DCT        jidctint.c    jpeg_idct_islow: 13-bit fixed point, two passes, range-limit table
//   chroma upsampling  jdsample.c    "fancy" triangle filters h2v1 / h2v2 (edge rows / columns replicated as jdmainct.c
//                                    does), or none for 4:4:4
//   colour conversion  jdcolor.c     YCbCr -> RGB with 16-bit fixed-point tables
// Handled: 8-bit Huffman JPEGs -- baseline / extended sequential (SOF0 / SOF1, interleaved or one scan per component) and
// progressive (SOF2: spectral selection + successive approximation, jdphuff.c) -- grey or YCbCr with chroma sampling
// 4:4:4, 4:2:2 (h2v1) or 4:2:0 (h2v2), restart intervals.  Arithmetic coding, CMYK / RGB-tagged files and other
// sampling layouts are reported as unsupported by parse() and stay on the host.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define TISE_JHD __host__ __device__ __forceinline__
#else
#define TISE_JHD inline
#endif

namespace tise {
namespace jpg {

enum Status : int32_t {
  kOk = 0,
  kBadHuffmanCode = 21,  // bit pattern that is no code
  kBadCoefficient = 22,  // run past the 64th coefficient, or a DC size category > 15
  kBadRestart = 23,      // expected RSTn marker missing
  kTruncated = 24,       // entropy-coded data ended before the last MCU
  kNotDecoded = 25
};

struct Component {
  int32_t id, h, v, tq;
  int32_t blocks_w, blocks_h;  // blocks per row / column of this component's plane (whole MCUs)
  int32_t ds_w, ds_h;          // ceil(width * h / hmax), ceil(height * v / vmax): the real samples
};

constexpr int kMaxScans = 24;   // progressive files written by libjpeg have ~10 scans
constexpr int kMaxTables = 40;  // every DHT definition met in the file, in order (tables may be redefined between scans)

struct Scan {
  int32_t ncomp, comp[3];        // indices into Header::comp
  int32_t dc_tab[3], ac_tab[3];  // indices into the table pool, -1 = none defined
  int32_t ss, se, ah, al;        // spectral selection / successive approximation (0, 63, 0, 0 for sequential scans)
  int32_t restart_interval;
  int32_t pad;
  int64_t off, bytes;            // entropy-coded segment of this scan, relative to the first byte of the file
};

struct HuffDef {
  uint8_t bits[17];  // code length -> number of codes
  uint8_t vals[256];
  uint8_t pad[15];
};

// What the kernels need of one image (the scans and Huffman definitions travel in pools next to it).
struct Core {
  int32_t width, height, ncomp, hmax, vmax;
  int32_t mcus_x, mcus_y;
  int32_t supported;
  int32_t chroma_mode;  // 0 grey, 1 4:4:4, 2 h2v1, 3 h2v2
  int32_t progressive, nscans, ntables;
  int64_t scan_bytes;   // entropy-coded bytes of all scans
  Component comp[3];
  uint16_t qt[4][64];   // quantisation tables in NATURAL order
  uint8_t qt_present[4];
  uint8_t pad[4];
};

struct Header : Core {
  Scan scans[kMaxScans];
  HuffDef tables[kMaxTables];
};

TISE_JHD int natural_order(int k) {
  // zigzag index -> natural (row-major) index, jutils.c jpeg_natural_order
  const uint8_t t[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                         41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                         30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
  return t[k];
}

// ---------------------------------------------------------------------------------------------------------
// Host parser.  Returns 0, or -1 if the bytes are not a JPEG container at all.
inline int parse(const uint8_t* f, int64_t n, Header* h) {
  *h = Header();
  if (n < 4 || f[0] != 0xFF || f[1] != 0xD8) return -1;
  int64_t p = 2;
  bool have_sof = false, ok = true, sequential_done = false;
  int adobe_transform = -1, restart_interval = 0;
  int cur[2][4] = {{-1, -1, -1, -1}, {-1, -1, -1, -1}};  // current table of each (class, id) -> pool index
  bool covered[3] = {false, false, false};
  while (p + 4 <= n && !sequential_done) {
    if (f[p] != 0xFF) return -1;
    while (p < n && f[p] == 0xFF) ++p;  // fill bytes
    if (p >= n) return -1;
    const int m = f[p++];
    if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
    if (m == 0xD9) break;
    if (p + 2 > n) return -1;
    const int len = (f[p] << 8) | f[p + 1];
    if (len < 2 || p + len > n) return -1;
    const uint8_t* d = f + p + 2;
    const int dl = len - 2;
    if (m == 0xDB) {  // DQT
      if (h->nscans) ok = false;  // (libjpeg latches a component's table at its first scan; redefinitions: host)
      int q = 0;
      while (q < dl) {
        const int pq = d[q] >> 4, tq = d[q] & 15;
        ++q;
        if (tq > 3 || q + (pq ? 128 : 64) > dl) return -1;
        for (int k = 0; k < 64; ++k) {
          const int v = pq ? ((d[q] << 8) | d[q + 1]) : d[q];
          q += pq ? 2 : 1;
          h->qt[tq][natural_order(k)] = (uint16_t)v;
        }
        h->qt_present[tq] = 1;
      }
    } else if (m == 0xC4) {  // DHT
      int q = 0;
      while (q < dl) {
        const int tc = d[q] >> 4, th = d[q] & 15;
        ++q;
        if (tc > 1 || th > 3 || q + 16 > dl) return -1;
        int total = 0;
        for (int i = 1; i <= 16; ++i) total += d[q + i - 1];
        if (total > 256 || q + 16 + total > dl) return -1;
        if (h->ntables >= kMaxTables) {
          ok = false;
        } else {
          const int t = h->ntables++;
          h->tables[t].bits[0] = 0;
          for (int i = 1; i <= 16; ++i) h->tables[t].bits[i] = d[q + i - 1];
          for (int i = 0; i < total; ++i) h->tables[t].vals[i] = d[q + 16 + i];
          cur[tc][th] = t;
        }
        q += 16 + total;
      }
    } else if (m == 0xC0 || m == 0xC1 || m == 0xC2) {  // SOF0 / SOF1 (sequential), SOF2 (progressive), Huffman
      if (dl < 6 || have_sof) return -1;
      const int prec = d[0];
      h->height = (d[1] << 8) | d[2];
      h->width = (d[3] << 8) | d[4];
      h->ncomp = d[5];
      h->progressive = m == 0xC2;
      if (prec != 8 || (h->ncomp != 1 && h->ncomp != 3) || dl < 6 + 3 * h->ncomp || h->width <= 0 || h->height <= 0) ok = false;
      for (int c = 0; c < h->ncomp && c < 3 && dl >= 6 + 3 * h->ncomp; ++c) {
        h->comp[c].id = d[6 + 3 * c];
        h->comp[c].h = d[7 + 3 * c] >> 4;
        h->comp[c].v = d[7 + 3 * c] & 15;
        h->comp[c].tq = d[8 + 3 * c];
      }
      have_sof = true;
    } else if (m >= 0xC3 && m <= 0xCF && m != 0xC8) {  // lossless, differential, arithmetic ...
      have_sof = true;
      ok = false;
      if (dl >= 6) {
        h->height = (d[1] << 8) | d[2];
        h->width = (d[3] << 8) | d[4];
      }
    } else if (m == 0xDD) {  // DRI
      if (dl < 2) return -1;
      restart_interval = (d[0] << 8) | d[1];
    } else if (m == 0xEE) {  // APP14 "Adobe"
      if (dl >= 12 && d[0] == 'A' && d[1] == 'd' && d[2] == 'o' && d[3] == 'b' && d[4] == 'e') adobe_transform = d[11];
    } else if (m == 0xDA) {  // SOS
      if (!have_sof || dl < 1) return -1;
      const int ns = d[0];
      if (ns < 1 || ns > 4 || dl < 1 + 2 * ns + 3) return -1;
      if (ns > 3 || ns > h->ncomp) {  // CMYK / YCCK: a JPEG, but not one the device decoder takes
        ok = false;
        break;
      }
      const int64_t data = p + len;
      int64_t end = n;
      if (h->nscans >= kMaxScans || !ok) {
        ok = false;
      } else {
        Scan& sc = h->scans[h->nscans];
        sc.ncomp = ns;
        for (int c = 0; c < ns; ++c) {
          int ci = -1;
          for (int k = 0; k < h->ncomp && k < 3; ++k)
            if (h->comp[k].id == d[1 + 2 * c]) ci = k;
          if (ci < 0 || (c > 0 && ci <= sc.comp[c - 1])) ok = false;  // unknown id, or not in frame order
          sc.comp[c] = ci < 0 ? 0 : ci;
          const int td = d[2 + 2 * c] >> 4, ta = d[2 + 2 * c] & 15;
          sc.dc_tab[c] = td < 4 ? cur[0][td] : -1;
          sc.ac_tab[c] = ta < 4 ? cur[1][ta] : -1;
          if (ci >= 0) covered[ci] = true;
        }
        sc.ss = d[1 + 2 * ns];
        sc.se = d[2 + 2 * ns];
        sc.ah = d[3 + 2 * ns] >> 4;
        sc.al = d[3 + 2 * ns] & 15;
        sc.restart_interval = restart_interval;
        sc.off = data;
        if (!h->progressive) {
          sc.ss = 0;
          sc.se = 63;
          sc.ah = sc.al = 0;
        } else if (sc.ss > sc.se || sc.se > 63 || sc.al > 13 || sc.ah > 13 || (sc.ss == 0 && sc.se != 0) || (sc.ss > 0 && ns != 1)) {
          ok = false;
        }
        // tables this kind of scan reads
        for (int c = 0; c < ns; ++c) {
          const bool need_dc = !h->progressive || (sc.ss == 0 && sc.ah == 0);
          const bool need_ac = !h->progressive || sc.ss > 0;
          if ((need_dc && sc.dc_tab[c] < 0) || (need_ac && sc.ac_tab[c] < 0)) ok = false;
        }
        bool all = true;
        for (int c = 0; c < h->ncomp && c < 3; ++c) all = all && covered[c];
        if (!h->progressive && all) {
          sequential_done = true;  // every component decoded: nothing else can follow but EOI; skip the byte walk
        } else {
          int64_t q = data;  // find the marker that ends this scan's entropy-coded segment
          end = n;
          while (q + 1 < n) {
            const uint8_t* ff = (const uint8_t*)memchr(f + q, 0xFF, (size_t)(n - q - 1));
            if (!ff) break;
            q = ff - f;
            const int nx = f[q + 1];
            if (nx == 0x00 || (nx >= 0xD0 && nx <= 0xD7)) {
              q += 2;
            } else if (nx == 0xFF) {
              q += 1;
            } else {
              end = q;
              break;
            }
          }
        }
        sc.bytes = end - data;
        h->scan_bytes += sc.bytes;
        h->nscans++;
      }
      p = end;
      continue;
    }
    p += len;
  }
  if (!have_sof) return -1;
  if (h->nscans == 0) ok = false;
  // what the device decoder handles
  if (ok) {
    for (int c = 0; c < h->ncomp; ++c) {
      const Component& k = h->comp[c];
      if (k.tq > 3 || !h->qt_present[k.tq] || !covered[c]) ok = false;
    }
  }
  if (ok && h->ncomp == 1) {
    h->comp[0].h = h->comp[0].v = 1;  // a single-component scan is never interleaved: sampling factors are moot
    h->chroma_mode = 0;
  } else if (ok) {
    const Component *y = &h->comp[0], *cb = &h->comp[1], *cr = &h->comp[2];
    if (cb->h != 1 || cb->v != 1 || cr->h != 1 || cr->v != 1) ok = false;
    if (y->h == 1 && y->v == 1) h->chroma_mode = 1;
    else if (y->h == 2 && y->v == 1) h->chroma_mode = 2;
    else if (y->h == 2 && y->v == 2) h->chroma_mode = 3;
    else ok = false;
    if (adobe_transform == 0) ok = false;                                    // RGB stored without a colour transform
    if (y->id == 'R' && cb->id == 'G' && cr->id == 'B') ok = false;          // libjpeg's component-id heuristic
    if (adobe_transform < 0 && !(y->id == 1 && cb->id == 2 && cr->id == 3) && !(y->id == 0 && cb->id == 1 && cr->id == 2)) {
      // no JFIF-style ids and no Adobe marker: libjpeg guesses the colour space from the ids; only the plain YCbCr
      // layouts are taken here
      ok = false;
    }
  }
  if (ok) {
    h->hmax = h->comp[0].h;
    h->vmax = h->comp[0].v;
    h->mcus_x = (h->width + 8 * h->hmax - 1) / (8 * h->hmax);
    h->mcus_y = (h->height + 8 * h->vmax - 1) / (8 * h->vmax);
    for (int c = 0; c < h->ncomp; ++c) {
      Component& k = h->comp[c];
      k.blocks_w = h->mcus_x * k.h;
      k.blocks_h = h->mcus_y * k.v;
      k.ds_w = (h->width * k.h + h->hmax - 1) / h->hmax;
      k.ds_h = (h->height * k.v + h->vmax - 1) / h->vmax;
    }
    if ((int64_t)h->width * h->height > ((int64_t)1 << 28)) ok = false;
  }
  h->supported = ok ? 1 : 0;
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Huffman decoding table (jdhuff.c jpeg_make_d_derived_tbl): 8-bit look-ahead + canonical slow path.
struct DTable {
  int32_t maxcode[18];    // largest code of length l, -1 if none; [17] = sentinel
  int32_t valoffset[17];  // vals index of the first code of length l minus that code
  uint16_t look[256];     // (length << 8) | symbol for codes of <= 8 bits, 0 otherwise
  uint8_t vals[256];
};

// serial; call from one lane
TISE_JHD void build_dtable(const uint8_t* bits, const uint8_t* vals, DTable* t) {
  int code = 0, p = 0;
  for (int i = 0; i < 256; ++i) t->look[i] = 0;
  for (int l = 1; l <= 16; ++l) {
    const int nb = bits[l];
    if (nb) {
      t->valoffset[l] = p - code;
      for (int i = 0; i < nb; ++i, ++p, ++code) {
        t->vals[p] = vals[p];
        if (l <= 8) {
          const int first = code << (8 - l);
          for (int j = 0; j < (1 << (8 - l)); ++j) t->look[(first + j) & 255] = (uint16_t)((l << 8) | vals[p]);
        }
      }
      t->maxcode[l] = code - 1;
    } else {
      t->maxcode[l] = -1;
      t->valoffset[l] = 0;
    }
    code <<= 1;
  }
  t->maxcode[17] = 0xFFFFF;
}

// MSB-first bit reader over the entropy-coded segment: removes FF 00 stuffing, stops at any other marker (and then
// feeds zero bits, like jdhuff.c after "Corrupt JPEG data: premature end of data segment").
struct BitReader {
  const uint8_t* p;
  const uint8_t* end;
  uint64_t buf;  // bits are consumed from the top
  int cnt;
  int hit_marker;  // marker byte seen (0 = none)
};

TISE_JHD void br_init(BitReader& r, const uint8_t* p, int64_t n) {
  r.p = p;
  r.end = p + n;
  r.buf = 0;
  r.cnt = 0;
  r.hit_marker = 0;
}
TISE_JHD void br_fill(BitReader& r) {  // at least 32 bits afterwards (zeros once the data has ended)
  while (r.cnt <= 56) {
    uint32_t b = 0;
    if (!r.hit_marker && r.p < r.end) {
      b = *r.p;
      if (b == 0xFF) {
        const uint32_t b2 = (r.p + 1 < r.end) ? r.p[1] : 0xD9u;
        if (b2 == 0) {
          r.p += 2;
        } else {
          r.hit_marker = (int)b2;
          b = 0;
        }
      } else {
        r.p += 1;
      }
    }
    r.buf |= (uint64_t)b << (56 - r.cnt);
    r.cnt += 8;
  }
}
TISE_JHD uint32_t br_peek(const BitReader& r, int n) { return (uint32_t)(r.buf >> (64 - n)); }
TISE_JHD void br_skip(BitReader& r, int n) {
  r.buf <<= n;
  r.cnt -= n;
}
TISE_JHD int br_get(BitReader& r, int n) {
  if (n == 0) return 0;
  const int v = (int)(r.buf >> (64 - n));
  r.buf <<= n;
  r.cnt -= n;
  return v;
}
TISE_JHD int huff_extend(int r, int s) { return r < (1 << (s - 1)) ? r - (1 << s) + 1 : r; }

TISE_JHD int huff_decode(const DTable* t, BitReader& r) {  // needs >= 16 buffered bits; -1 = no such code
  const uint32_t e = t->look[br_peek(r, 8)];
  if (e) {
    br_skip(r, (int)(e >> 8));
    return (int)(e & 255u);
  }
  int l = 9;
  int code = (int)br_peek(r, 9);
  while (l <= 16 && code > t->maxcode[l]) {
    ++l;
    code = (int)br_peek(r, l);
  }
  if (l > 16) return -1;
  br_skip(r, l);
  return t->vals[(code + t->valoffset[l]) & 255];
}

// Entropy decoding of all scans of one image (jdhuff.c decode_mcu for sequential files, jdphuff.c decode_mcu_DC_first /
// DC_refine / AC_first / AC_refine for progressive ones).  coef: per component a plane of blocks (blocks_w x blocks_h)
// x 64 int16 in natural order, zero-initialised by the caller; comp_off[c] = first int16 of component c.  work: 6
// tables of scratch (3 DC + 3 AC, rebuilt for every scan).  Runs on ONE thread.
struct ScanState {
  BitReader r;
  int pred[3];
  int eobrun;
  int restarts_left, next_rst;
};

TISE_JHD int scan_restart(ScanState& st, int interval) {
  BitReader& r = st.r;
  br_skip(r, r.cnt & 7);
  br_fill(r);
  if (r.hit_marker != 0xD0 + st.next_rst) return kBadRestart;
  r.p += 2;  // the reader stopped AT the marker: step over it and restart the bit buffer
  r.buf = 0;
  r.cnt = 0;
  r.hit_marker = 0;
  st.next_rst = (st.next_rst + 1) & 7;
  st.restarts_left = interval;
  st.pred[0] = st.pred[1] = st.pred[2] = 0;
  st.eobrun = 0;
  return kOk;
}

TISE_JHD int block_sequential(ScanState& st, int c, const DTable* dct, const DTable* act, int16_t* blk) {
  BitReader& r = st.r;
  br_fill(r);
  int s = huff_decode(dct, r);
  if (s < 0) return kBadHuffmanCode;
  if (s > 15) return kBadCoefficient;
  if (s) {
    br_fill(r);
    s = huff_extend(br_get(r, s), s);
  }
  st.pred[c] += s;
  blk[0] = (int16_t)st.pred[c];
  for (int kk = 1; kk < 64; ++kk) {
    br_fill(r);
    const int rs = huff_decode(act, r);
    if (rs < 0) return kBadHuffmanCode;
    const int run = rs >> 4;
    s = rs & 15;
    if (s) {
      kk += run;
      if (kk > 63) return kBadCoefficient;
      blk[natural_order(kk)] = (int16_t)huff_extend(br_get(r, s), s);
    } else {
      if (run != 15) break;
      kk += 15;
    }
  }
  return kOk;
}

TISE_JHD int block_dc_first(ScanState& st, int c, const DTable* dct, int al, int16_t* blk) {
  BitReader& r = st.r;
  br_fill(r);
  int s = huff_decode(dct, r);
  if (s < 0) return kBadHuffmanCode;
  if (s > 15) return kBadCoefficient;
  if (s) {
    br_fill(r);
    s = huff_extend(br_get(r, s), s);
  }
  st.pred[c] += s;
  blk[0] = (int16_t)(st.pred[c] * (1 << al));
  return kOk;
}

TISE_JHD int block_ac_first(ScanState& st, const DTable* act, int ss, int se, int al, int16_t* blk) {
  if (st.eobrun > 0) {
    st.eobrun--;
    return kOk;
  }
  BitReader& r = st.r;
  for (int k = ss; k <= se; ++k) {
    br_fill(r);
    const int rs = huff_decode(act, r);
    if (rs < 0) return kBadHuffmanCode;
    const int run = rs >> 4, s = rs & 15;
    if (s) {
      k += run;
      if (k > 63) return kBadCoefficient;
      blk[natural_order(k)] = (int16_t)(huff_extend(br_get(r, s), s) * (1 << al));
    } else if (run == 15) {
      k += 15;
    } else {
      st.eobrun = 1 << run;
      if (run) st.eobrun += br_get(r, run);
      st.eobrun--;
      break;
    }
  }
  return kOk;
}

TISE_JHD void refine_nonzero(BitReader& r, int16_t* coef, int p1, int m1) {
  br_fill(r);
  if (br_get(r, 1)) {
    if ((*coef & p1) == 0) *coef = (int16_t)(*coef >= 0 ? *coef + p1 : *coef + m1);
  }
}

TISE_JHD int block_ac_refine(ScanState& st, const DTable* act, int ss, int se, int al, int16_t* blk) {
  BitReader& r = st.r;
  const int p1 = 1 << al, m1 = -(1 << al);
  int k = ss;
  if (st.eobrun == 0) {
    for (; k <= se; ++k) {
      br_fill(r);
      const int rs = huff_decode(act, r);
      if (rs < 0) return kBadHuffmanCode;
      int run = rs >> 4, s = rs & 15;
      if (s) {  // a newly non-zero coefficient: its sign follows (size is always 1)
        br_fill(r);
        s = br_get(r, 1) ? p1 : m1;
      } else if (run != 15) {
        st.eobrun = 1 << run;
        if (run) {
          br_fill(r);
          st.eobrun += br_get(r, run);
        }
        break;  // end of band
      }
      // advance over already-non-zero coefficients and `run` still-zero ones, appending correction bits to the former
      do {
        int16_t* cf = blk + natural_order(k);
        if (*cf != 0) {
          refine_nonzero(r, cf, p1, m1);
        } else {
          if (--run < 0) break;
        }
        ++k;
      } while (k <= se);
      if (s) {
        if (k > 63) return kBadCoefficient;
        blk[natural_order(k)] = (int16_t)s;
      }
    }
  }
  if (st.eobrun > 0) {
    for (; k <= se; ++k) {
      int16_t* cf = blk + natural_order(k);
      if (*cf != 0) refine_nonzero(r, cf, p1, m1);
    }
    st.eobrun--;
  }
  return kOk;
}

TISE_JHD int decode_image(const Core* h, const Scan* scans, const HuffDef* tables, const uint8_t* file, DTable* work,
                          int16_t* coef, const int64_t* comp_off) {
  for (int si = 0; si < h->nscans; ++si) {
    const Scan& sc = scans[si];
    const bool prog = h->progressive != 0;
    const bool dc_scan = !prog || sc.ss == 0, ac_scan = !prog || sc.ss > 0;
    for (int c = 0; c < sc.ncomp; ++c) {
      if (dc_scan && (!prog || sc.ah == 0)) build_dtable(tables[sc.dc_tab[c]].bits, tables[sc.dc_tab[c]].vals, work + c);
      if (ac_scan) build_dtable(tables[sc.ac_tab[c]].bits, tables[sc.ac_tab[c]].vals, work + 3 + c);
    }
    ScanState st;
    br_init(st.r, file + sc.off, sc.bytes);
    st.pred[0] = st.pred[1] = st.pred[2] = 0;
    st.eobrun = 0;
    st.restarts_left = sc.restart_interval;
    st.next_rst = 0;
    // MCU grid of this scan: interleaved scans walk the frame's MCUs, a single-component scan walks that component's own
    // blocks (ceil(samples / 8): the MCU padding of the frame is not coded)
    const Component& k0 = h->comp[sc.comp[0]];
    const int gx = sc.ncomp > 1 ? h->mcus_x : (k0.ds_w + 7) / 8, gy = sc.ncomp > 1 ? h->mcus_y : (k0.ds_h + 7) / 8;
    for (int my = 0; my < gy; ++my)
      for (int mx = 0; mx < gx; ++mx) {
        if (sc.restart_interval) {
          if (st.restarts_left == 0) {
            const int e = scan_restart(st, sc.restart_interval);
            if (e) return e;
          }
          --st.restarts_left;
        }
        for (int c = 0; c < sc.ncomp; ++c) {
          const int ci = sc.comp[c];
          const Component& k = h->comp[ci];
          const int bh = sc.ncomp > 1 ? k.h : 1, bv = sc.ncomp > 1 ? k.v : 1;
          for (int by = 0; by < bv; ++by)
            for (int bx = 0; bx < bh; ++bx) {
              int16_t* blk = coef + comp_off[ci] + ((int64_t)(my * bv + by) * k.blocks_w + (mx * bh + bx)) * 64;
              int e;
              if (!prog) {
                e = block_sequential(st, ci, work + c, work + 3 + c, blk);
              } else if (sc.ss == 0) {
                if (sc.ah == 0) {
                  e = block_dc_first(st, ci, work + c, sc.al, blk);
                } else {
                  br_fill(st.r);
                  if (br_get(st.r, 1)) blk[0] = (int16_t)(blk[0] | (1 << sc.al));
                  e = kOk;
                }
              } else {
                e = sc.ah == 0 ? block_ac_first(st, work + 3 + c, sc.ss, sc.se, sc.al, blk)
                               : block_ac_refine(st, work + 3 + c, sc.ss, sc.se, sc.al, blk);
              }
              if (e) return e;
            }
        }
      }
    const int hm = st.r.hit_marker;  // entropy data that ends early shows up as an unexpected marker before the last MCU
    if (hm && hm != 0xD9 && hm != 0xDA && hm != 0xC4 && hm != 0xDD && !(hm >= 0xD0 && hm <= 0xD7) && hm != 0xFE && !(hm >= 0xE0 && hm <= 0xEF))
      return kTruncated;
  }
  return kOk;
}

// ---------------------------------------------------------------------------------------------------------
// jidctint.c jpeg_idct_islow, one 8 x 8 block: coef (natural order) x quant (natural order) -> samples
TISE_JHD uint8_t range_limit(int v) {
  const int idx = v & 1023;  // the idct range-limit table of jdmaster.c prepare_range_limit_table, centred by +128
  return (uint8_t)(idx < 128 ? idx + 128 : idx < 512 ? 255 : idx < 896 ? 0 : idx - 896);
}

TISE_JHD void idct_islow(const int16_t* in, const uint16_t* q, uint8_t* out, int out_stride) {
  constexpr int CB = 13, P1 = 2;
  constexpr int F0298 = 2446, F0390 = 3196, F0541 = 4433, F0765 = 6270, F0899 = 7373, F1175 = 9633, F1501 = 12299,
                F1847 = 15137, F1961 = 16069, F2053 = 16819, F2562 = 20995, F3072 = 25172;
  typedef int64_t L;  // JLONG of jidctint.c: 64-bit, so damaged streams cannot overflow (results identical for valid data)
  L ws[64];
  for (int c = 0; c < 8; ++c) {
    const L d0 = (L)in[c] * q[c], d1 = (L)in[8 + c] * q[8 + c], d2 = (L)in[16 + c] * q[16 + c], d3 = (L)in[24 + c] * q[24 + c],
            d4 = (L)in[32 + c] * q[32 + c], d5 = (L)in[40 + c] * q[40 + c], d6 = (L)in[48 + c] * q[48 + c], d7 = (L)in[56 + c] * q[56 + c];
    L z2 = d2, z3 = d6;
    L z1 = (z2 + z3) * F0541;
    L tmp2 = z1 + z3 * (-F1847);
    L tmp3 = z1 + z2 * F0765;
    z2 = d0;
    z3 = d4;
    L tmp0 = (z2 + z3) * (1 << CB);
    L tmp1 = (z2 - z3) * (1 << CB);
    const L tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = d7;
    tmp1 = d5;
    tmp2 = d3;
    tmp3 = d1;
    z1 = tmp0 + tmp3;
    z2 = tmp1 + tmp2;
    z3 = tmp0 + tmp2;
    L z4 = tmp1 + tmp3;
    const L z5 = (z3 + z4) * F1175;
    tmp0 *= F0298;
    tmp1 *= F2053;
    tmp2 *= F3072;
    tmp3 *= F1501;
    z1 *= -F0899;
    z2 *= -F2562;
    z3 *= -F1961;
    z4 *= -F0390;
    z3 += z5;
    z4 += z5;
    tmp0 += z1 + z3;
    tmp1 += z2 + z4;
    tmp2 += z2 + z3;
    tmp3 += z1 + z4;
    constexpr int S = CB - P1, R = 1 << (S - 1);
    ws[c] = (tmp10 + tmp3 + R) >> S;
    ws[56 + c] = (tmp10 - tmp3 + R) >> S;
    ws[8 + c] = (tmp11 + tmp2 + R) >> S;
    ws[48 + c] = (tmp11 - tmp2 + R) >> S;
    ws[16 + c] = (tmp12 + tmp1 + R) >> S;
    ws[40 + c] = (tmp12 - tmp1 + R) >> S;
    ws[24 + c] = (tmp13 + tmp0 + R) >> S;
    ws[32 + c] = (tmp13 - tmp0 + R) >> S;
  }
  for (int rr = 0; rr < 8; ++rr) {
    const L* w = ws + rr * 8;
    L z2 = w[2], z3 = w[6];
    L z1 = (z2 + z3) * F0541;
    L tmp2 = z1 + z3 * (-F1847);
    L tmp3 = z1 + z2 * F0765;
    L tmp0 = (w[0] + w[4]) * (1 << CB);
    L tmp1 = (w[0] - w[4]) * (1 << CB);
    const L tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = w[7];
    tmp1 = w[5];
    tmp2 = w[3];
    tmp3 = w[1];
    z1 = tmp0 + tmp3;
    z2 = tmp1 + tmp2;
    z3 = tmp0 + tmp2;
    L z4 = tmp1 + tmp3;
    const L z5 = (z3 + z4) * F1175;
    tmp0 *= F0298;
    tmp1 *= F2053;
    tmp2 *= F3072;
    tmp3 *= F1501;
    z1 *= -F0899;
    z2 *= -F2562;
    z3 *= -F1961;
    z4 *= -F0390;
    z3 += z5;
    z4 += z5;
    tmp0 += z1 + z3;
    tmp1 += z2 + z4;
    tmp2 += z2 + z3;
    tmp3 += z1 + z4;
    constexpr int S = CB + P1 + 3, R = 1 << (S - 1);
    uint8_t* o = out + rr * out_stride;
    o[0] = range_limit((int)((tmp10 + tmp3 + R) >> S));
    o[7] = range_limit((int)((tmp10 - tmp3 + R) >> S));
    o[1] = range_limit((int)((tmp11 + tmp2 + R) >> S));
    o[6] = range_limit((int)((tmp11 - tmp2 + R) >> S));
    o[2] = range_limit((int)((tmp12 + tmp1 + R) >> S));
    o[5] = range_limit((int)((tmp12 - tmp1 + R) >> S));
    o[3] = range_limit((int)((tmp13 + tmp0 + R) >> S));
    o[4] = range_limit((int)((tmp13 - tmp0 + R) >> S));
  }
}

// ---------------------------------------------------------------------------------------------------------
// One output pixel: chroma upsampling (jdsample.c) + YCbCr -> RGB (jdcolor.c).  planes: the IDCT output of each
// component (pitch = blocks_w * 8).
TISE_JHD int clamp255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

TISE_JHD int chroma_sample(const uint8_t* pl, int pitch, int ds_w, int ds_h, int mode, int x, int y) {
  if (mode == 1) return pl[(int64_t)y * pitch + x];
  // jdsample.c jinit_upsampler: the triangle filters need more than two chroma columns, else plain replication
  if (ds_w <= 2) return pl[(int64_t)(mode == 3 ? (y >> 1) : y) * pitch + (x >> 1)];
  if (mode == 2) {  // h2v1_fancy_upsample
    const uint8_t* row = pl + (int64_t)y * pitch;
    const int i = x >> 1;
    const int cur = row[i];
    if (x & 1) return i == ds_w - 1 ? cur : (cur * 3 + row[i + 1] + 2) >> 2;
    return i == 0 ? cur : (cur * 3 + row[i - 1] + 1) >> 2;
  }
  // h2v2_fancy_upsample: vertical 3:1 blend with the nearer neighbour row (replicated at the image's top / bottom, as the
  // context rows of jdmainct.c are), then horizontal 3:1 with rounding 8 / 7
  const int j = y >> 1, i = x >> 1;
  int jn = (y & 1) ? j + 1 : j - 1;
  jn = jn < 0 ? 0 : (jn > ds_h - 1 ? ds_h - 1 : jn);
  const uint8_t* r0 = pl + (int64_t)j * pitch;
  const uint8_t* r1 = pl + (int64_t)jn * pitch;
  const int cur = r0[i] * 3 + r1[i];
  if (x & 1) {
    if (i == ds_w - 1) return (cur * 4 + 7) >> 4;
    return (cur * 3 + (r0[i + 1] * 3 + r1[i + 1]) + 7) >> 4;
  }
  if (i == 0) return (cur * 4 + 8) >> 4;
  return (cur * 3 + (r0[i - 1] * 3 + r1[i - 1]) + 8) >> 4;
}

TISE_JHD void ycc_to_rgb(int y, int cb, int cr, uint8_t* rgb) {
  constexpr int ONE_HALF = 1 << 15;
  const int xb = cb - 128, xr = cr - 128;
  const int r_add = (91881 * xr + ONE_HALF) >> 16;            // FIX(1.40200)
  const int b_add = (116130 * xb + ONE_HALF) >> 16;           // FIX(1.77200)
  const int g_add = (-22554 * xb + ONE_HALF - 46802 * xr) >> 16;  // FIX(0.34414), FIX(0.71414)
  rgb[0] = (uint8_t)clamp255(y + r_add);
  rgb[1] = (uint8_t)clamp255(y + g_add);
  rgb[2] = (uint8_t)clamp255(y + b_add);
}

TISE_JHD void output_pixel(const Core* h, const uint8_t* const* planes, int x, int y, uint8_t* rgb) {
  const int py = h->comp[0].blocks_w * 8;
  const int yy = planes[0][(int64_t)y * py + x];
  if (h->ncomp == 1) {
    rgb[0] = rgb[1] = rgb[2] = (uint8_t)yy;
    return;
  }
  const int pc = h->comp[1].blocks_w * 8;
  const int cb = chroma_sample(planes[1], pc, h->comp[1].ds_w, h->comp[1].ds_h, h->chroma_mode, x, y);
  const int cr = chroma_sample(planes[2], pc, h->comp[2].ds_w, h->comp[2].ds_h, h->chroma_mode, x, y);
  ycc_to_rgb(yy, cb, cr, rgb);
}

}  // namespace jpg
}  // namespace tise
