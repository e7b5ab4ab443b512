// A1: PNG decode building blocks (zlib inflate, scanline unfiltering, conversion to RGB8) written once for
// the device and the host.  On the device one WARP decodes one image: every lane runs the same bit reader and
// Huffman decoder on registers (warp-uniform control flow, no divergence), lane 0 writes literals, all lanes
// copy LZ77 matches and build the lookup tables; `LANES` = 32 there.  Compiled with a plain C++ compiler
// (LANES = 1, no barriers) the very same functions decode on the CPU -- that build exists ONLY so that the
// CPU test-suite can check the decoder logic against zlib / Pillow without a GPU (tests/png_host.cpp);
// the library itself never decodes on the host.
//
// Replaces PIL.Image.open(f).convert("RGB") at image_realism/FID/img_data.py:21,
// object_fidelity/O-IS/object_centric_inception_score.py:101 for PNG files (every colour type; 8-bit, 16-bit except grey,
// 1/2/4-bit grey and palette; plain or Adam7); anything else is reported as unsupported by parse() and stays on the host.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define TISE_HD __host__ __device__ __forceinline__
#else
#define TISE_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define TISE_SYNCWARP() __syncwarp()
#else
#define TISE_SYNCWARP() ((void)0)
#endif

namespace tise {
namespace png {

enum Status : int32_t {
  kOk = 0,
  kBadZlibHeader = 1,
  kBadBlockType = 2,
  kBadCodeLengths = 3,  // over-subscribed Huffman code, repeat without previous length, no end-of-block code
  kBadSymbol = 4,       // bit pattern that is no code, length symbol > 285, distance symbol > 29
  kBadDistance = 5,     // match reaches before the start of the output
  kOutputOverflow = 6,  // more data than height * (1 + stride)
  kStreamOverrun = 7,   // compressed stream ended early
  kSizeMismatch = 8,    // fewer bytes than height * (1 + stride)
  kBadFilter = 9,       // scanline filter type > 4
  kBadChecksum = 10,    // Adler-32 of the inflated bytes differs from the zlib trailer
  kStoredLength = 11,   // stored block LEN != ~NLEN
  kNotDecoded = 12      // never written: the kernel did not reach this image
};

// ---------------------------------------------------------------------------------------------------------
// Bit reader over a 4-byte aligned little-endian word stream.  All state is per lane, in registers.
struct BitReader {
  const uint32_t* words;
  uint32_t n_words;  // words that may be read (the last one may be partial: its tail bytes are never consumed)
  uint32_t pos;      // next word
  uint64_t buf;
  int cnt;
};

TISE_HD void br_init(BitReader& r, const uint32_t* words, uint32_t stream_bytes) {
  r.words = words;
  r.n_words = (stream_bytes + 3u) >> 2;
  r.pos = 0;
  r.buf = 0;
  r.cnt = 0;
}
// after a refill at least 33 bits are buffered (zeros past the end; br_overrun() tells)
TISE_HD void br_refill(BitReader& r) {
  if (r.cnt <= 32) {
    const uint32_t w = r.pos < r.n_words ? r.words[r.pos] : 0u;
    r.pos++;
    r.buf |= (uint64_t)w << r.cnt;
    r.cnt += 32;
  }
}
TISE_HD uint32_t br_bits(BitReader& r, int n) {
  const uint32_t v = (uint32_t)r.buf & ((1u << n) - 1u);
  r.buf >>= n;
  r.cnt -= n;
  return v;
}
// byte offset in the stream of the next unread bit (rounded down)
TISE_HD int64_t br_byte_pos(const BitReader& r) { return (int64_t)r.pos * 4 - (r.cnt >> 3); }
TISE_HD void br_seek(BitReader& r, uint32_t byte_pos) {
  r.pos = byte_pos >> 2;
  r.buf = 0;
  r.cnt = 0;
  br_refill(r);
  const int drop = (int)(byte_pos & 3u) * 8;
  r.buf >>= drop;
  r.cnt -= drop;
}

// ---------------------------------------------------------------------------------------------------------
// Canonical Huffman code of up to 288 symbols: a 10-bit direct lookup table and the count / sorted-symbol
// arrays of the classic bit-at-a-time canonical decoder for the (rare) longer codes.  Lives in shared memory.
constexpr int kFastBits = 10;
constexpr int kFastSize = 1 << kFastBits;
constexpr int kMaxSyms = 288;
struct Huff {
  uint16_t fast[kFastSize];  // (length << 9) | symbol, 0 = no code of <= kFastBits bits starts like this
  uint16_t count[16];        // codes per length
  uint16_t sorted[kMaxSyms]; // symbols ordered by (length, symbol)
  uint16_t code[kMaxSyms];   // canonical code of each symbol (scratch of the build)
  int32_t err;
};

TISE_HD uint32_t bit_reverse16(uint32_t v, int len) {
  v = ((v & 0xAAAAu) >> 1) | ((v & 0x5555u) << 1);
  v = ((v & 0xCCCCu) >> 2) | ((v & 0x3333u) << 2);
  v = ((v & 0xF0F0u) >> 4) | ((v & 0x0F0Fu) << 4);
  v = ((v & 0xFF00u) >> 8) | ((v & 0x00FFu) << 8);
  return v >> (16 - len);
}

// lens[0..n): code length of each symbol (0 = unused).  Lane 0 does the serial part, all lanes fill the table.
template <int LANES>
TISE_HD int huff_build(Huff* h, const uint8_t* lens, int n, int lane) {
  TISE_SYNCWARP();  // lens[] (written by lane 0) and earlier readers of *h are done
  if (lane == 0) {
    uint16_t cnt[16];
    for (int l = 0; l < 16; ++l) cnt[l] = 0;
    for (int s = 0; s < n; ++s) cnt[lens[s]]++;
    cnt[0] = 0;
    int left = 1, err = kOk;
    for (int l = 1; l < 16; ++l) {
      left = left * 2 - (int)cnt[l];
      if (left < 0) {  // over-subscribed
        err = kBadCodeLengths;
        break;
      }
    }
    uint16_t next_code[16], next_idx[16];
    uint32_t c = 0, idx = 0;
    for (int l = 1; l < 16; ++l) {
      next_code[l] = (uint16_t)c;
      next_idx[l] = (uint16_t)idx;
      c = (c + cnt[l]) << 1;
      idx += cnt[l];
    }
    if (!err)
      for (int s = 0; s < n; ++s) {
        const int l = lens[s];
        if (l) {
          h->sorted[next_idx[l]++] = (uint16_t)s;
          h->code[s] = next_code[l]++;
        }
      }
    for (int l = 0; l < 16; ++l) h->count[l] = cnt[l];
    h->err = err;
  }
  for (int i = lane; i < kFastSize; i += LANES) h->fast[i] = 0;
  TISE_SYNCWARP();
  const int err = h->err;
  if (!err)
    for (int s = lane; s < n; s += LANES) {
      const int l = lens[s];
      if (l && l <= kFastBits) {
        const uint16_t e = (uint16_t)((l << 9) | s);
        for (uint32_t j = bit_reverse16(h->code[s], l); j < (uint32_t)kFastSize; j += 1u << l) h->fast[j] = e;
      }
    }
  TISE_SYNCWARP();
  return err;
}

// needs >= 15 buffered bits; returns the symbol or -1
TISE_HD int huff_decode(const Huff* h, BitReader& r) {
  const uint32_t e = h->fast[(uint32_t)r.buf & (kFastSize - 1)];
  if (e) {
    const int l = (int)(e >> 9);
    r.buf >>= l;
    r.cnt -= l;
    return (int)(e & 511u);
  }
  int code = 0, first = 0, index = 0;
  for (int l = 1; l < 16; ++l) {
    code |= (int)(r.buf & 1u);
    r.buf >>= 1;
    r.cnt -= 1;
    const int count = h->count[l];
    if (code - count < first) return h->sorted[index + (code - first)];
    index += count;
    first += count;
    first <<= 1;
    code <<= 1;
  }
  return -1;
}

// ---------------------------------------------------------------------------------------------------------
// zlib (RFC 1950) stream -> out[0 .. out_cap).  lens: 320 bytes, win: 32 x uint16 of scratch next to the two tables.
// Returns a Status; *out_len = bytes produced, *adler = the stream's Adler-32 trailer.
template <int LANES>
TISE_HD int inflate_zlib(const uint32_t* words, uint32_t stream_bytes, uint8_t* out, uint32_t out_cap, uint32_t* out_len,
                         uint32_t* adler, Huff* hl, Huff* hd, uint8_t* lens, uint16_t* win, int lane) {
  (void)win;  // 32 entries of per-warp scratch for the literal-run decoder (device build only)
  BitReader r;
  br_init(r, words, stream_bytes);
  *out_len = 0;
  *adler = 0;
  if (stream_bytes < 6) return kStreamOverrun;
  br_refill(r);
  const uint32_t cmf = br_bits(r, 8), flg = br_bits(r, 8);
  if ((cmf & 15u) != 8u || (cmf >> 4) > 7u || ((cmf << 8) | flg) % 31u != 0u || (flg & 32u)) return kBadZlibHeader;

  uint32_t op = 0;
  int last = 0;
  while (!last) {
    br_refill(r);
    last = (int)br_bits(r, 1);
    const uint32_t type = br_bits(r, 2);
    if (type == 0) {  // stored
      br_bits(r, r.cnt & 7);
      br_refill(r);
      const uint32_t len = br_bits(r, 16), nlen = br_bits(r, 16);
      if ((len ^ 0xFFFFu) != nlen) return kStoredLength;
      const int64_t bp = br_byte_pos(r);
      if (bp + len > (int64_t)stream_bytes) return kStreamOverrun;
      if (op + len > out_cap) return kOutputOverflow;
      const uint8_t* src = reinterpret_cast<const uint8_t*>(words) + bp;
      for (uint32_t i = lane; i < len; i += LANES) out[op + i] = src[i];
      op += len;
      br_seek(r, (uint32_t)(bp + len));
      continue;
    }
    if (type == 3) return kBadBlockType;
    int n_lit, n_dist;
    if (type == 1) {  // fixed code
      n_lit = 288;
      n_dist = 32;
      TISE_SYNCWARP();
      for (int i = lane; i < 320; i += LANES)
        lens[i] = (uint8_t)(i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : i < 288 ? 8 : 5);
    } else {  // dynamic code
      n_lit = (int)br_bits(r, 5) + 257;
      n_dist = (int)br_bits(r, 5) + 1;
      const int n_clen = (int)br_bits(r, 4) + 4;
      if (n_lit > 286 || n_dist > 30) return kBadCodeLengths;
      TISE_SYNCWARP();
      if (lane == 0)
        for (int i = 0; i < 19; ++i) lens[i] = 0;
      for (int i = 0; i < n_clen; ++i) {
        br_refill(r);
        const uint32_t v = br_bits(r, 3);
        // order 16 17 18 0 8 7 9 6 10 5 11 4 12 3 13 2 14 1 15
        const int slot = i < 3 ? 16 + i : i == 3 ? 0 : (i & 1) ? (19 - i) / 2 : (12 + i) / 2;
        if (lane == 0) lens[slot] = (uint8_t)v;
      }
      int e = huff_build<LANES>(hd, lens, 19, lane);  // the code-length code borrows the distance table
      if (e) return e;
      int i = 0, prev = -1;
      const int total = n_lit + n_dist;
      while (i < total) {
        br_refill(r);
        const int sym = huff_decode(hd, r);
        if (sym < 0) return kBadSymbol;
        if (sym < 16) {
          if (lane == 0) lens[i] = (uint8_t)sym;
          prev = sym;
          ++i;
          continue;
        }
        int rep, val = 0;
        if (sym == 16) {
          if (prev < 0) return kBadCodeLengths;
          val = prev;
          rep = 3 + (int)br_bits(r, 2);
        } else if (sym == 17) {
          rep = 3 + (int)br_bits(r, 3);
        } else {
          rep = 11 + (int)br_bits(r, 7);
        }
        if (i + rep > total) return kBadCodeLengths;
        if (lane == 0)
          for (int k = 0; k < rep; ++k) lens[i + k] = (uint8_t)val;
        i += rep;
        prev = val;
      }
      TISE_SYNCWARP();
      if (lens[256] == 0) return kBadCodeLengths;
    }
    int e = huff_build<LANES>(hl, lens, n_lit, lane);
    if (e) return e;
    e = huff_build<LANES>(hd, lens + n_lit, n_dist, lane);
    if (e) return e;

    for (;;) {
      br_refill(r);
      // Literal runs, the bulk of a photographic PNG: lane k looks up the code that would start at bit offset k of the
      // buffered window (one shared-memory access for 32 candidate positions); the chain of actual code boundaries
      // 0 -> len_0 -> len_0 + len_1 ... is then followed through registers (one shuffle per symbol) instead of one
      // dependent table access, 64-bit shift and store per symbol, and the literals found are stored side by side.
      // The chain stops at the first code that is not a <= 10-bit literal or does not lie inside the buffered bits;
      // that symbol takes the general path below.
      {
        int p = 0, nlit = 0;
#if defined(__CUDA_ARCH__)
        // lane k: (bit position after the code starting at k) << 8 | literal, or 0xff00 if no <= 10-bit literal code
        // starts at k / it does not end inside the buffered bits.  The 32 candidates go to shared memory so that a hop
        // is one broadcast load instead of a warp-synchronous shuffle.
        const uint32_t mine = hl->fast[(uint32_t)(r.buf >> lane) & (kFastSize - 1)];
        const int l_mine = (int)(mine >> 9);
        const bool ok = mine != 0u && !(mine & 256u) && lane + l_mine <= r.cnt;
        TISE_SYNCWARP();  // the previous window's readers are done
        win[lane] = ok ? (uint16_t)(((lane + l_mine) << 8) | (mine & 255u)) : (uint16_t)0xff00;
        TISE_SYNCWARP();
        uint32_t my_sym = 0;
        while (p < 32) {
          const uint32_t e = win[p];
          const int np = (int)(e >> 8);
          if (np == 0xff) break;
          if (lane == nlit) my_sym = e;
          ++nlit;
          p = np;
        }
        if (nlit) {
          if (op + (uint32_t)nlit > out_cap) return kOutputOverflow;
          if (lane < nlit) out[op + lane] = (uint8_t)my_sym;
        }
#else
        uint8_t syms[32];
        while (p < 32) {
          const uint32_t e = hl->fast[(uint32_t)(r.buf >> p) & (kFastSize - 1)];
          const int l = (int)(e >> 9);
          if (e == 0u || (e & 256u) || p + l > r.cnt) break;
          syms[nlit++] = (uint8_t)(e & 255u);
          p += l;
        }
        if (nlit) {
          if (op + (uint32_t)nlit > out_cap) return kOutputOverflow;
          for (int k = 0; k < nlit; ++k) out[op + k] = syms[k];
        }
#endif
        if (nlit) {
          op += (uint32_t)nlit;
          r.buf >>= p;
          r.cnt -= p;
          continue;
        }
      }
      int sym = huff_decode(hl, r);
      if (sym < 256) {
        if (sym < 0) return kBadSymbol;
        if (op >= out_cap) return kOutputOverflow;
        if (lane == 0) out[op] = (uint8_t)sym;
        ++op;
        continue;
      }
      if (sym == 256) break;
      sym -= 257;
      if (sym >= 29) return kBadSymbol;
      uint32_t length;
      if (sym < 8) {
        length = 3u + sym;
      } else if (sym == 28) {
        length = 258u;
      } else {
        const int xb = (sym - 4) >> 2;
        length = ((4u + (sym & 3u)) << xb) + 3u + br_bits(r, xb);
      }
      br_refill(r);
      const int ds = huff_decode(hd, r);
      if (ds < 0 || ds >= 30) return kBadSymbol;
      uint32_t dist;
      if (ds < 4) {
        dist = 1u + ds;
      } else {
        const int xb = (ds >> 1) - 1;
        dist = ((2u + (ds & 1u)) << xb) + 1u + br_bits(r, xb);
      }
      if (dist > op) return kBadDistance;
      if (op + length > out_cap) return kOutputOverflow;
      TISE_SYNCWARP();  // every earlier store of every lane is visible
      const uint8_t* from = out + op - dist;
      for (uint32_t i = lane; i < length; i += LANES) out[op + i] = from[i < dist ? i : i % dist];
      op += length;
    }
    if (br_byte_pos(r) > (int64_t)stream_bytes + 4) return kStreamOverrun;
  }
  // trailer: Adler-32, big endian, byte aligned
  br_bits(r, r.cnt & 7);
  uint32_t a = 0;
  for (int k = 0; k < 4; ++k) {
    br_refill(r);
    a = (a << 8) | br_bits(r, 8);
  }
  if (br_byte_pos(r) > (int64_t)stream_bytes) return kStreamOverrun;
  *adler = a;
  *out_len = op;
  TISE_SYNCWARP();
  return kOk;
}

// ---------------------------------------------------------------------------------------------------------
// What the host parser extracts from a PNG file (offsets are relative to the first byte of the file).
constexpr int kMaxIdat = 1 << 20;
TISE_HD int samples_of(int color_type) {
  return color_type == 0 ? 1 : color_type == 2 ? 3 : color_type == 3 ? 1 : color_type == 4 ? 2 : color_type == 6 ? 4 : 0;
}
// bytes of one scanline (without its filter byte) and the filter's "bytes per pixel" unit (PNG spec 9.2)
TISE_HD int stride_of(int width, int color_type, int bit_depth) {
  return (int)(((int64_t)width * samples_of(color_type) * bit_depth + 7) >> 3);
}
TISE_HD int filter_bpp_of(int color_type, int bit_depth) {
  const int b = samples_of(color_type) * bit_depth / 8;
  return b < 1 ? 1 : b;
}
// What convert("RGB") of Pillow's decoder is reproduced for: 8-bit everything; 16-bit RGB / RGBA / grey+alpha (Pillow
// keeps the high byte); 1/2/4-bit grey (scaled to 0..255) and palette.  16-bit grey opens as "I;16", whose conversion
// to RGB saturates instead of scaling -- left to the host.  (Adam7-interlaced files of the supported depths are decoded
// on the device: seven filtered sub-images scattered into place.)
TISE_HD bool depth_supported(int color_type, int bit_depth) {
  if (bit_depth == 8) return samples_of(color_type) > 0;
  if (bit_depth == 16) return color_type == 2 || color_type == 4 || color_type == 6;
  if (bit_depth == 1 || bit_depth == 2 || bit_depth == 4) return color_type == 0 || color_type == 3;
  return false;
}

struct Header {
  int32_t width, height, bit_depth, color_type, interlace;
  int32_t channels;  // filter unit: bytes per pixel of a scanline, at least 1
  int32_t n_idat;
  int32_t supported;
  int64_t idat_bytes;
  int32_t plte_off, plte_entries;
};

inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

// Walks the chunk list.  on_idat(offset_of_data, bytes) is called for every IDAT chunk.  Returns 0, or -1 if the file is
// not a well-formed PNG container (bad signature, truncated chunk, no IHDR / IDAT / IEND).
template <class F>
inline int parse(const uint8_t* f, int64_t n, Header* h, F on_idat) {
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
  *h = Header();
  h->plte_off = -1;
  if (n < 8 + 25) return -1;
  for (int i = 0; i < 8; ++i)
    if (f[i] != sig[i]) return -1;
  int64_t p = 8;
  bool have_ihdr = false, have_iend = false;
  while (p + 12 <= n) {
    const uint32_t len = be32(f + p);
    const uint8_t* t = f + p + 4;
    if ((int64_t)len > n - p - 12) return -1;
    const int64_t data = p + 8;
    if (!have_ihdr) {
      if (!(t[0] == 'I' && t[1] == 'H' && t[2] == 'D' && t[3] == 'R') || len != 13) return -1;
      h->width = (int32_t)be32(f + data);
      h->height = (int32_t)be32(f + data + 4);
      h->bit_depth = f[data + 8];
      h->color_type = f[data + 9];
      const int compression = f[data + 10], filter = f[data + 11];
      h->interlace = f[data + 12];
      const bool fmt_ok = depth_supported(h->color_type, h->bit_depth);
      h->channels = fmt_ok ? filter_bpp_of(h->color_type, h->bit_depth) : 0;
      h->supported = (fmt_ok && h->interlace <= 1 && compression == 0 && filter == 0 &&
                      h->width > 0 && h->height > 0 && (int64_t)h->width * h->height <= (int64_t)1 << 28)
                         ? 1
                         : 0;
      have_ihdr = true;
    } else if (t[0] == 'I' && t[1] == 'D' && t[2] == 'A' && t[3] == 'T') {
      if (len) {
        on_idat(data, (int64_t)len);
        h->n_idat++;
        h->idat_bytes += len;
      }
    } else if (t[0] == 'P' && t[1] == 'L' && t[2] == 'T' && t[3] == 'E') {
      h->plte_off = (int32_t)data;
      h->plte_entries = (int32_t)(len / 3);
    } else if (t[0] == 'I' && t[1] == 'E' && t[2] == 'N' && t[3] == 'D') {
      have_iend = true;
      break;
    }
    p = data + len + 4;
  }
  if (!have_ihdr || !have_iend || h->n_idat == 0) return -1;
  if (h->color_type == 3 && (h->plte_off < 0 || h->plte_entries == 0)) h->supported = 0;
  if (h->idat_bytes > (int64_t)0x7fffffff) h->supported = 0;
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Scanline unfiltering (PNG spec 9.2) + conversion to RGB8, one image per warp.
// raw: height rows of [filter type | stride bytes]; rows: 2 * stride bytes of scratch (shared memory);
// palette: PLTE bytes (colour type 3) or null.  Also accumulates the Adler-32 of raw for the zlib check.
TISE_HD int paeth(int a, int b, int c) {
  const int p = a + b - c;
  int pa = p - a, pb = p - b, pc = p - c;
  pa = pa < 0 ? -pa : pa;
  pb = pb < 0 ? -pb : pb;
  pc = pc < 0 ? -pc : pc;
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

#if defined(__CUDA_ARCH__)
TISE_HD uint64_t lanes_sum(uint64_t v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#else
TISE_HD uint64_t lanes_sum(uint64_t v) { return v; }
#endif

// One pass of scanlines (the whole image, or one of the seven Adam7 sub-images): unfilter row by row and write pixel
// (r, c) of the pass to image position (y0 + r * ys, x0 + c * xs).  A / B carry the Adler-32 state across passes.
template <int LANES>
TISE_HD int unfilter_pass(const uint8_t* raw, int pw, int ph, int bit_depth, int color_type, const uint8_t* palette,
                          int palette_entries, uint8_t* rows, uint8_t* out, int out_width, int x0, int xs, int y0, int ys,
                          uint64_t& A, uint64_t& B, int lane) {
  const int stride = stride_of(pw, color_type, bit_depth), bpp = filter_bpp_of(color_type, bit_depth);
  const int samples = samples_of(color_type);
  uint8_t* cur = rows;
  uint8_t* prev = rows + stride;
  TISE_SYNCWARP();
  for (int i = lane; i < stride; i += LANES) prev[i] = 0;
  const uint64_t m = (uint64_t)stride + 1;
  int status = kOk;
  for (int y = 0; y < ph; ++y) {
    const uint8_t* src = raw + (size_t)y * (stride + 1);
    const int ft = src[0];
    uint64_t s1 = 0, s2 = 0;
    if (lane == 0) {
      s1 = (uint64_t)ft;
      s2 = m * (uint64_t)ft;
    }
    TISE_SYNCWARP();  // the previous row's readers are done with `cur` (it was `prev` two rows ago)
    for (int i = lane; i < stride; i += LANES) {
      const uint32_t d = src[1 + i];
      cur[i] = (uint8_t)d;
      s1 += d;
      s2 += (m - 1 - (uint64_t)i) * d;
    }
    s1 = lanes_sum(s1);
    s2 = lanes_sum(s2);
    B = (B + m * A + s2) % 65521u;
    A = (A + s1) % 65521u;
    TISE_SYNCWARP();
    if (ft == 1) {
      if (lane < bpp) {
        int a = 0;
        for (int x = lane; x < stride; x += bpp) {
          a = (cur[x] + a) & 255;
          cur[x] = (uint8_t)a;
        }
      }
    } else if (ft == 2) {
      for (int i = lane; i < stride; i += LANES) cur[i] = (uint8_t)(cur[i] + prev[i]);
    } else if (ft == 3) {
      if (lane < bpp) {
        int a = 0;
        for (int x = lane; x < stride; x += bpp) {
          a = (cur[x] + ((a + prev[x]) >> 1)) & 255;
          cur[x] = (uint8_t)a;
        }
      }
    } else if (ft == 4) {
      if (lane < bpp) {
        int a = 0, c = 0;
        for (int x = lane; x < stride; x += bpp) {
          const int b = prev[x];
          a = (cur[x] + paeth(a, b, c)) & 255;
          cur[x] = (uint8_t)a;
          c = b;
        }
      }
    } else if (ft != 0) {
      status = kBadFilter;
    }
    if (LANES == 1 && bpp > 1 && ft != 0 && ft != 2) {
      // single-lane build: the loops above ran only chain 0; run the other byte chains of the pixel
      for (int ch = 1; ch < bpp; ++ch) {
        int a = 0, c = 0;
        for (int x = ch; x < stride; x += bpp) {
          const int b = prev[x];
          const int pred = ft == 1 ? a : ft == 3 ? ((a + b) >> 1) : paeth(a, b, c);
          a = (cur[x] + pred) & 255;
          cur[x] = (uint8_t)a;
          c = b;
        }
      }
    }
    TISE_SYNCWARP();
    uint8_t* o = out + ((size_t)(y0 + y * ys) * out_width + x0) * 3;
    const int ostep = xs * 3;  // bytes between consecutive pixels of this pass in the output row
    if (color_type == 2 && bit_depth == 8 && xs == 1) {
      for (int i = lane; i < stride; i += LANES) o[i] = cur[i];
    } else if (bit_depth < 8) {  // 1 / 2 / 4-bit grey (scaled to 0..255) or palette indices, packed MSB first
      const int mask = (1 << bit_depth) - 1, mul = 255 / mask;
      for (int i = lane; i < pw * 3; i += LANES) {
        const int p = i / 3, k = i - p * 3, bit = p * bit_depth;
        const int v = (cur[bit >> 3] >> (8 - bit_depth - (bit & 7))) & mask;
        o[p * ostep + k] = color_type == 3 ? (v < palette_entries ? palette[v * 3 + k] : (uint8_t)0) : (uint8_t)(v * mul);
      }
    } else if (color_type == 3) {
      for (int i = lane; i < pw * 3; i += LANES) {
        const int p = i / 3, k = i - p * 3, idx = cur[p];
        o[p * ostep + k] = idx < palette_entries ? palette[idx * 3 + k] : (uint8_t)0;
      }
    } else {  // grey (0), grey + alpha (4): replicate; RGB(A) (2, 6): first three samples; 16-bit: the high byte --
              // what Pillow's decoder + convert("RGB") produce
      const int sb = bit_depth >> 3;  // bytes per sample
      for (int i = lane; i < pw * 3; i += LANES) {
        const int p = i / 3, k = i - p * 3;
        o[p * ostep + k] = cur[(p * samples + (samples >= 3 ? k : 0)) * sb];
      }
    }
    uint8_t* t = cur;
    cur = prev;
    prev = t;
  }
  return status;
}

// Adam7 (PNG spec 8.2): pass p holds the pixels (ystart + r * ystep, xstart + c * xstep)
TISE_HD void adam7_pass(int p, int width, int height, int* x0, int* xs, int* y0, int* ys, int* pw, int* ph) {
  const int xst[7] = {0, 4, 0, 2, 0, 1, 0}, yst[7] = {0, 0, 4, 0, 2, 0, 1};
  const int xsp[7] = {8, 8, 4, 4, 2, 2, 1}, ysp[7] = {8, 8, 8, 4, 4, 2, 2};
  *x0 = xst[p];
  *y0 = yst[p];
  *xs = xsp[p];
  *ys = ysp[p];
  *pw = (width - xst[p] + xsp[p] - 1) / xsp[p];
  *ph = (height - yst[p] + ysp[p] - 1) / ysp[p];
  if (*pw < 0) *pw = 0;
  if (*ph < 0) *ph = 0;
}

// bytes of the inflated stream: filter byte + scanline per row, per pass when interlaced (empty passes are absent)
TISE_HD int64_t raw_bytes_of(int width, int height, int color_type, int bit_depth, int interlace) {
  if (!interlace) return (int64_t)height * (1 + (int64_t)stride_of(width, color_type, bit_depth));
  int64_t total = 0;
  for (int p = 0; p < 7; ++p) {
    int x0, xs, y0, ys, pw, ph;
    adam7_pass(p, width, height, &x0, &xs, &y0, &ys, &pw, &ph);
    if (pw > 0 && ph > 0) total += (int64_t)ph * (1 + (int64_t)stride_of(pw, color_type, bit_depth));
  }
  return total;
}

template <int LANES>
TISE_HD int unfilter_to_rgb(const uint8_t* raw, int width, int height, int bit_depth, int color_type, int interlace,
                            const uint8_t* palette, int palette_entries, uint8_t* rows, uint8_t* out,
                            uint32_t expect_adler, int lane) {
  uint64_t A = 1, B = 0;  // Adler-32 state, reduced once per row
  int status = kOk;
  if (!interlace) {
    status = unfilter_pass<LANES>(raw, width, height, bit_depth, color_type, palette, palette_entries, rows, out, width, 0, 1, 0,
                                  1, A, B, lane);
  } else {
    for (int p = 0; p < 7; ++p) {
      int x0, xs, y0, ys, pw, ph;
      adam7_pass(p, width, height, &x0, &xs, &y0, &ys, &pw, &ph);
      if (pw <= 0 || ph <= 0) continue;
      const int e = unfilter_pass<LANES>(raw, pw, ph, bit_depth, color_type, palette, palette_entries, rows, out, width, x0, xs,
                                         y0, ys, A, B, lane);
      if (e) status = e;
      raw += (size_t)ph * (1 + (size_t)stride_of(pw, color_type, bit_depth));
    }
  }
  if (status) return status;
  const uint32_t adler = (uint32_t)((B << 16) | A);
  return adler == expect_adler ? kOk : kBadChecksum;
}

}  // namespace png
}  // namespace tise
