"""CPU checks of the PNG ingest.  The device decoder's algorithms live in
tise_toolbox_b200/csrc/png_core.h, written for host and device alike; tests/png_host.cpp compiles that header
with g++ (one "lane") so the inflate / unfilter / chunk-parser LOGIC is checked here against zlib and Pillow
(the oracle for this row: the reference calls PIL.Image.open(f).convert("RGB"), img_data.py:21).
The GPU build of the same source is checked in test_gpu_png.py.  Nothing in the package uses this host build."""
import ctypes as C
import io
import os
import subprocess
import zlib

import numpy as np
import pytest
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def host():
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "png_host.so")
    src = os.path.join(HERE, "png_host.cpp")
    core = os.path.join(HERE, "..", "tise_toolbox_b200", "csrc", "png_core.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(core)):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", src, "-o", so])
    lib = C.CDLL(so)
    lib.host_inflate.argtypes = [C.c_char_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32),
                                 C.POINTER(C.c_uint32)]
    lib.host_png_decode.argtypes = [C.c_char_p, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_int32)]
    return lib


def _inflate(lib, z, cap):
    out = np.zeros(cap + 8, np.uint8)
    n, ad = C.c_uint32(), C.c_uint32()
    st = lib.host_inflate(z, len(z), out.ctypes.data, cap, C.byref(n), C.byref(ad))
    return st, out[:n.value].tobytes(), ad.value


def _decode(lib, blob, cap=1 << 24):
    hw = (C.c_int32 * 2)()
    out = np.zeros(cap, np.uint8)
    st = lib.host_png_decode(blob, len(blob), out.ctypes.data, out.size, hw)
    return st, (out[:hw[0] * hw[1] * 3].reshape(hw[0], hw[1], 3) if st == 0 else None)


def smooth_image(rng, h, w, c):
    """Low-frequency colour field + noise (like the synthetic sets): exercises matches AND literals."""
    gh, gw = max(h // 16, 1) + 1, max(w // 16, 1) + 1
    planes = [np.asarray(Image.fromarray(rng.integers(0, 256, (gh, gw)).astype(np.uint8)).resize((w, h), Image.BICUBIC))
              for _ in range(c)]
    a = np.stack(planes, -1).astype(np.int32) + rng.integers(-6, 7, (h, w, c))
    return np.clip(a, 0, 255).astype(np.uint8)


def png_bytes(arr_or_img, **kw):
    im = arr_or_img if isinstance(arr_or_img, Image.Image) else Image.fromarray(arr_or_img)
    b = io.BytesIO()
    im.save(b, "PNG", **kw)
    return b.getvalue()


@pytest.mark.parametrize("level", [0, 1, 6, 9])
def test_inflate_matches_zlib(host, level):
    rng = np.random.default_rng(level)
    for kind in range(6):
        n = int(rng.integers(0, 150000))
        if kind == 0:
            d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()  # incompressible -> stored blocks
        elif kind == 1:
            d = bytes(n)  # one long run: distance 1, length 258
        elif kind == 2:
            d = (b"abcabcabd" * (n // 9 + 1))[:n]
        elif kind == 3:
            d = rng.integers(0, 4, n, dtype=np.uint8).tobytes()
        elif kind == 4:
            d = np.cumsum(rng.integers(-2, 3, n)).astype(np.uint8).tobytes()
        else:
            d = b""
        z = zlib.compress(d, level)
        st, out, adler = _inflate(host, z, len(d))
        assert st == 0 and out == d and adler == zlib.adler32(d), (kind, n, st)


def test_inflate_fixed_huffman_and_long_codes(host):
    co = zlib.compressobj(6, zlib.DEFLATED, 15, 8, zlib.Z_FIXED)
    d = b"hello hello hello world" * 100
    z = co.compress(d) + co.flush()
    st, out, _ = _inflate(host, z, len(d))
    assert st == 0 and out == d
    # a very skewed literal distribution forces code lengths beyond the 10-bit lookup table (slow path)
    rng = np.random.default_rng(1)
    p = 0.5 ** np.arange(1, 41)
    p = np.concatenate([p, np.full(216, (1 - p.sum()) / 216)])
    d = rng.choice(256, size=300000, p=p / p.sum()).astype(np.uint8).tobytes()
    co = zlib.compressobj(9, zlib.DEFLATED, 15, 9, zlib.Z_HUFFMAN_ONLY)
    z = co.compress(d) + co.flush()
    st, out, _ = _inflate(host, z, len(d))
    assert st == 0 and out == d


def test_inflate_rejects_corrupt_streams(host):
    rng = np.random.default_rng(7)
    d = smooth_image(rng, 64, 64, 3).tobytes()
    z = bytearray(zlib.compress(d, 6))
    assert _inflate(host, bytes(z[:len(z) // 2]), len(d))[0] != 0  # truncated
    assert _inflate(host, bytes(z), len(d) - 100)[0] != 0  # output too small
    bad = bytearray(z)
    bad[0] = 0x79
    assert _inflate(host, bytes(bad), len(d))[0] != 0  # header
    n_detected = 0
    for k in range(200):  # random damage must never crash or overrun; most flips are detected
        bad = bytearray(z)
        i = int(rng.integers(2, len(bad) - 4))
        bad[i] ^= 1 << int(rng.integers(0, 8))
        st, out, adler = _inflate(host, bytes(bad), len(d))
        n_detected += int(st != 0 or out != d or adler != zlib.adler32(out))
    assert n_detected >= 190


MODES = [("RGB", 3), ("L", 1), ("RGBA", 4), ("LA", 2), ("P", 3)]
SIZES = [(256, 256), (1, 1), (17, 333), (64, 48), (500, 375), (3, 1000), (299, 299)]


@pytest.mark.parametrize("mode,c", MODES)
def test_png_decode_matches_pillow(host, mode, c):
    rng = np.random.default_rng(len(mode) * 31 + c)
    n_ok = 0
    for (h, w) in SIZES:
        for level in (1, 6, 9):
            a = smooth_image(rng, h, w, c)
            if mode == "P":
                im = Image.fromarray(a).convert("P", palette=Image.ADAPTIVE, colors=200)
            elif mode == "L":
                im = Image.fromarray(a[..., 0])
            else:
                im = Image.fromarray(a, mode)
            blob = png_bytes(im, compress_level=level)
            ref = np.asarray(Image.open(io.BytesIO(blob)).convert("RGB"))
            st, got = _decode(host, blob)
            if st == -2:  # Pillow packs tiny palettes at 1/2/4 bits: outside the device decoder, reported as such
                assert blob[24] != 8
                continue
            assert st == 0, (mode, h, w, level, st)
            assert np.array_equal(got, ref), (mode, h, w, level)
            n_ok += 1
    assert n_ok >= 15


def test_png_every_filter_type(host):
    """Hand-built PNGs with each scanline filter (0..4) forced for every row, 3 and 4 bytes per pixel."""
    import struct

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))

    def paeth(a, b, c):
        p = a + b - c
        pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
        return a if pa <= pb and pa <= pc else (b if pb <= pc else c)

    rng = np.random.default_rng(3)
    for ch, ctype in ((3, 2), (4, 6), (1, 0), (2, 4)):
        img = smooth_image(rng, 37, 53, ch)
        for ft in range(5):
            raw = bytearray()
            prev = np.zeros(53 * ch, np.int32)
            for y in range(37):
                cur = img[y].reshape(-1).astype(np.int32)
                line = np.zeros_like(cur)
                for x in range(cur.size):
                    a = cur[x - ch] if x >= ch else 0
                    b = prev[x]
                    c = prev[x - ch] if x >= ch else 0
                    pred = (0, a, b, (a + b) >> 1, paeth(a, b, c))[ft]
                    line[x] = (cur[x] - pred) & 255
                raw += bytes([ft]) + line.astype(np.uint8).tobytes()
                prev = cur
            ihdr = struct.pack(">IIBBBBB", 53, 37, 8, ctype, 0, 0, 0)
            z = zlib.compress(bytes(raw), 6)
            # IDAT split into three chunks, one of them tiny: the gather path
            blob = (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", ihdr) + chunk(b"IDAT", z[:5]) + chunk(b"IDAT", z[5:len(z) // 2]) +
                    chunk(b"IDAT", z[len(z) // 2:]) + chunk(b"IEND", b""))
            ref = np.asarray(Image.open(io.BytesIO(blob)).convert("RGB"))
            st, got = _decode(host, blob)
            assert st == 0 and np.array_equal(got, ref), (ch, ft, st)


def test_png_transparency_chunks_are_dropped_like_convert_rgb(host):
    rng = np.random.default_rng(11)
    a = smooth_image(rng, 40, 60, 3)
    p = Image.fromarray(a).convert("P", palette=Image.ADAPTIVE, colors=256)
    blob = png_bytes(p, transparency=3)
    assert blob[24] == 8
    st, got = _decode(host, blob)
    assert st == 0 and np.array_equal(got, np.asarray(Image.open(io.BytesIO(blob)).convert("RGB")))
    blob = png_bytes(Image.fromarray(a), transparency=(1, 2, 3))
    st, got = _decode(host, blob)
    assert st == 0 and np.array_equal(got, np.asarray(Image.open(io.BytesIO(blob)).convert("RGB")))


def test_png_detects_damage(host):
    rng = np.random.default_rng(5)
    blob = bytearray(png_bytes(smooth_image(rng, 64, 64, 3)))
    ref = np.asarray(Image.open(io.BytesIO(bytes(blob))).convert("RGB"))
    assert _decode(host, bytes(blob[:50]))[0] == -1  # truncated container
    assert _decode(host, b"\xff\xd8\xff\xe0" + bytes(100))[0] == -1  # a JPEG
    idat = bytes(blob).find(b"IDAT") + 4
    hits = 0
    for k in range(100):
        bad = bytearray(blob)
        bad[idat + 2 + int(rng.integers(0, len(blob) - idat - 30))] ^= 1 << int(rng.integers(0, 8))
        st, got = _decode(host, bytes(bad))
        hits += int(st != 0)
        assert st != 0 or np.array_equal(got, ref)  # the Adler-32 check catches silent corruption
    assert hits >= 95


def test_c_abi_probe_reports_what_the_device_handles():
    """tise_png_probe is host-only: callable without a GPU."""
    from tise_toolbox_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(2)
    info = _lib.PngInfo()
    blob = png_bytes(smooth_image(rng, 300, 200, 3))
    assert lib.tise_png_probe(blob, len(blob), C.byref(info)) == 0
    assert (info.width, info.height, info.bit_depth, info.color_type, info.channels, info.supported) == (200, 300, 8, 2, 3, 1)
    assert info.n_idat >= 1 and info.idat_bytes > 0
    assert lib.tise_png_decode_workspace_bytes(C.byref(info), 1) >= info.idat_bytes + 300 * 601
    a16 = (rng.integers(0, 65536, (20, 20))).astype(np.uint16)
    blob = png_bytes(Image.fromarray(a16))
    assert lib.tise_png_probe(blob, len(blob), C.byref(info)) == 0 and info.supported == 0 and info.bit_depth == 16
    jpg = io.BytesIO()
    Image.fromarray(smooth_image(rng, 32, 32, 3)).save(jpg, "JPEG")
    assert lib.tise_png_probe(jpg.getvalue(), len(jpg.getvalue()), C.byref(info)) < 0 and info.supported == 0
    # a scanline that does not fit the unfilter kernel's shared memory (two rows) is reported at probe time, so that one
    # such file in a folder takes the host decoder instead of failing the whole tise_png_decode call
    for width, want in ((38000, 1), (40000, 0)):
        blob = png_bytes(np.zeros((1, width, 3), dtype=np.uint8))
        assert lib.tise_png_probe(blob, len(blob), C.byref(info)) == 0 and info.width == width
        assert info.supported == want, (width, info.supported)


def _filter_rows(rows, bpp, filters):
    """PNG scanline filtering (spec 9.2) of raw row bytes with the given filter type per row."""
    def paeth(a, b, c):
        p = a + b - c
        pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
        return a if pa <= pb and pa <= pc else (b if pb <= pc else c)

    out = bytearray()
    prev = np.zeros(len(rows[0]), np.int32)
    for y, r in enumerate(rows):
        cur = np.frombuffer(r, np.uint8).astype(np.int32)
        ft = filters[y % len(filters)]
        line = np.zeros_like(cur)
        for x in range(cur.size):
            a = cur[x - bpp] if x >= bpp else 0
            b = prev[x]
            c = prev[x - bpp] if x >= bpp else 0
            line[x] = (cur[x] - (0, a, b, (a + b) >> 1, paeth(a, b, c))[ft]) & 255
        out += bytes([ft]) + line.astype(np.uint8).tobytes()
        prev = cur
    return bytes(out)


def build_png(width, height, depth, ctype, rows, bpp, filters=(0, 1, 2, 3, 4), plte=None):
    import struct

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))

    blob = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", width, height, depth, ctype, 0, 0, 0))
    if plte is not None:
        blob += chunk(b"PLTE", plte)
    return blob + chunk(b"IDAT", zlib.compress(_filter_rows(rows, bpp, filters), 6)) + chunk(b"IEND", b"")


def exotic_pngs(rng):
    """16-bit RGB / RGBA / grey+alpha and 1 / 2 / 4-bit grey and palette files, every filter type, odd widths."""
    out = []
    h = 9
    for ctype, samples in ((2, 3), (6, 4), (4, 2)):
        w = 13
        a = rng.integers(0, 65536, (h, w, samples)).astype(">u2")
        out.append(("16-bit ctype %d" % ctype, build_png(w, h, 16, ctype, [a[y].tobytes() for y in range(h)], samples * 2)))
    for depth in (1, 2, 4):
        for w in (1, 7, 8, 37):
            vals = rng.integers(0, 1 << depth, (h, w))
            rows = []
            for y in range(h):
                bits = "".join(format(int(v), "0%db" % depth) for v in vals[y])
                bits += "0" * ((-len(bits)) % 8)
                rows.append(bytes(int(bits[i:i + 8], 2) for i in range(0, len(bits), 8)))
            out.append(("%d-bit grey w%d" % (depth, w), build_png(w, h, depth, 0, rows, 1)))
            pal = rng.integers(0, 256, ((1 << depth) - (1 if depth > 1 else 0)) * 3, dtype=np.uint8).tobytes()  # short palette
            out.append(("%d-bit palette w%d" % (depth, w), build_png(w, h, depth, 3, rows, 1, plte=pal)))
    out.append(("PIL mode 1", png_bytes(Image.fromarray((rng.integers(0, 2, (21, 45)) * 255).astype(np.uint8)).convert("1"))))
    p = Image.fromarray(smooth_image(rng, 33, 50, 3)).convert("P", palette=Image.ADAPTIVE, colors=12)
    out.append(("PIL P 4-bit", png_bytes(p, bits=4)))
    return out


def test_png_16bit_and_subbyte_formats(host):
    rng = np.random.default_rng(21)
    for name, blob in exotic_pngs(rng):
        try:
            ref = np.asarray(Image.open(io.BytesIO(blob)).convert("RGB"))
        except IndexError:  # Pillow refuses a palette index beyond PLTE; the device decoder writes black
            continue
        st, got = _decode(host, blob)
        assert st == 0, (name, st)
        assert np.array_equal(got, ref), name
    g16 = rng.integers(0, 65536, (4, 4)).astype(">u2")
    blob = build_png(4, 4, 16, 0, [g16[y].tobytes() for y in range(4)], 2)
    assert _decode(host, blob)[0] == -2  # 16-bit grey ("I;16" saturates in convert("RGB")): host decode, reported as such


def build_adam7_png(a, depth, ctype, row_bytes, bpp, plte=None):
    """Interlaced (Adam7) PNG of array a: the seven sub-images are filtered and concatenated as the spec prescribes."""
    import struct

    h, w = a.shape[:2]
    xst, yst = [0, 4, 0, 2, 0, 1, 0], [0, 0, 4, 0, 2, 0, 1]
    xsp, ysp = [8, 8, 4, 4, 2, 2, 1], [8, 8, 8, 4, 4, 2, 2]
    raw = b""
    for p in range(7):
        sub = a[yst[p]::ysp[p], xst[p]::xsp[p]]
        if sub.shape[0] and sub.shape[1]:
            raw += _filter_rows([row_bytes(sub[y]) for y in range(sub.shape[0])], bpp, (0, 1, 2, 3, 4))

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))

    blob = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 1))
    if plte is not None:
        blob += chunk(b"PLTE", plte)
    return blob + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b"")


def interlaced_pngs(rng, sizes=((1, 1), (2, 2), (3, 5), (8, 8), (9, 17), (33, 70), (64, 48), (5, 1), (1, 9))):
    def pack4(r):
        bits = "".join(format(int(v), "04b") for v in r)
        bits += "0" * ((-len(bits)) % 8)
        return bytes(int(bits[i:i + 8], 2) for i in range(0, len(bits), 8))

    out = []
    for (h, w) in sizes:
        a = smooth_image(rng, h, w, 4)
        out.append(("adam7 rgb8 %dx%d" % (h, w), build_adam7_png(a[..., :3], 8, 2, lambda r: r.tobytes(), 3)))
        out.append(("adam7 rgba8 %dx%d" % (h, w), build_adam7_png(a, 8, 6, lambda r: r.tobytes(), 4)))
        out.append(("adam7 grey8 %dx%d" % (h, w), build_adam7_png(a[..., 0], 8, 0, lambda r: r.tobytes(), 1)))
        out.append(("adam7 rgb16 %dx%d" % (h, w), build_adam7_png(a[..., :3].astype(">u2") * 257, 16, 2, lambda r: r.tobytes(), 6)))
        out.append(("adam7 grey4 %dx%d" % (h, w), build_adam7_png(a[..., 0] >> 4, 4, 0, pack4, 1)))
        pal = rng.integers(0, 256, 48, dtype=np.uint8).tobytes()
        out.append(("adam7 pal4 %dx%d" % (h, w), build_adam7_png(a[..., 0] >> 4, 4, 3, pack4, 1, pal)))
    return out


def test_png_adam7_interlaced(host):
    rng = np.random.default_rng(8)
    for name, blob in interlaced_pngs(rng):
        st, got = _decode(host, blob)
        assert st == 0, (name, st)
        assert np.array_equal(got, np.asarray(Image.open(io.BytesIO(blob)).convert("RGB"))), name


def test_host_ingest_helpers_match_per_file_calls(tmp_path):
    """tise_file_sizes / tise_read_files / tise_probe_many (one call per decode group, host threads inside the library)
    against os.path.getsize, plain reads and the per-file probes; a missing file is reported, not skipped."""
    from tise_toolbox_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(7)
    paths, blobs = [], []
    for i in range(37):
        kind = i % 3
        if kind == 0:
            blob = png_bytes(smooth_image(rng, 20 + i, 30 + 2 * i, 3))
        elif kind == 1:
            b = io.BytesIO()
            Image.fromarray(smooth_image(rng, 24 + i, 40, 3)).save(b, "JPEG", quality=85)
            blob = b.getvalue()
        else:
            blob = bytes(rng.integers(0, 256, 100 + i, dtype=np.uint8))  # neither container
        p = tmp_path / ("f%02d.bin" % i)
        p.write_bytes(blob)
        paths.append(str(p))
        blobs.append(blob)
    n = len(paths)
    arr = (C.c_char_p * n)(*[os.fsencode(p) for p in paths])
    sizes = np.zeros(n, dtype=np.int64)
    assert lib.tise_file_sizes(arr, n, sizes.ctypes.data, 4) == 0
    assert sizes.tolist() == [len(b) for b in blobs]
    offs = np.zeros(n, dtype=np.int64)
    offs[1:] = np.cumsum((sizes + 15) // 16 * 16)[:-1]
    buf = np.zeros(int(offs[-1] + sizes[-1]) + 16, dtype=np.uint8)
    failed = C.c_int(-1)
    assert lib.tise_read_files(arr, offs.ctypes.data, sizes.ctypes.data, n, buf.ctypes.data, 4, C.byref(failed)) == 0
    for i in range(n):
        assert bytes(buf[offs[i]:offs[i] + sizes[i]]) == blobs[i]
    pinfos, jinfos = (_lib.PngInfo * n)(), (_lib.JpegInfo * n)()
    kind = np.full(n, -1, dtype=np.int8)
    hw = np.zeros((n, 2), dtype=np.int32)
    assert lib.tise_probe_many(buf.ctypes.data, offs.ctypes.data, sizes.ctypes.data, n, 1, pinfos, jinfos, kind.ctypes.data,
                               hw.ctypes.data, 3) == 0
    for i in range(n):
        one_p, one_j = _lib.PngInfo(), _lib.JpegInfo()
        if lib.tise_png_probe(blobs[i], len(blobs[i]), C.byref(one_p)) == 0 and one_p.supported:
            assert kind[i] == 1 and tuple(hw[i]) == (one_p.height, one_p.width) and pinfos[i].idat_bytes == one_p.idat_bytes
        elif lib.tise_jpeg_probe(blobs[i], len(blobs[i]), C.byref(one_j)) == 0 and one_j.supported:
            assert kind[i] == 2 and tuple(hw[i]) == (one_j.height, one_j.width)
        else:
            assert kind[i] == 0
    assert set(kind.tolist()) == {0, 1, 2}
    assert lib.tise_probe_many(buf.ctypes.data, offs.ctypes.data, sizes.ctypes.data, n, 0, pinfos, jinfos, kind.ctypes.data,
                               hw.ctypes.data, 1) == 0 and 2 not in kind.tolist()  # JPEG decode switched off
    # a missing file
    arr2 = (C.c_char_p * 2)(os.fsencode(paths[0]), os.fsencode(str(tmp_path / "nope.png")))
    s2 = np.zeros(2, dtype=np.int64)
    assert lib.tise_file_sizes(arr2, 2, s2.ctypes.data, 2) != 0 and s2[1] == -1 and s2[0] == len(blobs[0])
    s2[1] = 10
    assert lib.tise_read_files(arr2, np.array([0, 4096], dtype=np.int64).ctypes.data, s2.ctypes.data, 2,
                               np.zeros(8192, dtype=np.uint8).ctypes.data, 2, C.byref(failed)) != 0 and failed.value == 1
