"""CPU checks of the JPEG ingest: tise_toolbox_b200/csrc/jpeg_core.h -- the source of the device
decoder -- compiled for the host (tests/jpeg_host.cpp) and compared bit for bit with Pillow, i.e. with libjpeg-turbo's
default decoder (islow IDCT, fancy upsampling), which is what the reference's Image.open(f).convert("RGB") runs
(img_data.py:21).  The GPU build of the same source is checked in test_gpu_png.py.  Nothing in the package uses this
host build."""
import ctypes as C
import io
import os
import subprocess

import numpy as np
import pytest
from PIL import Image

from test_png_host import smooth_image

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def host():
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "jpeg_host.so")
    src = os.path.join(HERE, "jpeg_host.cpp")
    core = os.path.join(HERE, "..", "tise_toolbox_b200", "csrc", "jpeg_core.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(core)):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", src, "-o", so])
    lib = C.CDLL(so)
    lib.host_jpeg_decode.argtypes = [C.c_char_p, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_int32)]
    return lib


def _decode(lib, blob, cap=1 << 24):
    hw = (C.c_int32 * 2)()
    out = np.zeros(cap, np.uint8)
    st = lib.host_jpeg_decode(blob, len(blob), out.ctypes.data, out.size, hw)
    return st, (out[:hw[0] * hw[1] * 3].reshape(hw[0], hw[1], 3) if st == 0 else None)


def jpeg_bytes(a, **kw):
    b = io.BytesIO()
    Image.fromarray(a).save(b, "JPEG", **kw)
    return b.getvalue()


def _pil(blob):
    return np.asarray(Image.open(io.BytesIO(blob)).convert("RGB"))


JPEG_SIZES = [(256, 256), (64, 48), (17, 33), (1, 1), (2, 3), (5, 4), (500, 375), (33, 700), (8, 8), (16, 16), (15, 17)]


def jpeg_cases(rng, sizes=JPEG_SIZES):
    """(name, blob): chroma 4:4:4 / 4:2:2 / 4:2:0 and grey, several qualities, optimised tables, restart intervals."""
    out = []
    for (h, w) in sizes:
        a = smooth_image(rng, h, w, 3)
        for sub in (0, 1, 2):
            for q in (30, 75, 95, 100):
                out.append(("rgb %dx%d sub%d q%d" % (h, w, sub, q), jpeg_bytes(a, quality=q, subsampling=sub)))
        out.append(("grey %dx%d" % (h, w), jpeg_bytes(a[..., 0], quality=85)))
        out.append(("optimised %dx%d" % (h, w), jpeg_bytes(a, quality=80, optimize=True)))
        out.append(("restart rows %dx%d" % (h, w), jpeg_bytes(a, quality=80, restart_marker_rows=1)))
        out.append(("restart blocks %dx%d" % (h, w), jpeg_bytes(a, quality=90, subsampling=0, restart_marker_blocks=3)))
        for sub in (0, 1, 2):  # progressive: spectral selection + successive approximation, ~10 scans, tables per scan
            out.append(("progressive %dx%d sub%d" % (h, w, sub), jpeg_bytes(a, quality=[40, 85, 100][sub], subsampling=sub, progressive=True)))
        out.append(("progressive grey %dx%d" % (h, w), jpeg_bytes(a[..., 1], quality=70, progressive=True)))
        out.append(("progressive restart %dx%d" % (h, w), jpeg_bytes(a, quality=80, progressive=True, restart_marker_rows=1)))
    return out


def test_jpeg_decode_matches_pillow_bit_for_bit(host):
    rng = np.random.default_rng(0)
    n = 0
    for name, blob in jpeg_cases(rng):
        st, got = _decode(host, blob)
        assert st == 0, (name, st)
        assert np.array_equal(got, _pil(blob)), name
        n += 1
    assert n > 200


def test_jpeg_extreme_content(host):
    """Saturated / high-contrast content drives the range-limit table and the colour clamps."""
    rng = np.random.default_rng(1)
    for k in range(12):
        a = (rng.integers(0, 2, (40, 56, 3)) * 255).astype(np.uint8) if k % 2 else rng.integers(0, 256, (40, 56, 3), dtype=np.uint8)
        for sub in (0, 2):
            blob = jpeg_bytes(a, quality=[10, 50, 100][k % 3], subsampling=sub)
            st, got = _decode(host, blob)
            assert st == 0 and np.array_equal(got, _pil(blob)), (k, sub)


def test_jpeg_variants_left_to_the_host(host):
    rng = np.random.default_rng(2)
    a = smooth_image(rng, 64, 64, 3)
    b = io.BytesIO()
    Image.fromarray(np.dstack([a, a[..., :1]]), "CMYK").save(b, "JPEG")
    assert _decode(host, b.getvalue())[0] == -2
    assert _decode(host, b"\x89PNG\r\n\x1a\n" + bytes(64))[0] == -1
    assert _decode(host, jpeg_bytes(a)[:200])[0] in (-1, 24, 21, 22)  # truncated: rejected, never a crash


def test_jpeg_damage_never_crashes(host):
    rng = np.random.default_rng(3)
    blob = bytearray(jpeg_bytes(smooth_image(rng, 64, 64, 3), quality=90))
    sos = bytes(blob).find(b"\xff\xda")
    for k in range(200):
        bad = bytearray(blob)
        bad[sos + 14 + int(rng.integers(0, len(blob) - sos - 20))] ^= 1 << int(rng.integers(0, 8))
        st, got = _decode(host, bytes(bad))
        assert st >= 0 and (st != 0 or got.shape == (64, 64, 3))


def test_c_abi_jpeg_probe():
    from tise_toolbox_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(4)
    a = smooth_image(rng, 300, 200, 3)
    info = _lib.JpegInfo()
    blob = jpeg_bytes(a, quality=90)  # Pillow default: 4:2:0
    assert lib.tise_jpeg_probe(blob, len(blob), C.byref(info)) == 0
    assert (info.width, info.height, info.components, info.chroma_mode, info.supported) == (200, 300, 3, 3, 1)
    assert info.blocks == 13 * 19 * 4 + 2 * 13 * 19 and info.scan_bytes > 0
    assert lib.tise_jpeg_decode_workspace_bytes(C.byref(info), 1) >= info.blocks * 192
    blob = jpeg_bytes(a, progressive=True)
    assert lib.tise_jpeg_probe(blob, len(blob), C.byref(info)) == 0 and info.supported == 1 and info.progressive == 1 and info.scans > 3
    cmyk = io.BytesIO()
    Image.fromarray(np.dstack([a, a[..., :1]]), "CMYK").save(cmyk, "JPEG")
    assert lib.tise_jpeg_probe(cmyk.getvalue(), len(cmyk.getvalue()), C.byref(info)) == 0 and info.supported == 0
    assert lib.tise_jpeg_probe(b"not a jpeg at all", 17, C.byref(info)) < 0


def test_decoders_survive_fuzzing_under_sanitizers():
    """tools/fuzz_decoders.py: 1200 damaged PNG / JPEG files through the host builds under ASan + UBSan."""
    import shutil
    import sys

    if not shutil.which("g++"):
        pytest.skip("g++ not found")
    asan = subprocess.run(["g++", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan):
        pytest.skip("libasan not available")
    r = subprocess.run([sys.executable, os.path.join(HERE, "..", "tools", "fuzz_decoders.py"), "10"], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0 and "no sanitizer reports" in r.stdout, (r.stdout[-500:], r.stderr[-2000:])
