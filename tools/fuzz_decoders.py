"""Fuzzes the HOST builds of the device decoders (csrc/png_core.h, csrc/jpeg_core.h) under AddressSanitizer + UBSan:
bit flips, truncations, random splices and inserted 0xFF runs on valid PNG / JPEG files.  Every out-of-bounds access or
undefined operation the decoder logic could make on damaged input shows up here, without a GPU (on the device the same
accesses would be memory faults inside a kernel).
    python tools/fuzz_decoders.py [trials=60]     (re-executes itself with libasan preloaded)"""
import ctypes as C
import io
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(out_dir):
    libs = {}
    for name in ("png", "jpeg"):
        so = os.path.join(out_dir, "%s_host_asan.so" % name)
        subprocess.check_call(["g++", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-shared",
                               "-fPIC", os.path.join(ROOT, "tests", "%s_host.cpp" % name), "-o", so])
        libs[name] = so
    return libs


def run(trials, libs):
    import numpy as np
    from PIL import Image

    png, jpg = C.CDLL(libs["png"]), C.CDLL(libs["jpeg"])
    for f in (png.host_png_decode, jpg.host_jpeg_decode):
        f.argtypes = [C.c_char_p, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_int32)]
    rng = np.random.default_rng(0)
    out = np.zeros(1 << 22, np.uint8)
    hw = (C.c_int32 * 2)()
    n = 0
    for trial in range(trials):
        h, w = int(rng.integers(1, 90)), int(rng.integers(1, 90))
        a = np.stack([np.asarray(Image.fromarray(rng.integers(0, 256, (h // 8 + 2, w // 8 + 2)).astype(np.uint8))
                                 .resize((w, h), Image.BICUBIC)) for _ in range(3)], -1).astype(int) + rng.integers(-8, 9, (h, w, 3))
        a = np.clip(a, 0, 255).astype(np.uint8)
        b = io.BytesIO()
        Image.fromarray(a).save(b, "PNG")
        pb = b.getvalue()
        b = io.BytesIO()
        kw = {"restart_marker_blocks": int(rng.integers(1, 5))} if trial % 3 == 0 else {}
        if trial % 2:
            kw["progressive"] = True  # multi-scan files: spectral selection, successive approximation, EOB runs
        Image.fromarray(a).save(b, "JPEG", quality=int(rng.integers(5, 100)), subsampling=int(rng.integers(0, 3)), **kw)
        jb = b.getvalue()
        for blob, f in ((pb, png.host_png_decode), (jb, jpg.host_jpeg_decode)):
            for k in range(60):
                bad = bytearray(blob)
                if k % 4 == 0:
                    for _ in range(int(rng.integers(1, 4))):
                        bad[int(rng.integers(0, len(bad)))] ^= 1 << int(rng.integers(0, 8))
                elif k % 4 == 1:
                    bad = bad[:int(rng.integers(0, len(bad)))]
                elif k % 4 == 2:
                    i, m = int(rng.integers(0, len(bad))), int(rng.integers(1, 9))
                    bad[i:i + m] = bytes(rng.integers(0, 256, m, dtype=np.uint8))
                else:
                    i = int(rng.integers(0, len(bad)))
                    bad[i:i] = bytes([0xFF] * int(rng.integers(1, 4)))
                bad = bytes(bad)
                f(bad, len(bad), out.ctypes.data, out.size, hw)
                n += 1
    print("fuzz cases %d: no sanitizer reports" % n)


if __name__ == "__main__":
    trials = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 60
    if os.environ.get("TISE_FUZZ_CHILD"):
        run(trials, {"png": os.environ["TISE_FUZZ_PNG"], "jpeg": os.environ["TISE_FUZZ_JPEG"]})
    else:
        asan = subprocess.check_output(["g++", "-print-file-name=libasan.so"], text=True).strip()
        if not os.path.isabs(asan):
            sys.exit("libasan.so not found next to g++")
        with tempfile.TemporaryDirectory() as d:
            libs = build(d)
            env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0", TISE_FUZZ_CHILD="1", TISE_FUZZ_PNG=libs["png"],
                       TISE_FUZZ_JPEG=libs["jpeg"])
            sys.exit(subprocess.call([sys.executable, os.path.abspath(__file__), str(trials)], env=env))
