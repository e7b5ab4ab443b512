"""-m gpu: PNG ingest on the device through the C ABI (tise_png_decode), bit-exact against
Pillow's ``Image.open(f).convert("RGB")`` -- the call it replaces (image_realism/FID/img_data.py:21)."""
import io
import os

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import restate
from tise_toolbox_b200 import img_data, synth
from tise_toolbox_b200.engine import ImageDecodeError

from test_png_host import MODES, SIZES, exotic_pngs, interlaced_pngs, png_bytes, smooth_image

pytestmark = pytest.mark.gpu


def _pil(blob):
    return np.asarray(Image.open(io.BytesIO(blob)).convert("RGB"))


def test_decode_mixed_batch_bit_exact(engine):
    rng = np.random.default_rng(0)
    blobs = []
    for mode, c in MODES:
        for (h, w) in SIZES:
            a = smooth_image(rng, h, w, c)
            if mode == "P":
                im = Image.fromarray(a).convert("P", palette=Image.ADAPTIVE, colors=200)
            elif mode == "L":
                im = Image.fromarray(a[..., 0])
            else:
                im = Image.fromarray(a, mode)
            blobs.append(png_bytes(im, compress_level=int(rng.integers(0, 10))))
    jpg = io.BytesIO()
    a = smooth_image(rng, 40, 50, 3)
    Image.fromarray(np.dstack([a, a[..., :1]]), "CMYK").save(jpg, "JPEG")
    blobs.append(jpg.getvalue())  # CMYK JPEG: decoded by PIL like the reference does, uploaded into the same pack
    blobs.append(png_bytes(Image.fromarray(rng.integers(0, 65536, (20, 20)).astype(np.uint16))))  # 16-bit: host
    exotic = [b for _, b in exotic_pngs(rng)]  # 16-bit RGB(A) / GA, 1 / 2 / 4-bit grey and palette, all filter types
    exotic += [b for _, b in interlaced_pngs(rng, sizes=((1, 1), (3, 5), (9, 17), (33, 70)))]  # Adam7
    usable = []
    for b in exotic:
        try:
            _pil(b)
            usable.append(b)
        except IndexError:  # Pillow refuses palette indices beyond PLTE
            pass
    assert all(engine.png_probe(b).supported for b in usable)
    blobs += usable
    packed = engine.decode_images(blobs)
    assert len(packed) == len(blobs)
    for i, b in enumerate(blobs):
        ref = _pil(b)
        got = packed[i].cpu().numpy()
        assert got.shape == ref.shape and np.array_equal(got, ref), i
    with pytest.raises(ValueError):
        engine.decode_images([jpg.getvalue()], host_fallback=False)


def test_decode_synthetic_set_and_resize(engine):
    """cfg1-shaped inputs (256 x 256 RGB) + ragged crops: decode -> resize without a copy in between."""
    imgs = [im.numpy() for im in synth.images(1, 0, 0, 24)]
    sizes = synth.crop_sizes(8, 0, 8)
    imgs += [synth.images(5, 1, i, 1, h, w)[0].numpy() for i, (h, w) in enumerate(sizes)]
    blobs = [png_bytes(a) for a in imgs]
    packed = engine.decode_images(blobs)
    for i, a in enumerate(imgs):
        assert np.array_equal(packed[i].cpu().numpy(), a), i
    out = engine.resize(packed).cpu().numpy()
    ref = np.stack([restate.pillow_resize_bilinear(a) for a in imgs])
    assert np.array_equal(out, ref)


def test_decode_large_batch_many_idat_chunks(engine):
    rng = np.random.default_rng(4)
    a = rng.integers(0, 256, (700, 900, 3), dtype=np.uint8)  # ~1.9 MB incompressible: dozens of IDAT chunks
    blobs = [png_bytes(a), png_bytes(a[:1]), png_bytes(np.zeros((1000, 1000, 3), np.uint8))] + \
            [png_bytes(smooth_image(rng, 64, 64, 3)) for _ in range(300)]
    packed = engine.decode_images(blobs)
    for i in (0, 1, 2, 3, 150, 302):
        assert np.array_equal(packed[i].cpu().numpy(), _pil(blobs[i])), i


def test_corrupt_png_is_reported(engine):
    rng = np.random.default_rng(9)
    good = png_bytes(smooth_image(rng, 64, 64, 3))
    bad = bytearray(good)
    idat = good.find(b"IDAT") + 4
    bad[idat + 40] ^= 0x10
    with pytest.raises(ImageDecodeError):
        engine.decode_images([good, bytes(bad), good])
    assert np.array_equal(engine.decode_images([good])[0].cpu().numpy(), _pil(good))  # the context stays usable
    assert len(engine.decode_images([])) == 0


def test_device_image_loader_order_and_drop_last(tmp_path, engine):
    imgs = synth.images(3, 0, 0, 23)
    for i, im in enumerate(imgs):
        Image.fromarray(im.numpy()).save(str(tmp_path / ("img_%05d.png" % i)))
    names = img_data.Dataset(str(tmp_path)).file_names
    loader = img_data.DeviceImageLoader(names, engine, batch_size=5, drop_last=True)
    assert len(loader) == 4
    seen = 0
    for k, batch in enumerate(loader):
        assert len(batch) == 5
        for j in range(5):
            ref = np.asarray(Image.open(names[k * 5 + j]).convert("RGB"))
            assert np.array_equal(batch[j].cpu().numpy(), ref)
            seen += 1
    assert seen == 20
    assert sum(len(b) for b in img_data.DeviceImageLoader(names, engine, 5, drop_last=False)) == 23


def test_shard_loader_matches_the_png_folder(tmp_path, engine, state_dict):
    """Raw uint8 shard ingest (shards.py): same images, same order, same statistics as the PNG folder path."""
    from tise_toolbox_b200 import fid_score, shards
    from tise_toolbox_b200.inception import InceptionV3

    imgs = [im.numpy() for im in synth.images(3, 0, 40, 21)] + [synth.images(5, 1, 7, 1, 90, 140)[0].numpy()]
    folder = tmp_path / "pngs"
    folder.mkdir()
    for i, a in enumerate(imgs):
        Image.fromarray(a).save(str(folder / ("img_%05d.png" % i)))
    shard = str(tmp_path / "set.shard")
    _, n = shards.pack_folder(str(folder), shard)
    assert n == 22
    names = img_data.Dataset(str(folder)).file_names
    loader = shards.ShardLoader(shard, engine, batch_size=5, drop_last=True, group=10)
    assert len(loader) == 4
    k = 0
    for batch in loader:
        for j in range(len(batch)):
            assert np.array_equal(batch[j].cpu().numpy(), np.asarray(Image.open(names[k]).convert("RGB"))), k
            k += 1
    assert k == 20
    model = InceptionV3([3], state_dict=state_dict, micro_batch=32)
    m1, s1 = fid_score._compute_statistics_of_path(str(folder), model, 5, 2048, "0")
    m2, s2 = fid_score._compute_statistics_of_path(shard, model, 5, 2048, "0")
    assert np.abs(m1 - m2).max() <= 1e-12 * np.abs(m1).max() and np.abs(s1 - s2).max() <= 1e-10 * np.abs(s1).max()


def test_jpeg_decode_bit_exact(engine):
    """Baseline JPEGs on the device (tise_jpeg_decode) == Pillow / libjpeg-turbo, mixed with PNGs in one call."""
    from test_jpeg_host import jpeg_cases

    rng = np.random.default_rng(0)
    cases = jpeg_cases(rng, sizes=[(256, 256), (64, 48), (17, 33), (1, 1), (2, 3), (500, 375), (15, 17)])
    blobs = [b for _, b in cases] + [png_bytes(smooth_image(rng, 33, 44, 3))]
    packed = engine.decode_images(blobs, host_fallback=False)
    for i, b in enumerate(blobs):
        assert np.array_equal(packed[i].cpu().numpy(), _pil(b)), cases[i][0] if i < len(cases) else "png"
    # a damaged entropy-coded segment is reported, not decoded into garbage silently when the stream breaks a code
    good = cases[0][1]
    bad = bytearray(good)
    sos = good.find(b"\xff\xda")
    for k in range(sos + 20, sos + 60):
        bad[k] = 0xFF if k % 2 else 0x00 ^ bad[k]
    try:
        out = engine.decode_images([bytes(bad)], host_fallback=False)
        assert tuple(out[0].shape) == (256, 256, 3)
    except ImageDecodeError:
        pass
    assert np.array_equal(engine.decode_images([good])[0].cpu().numpy(), _pil(good))
