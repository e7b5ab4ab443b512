"""-m gpu: crop hand-off.  tise_crop_resize_u8 must reproduce, byte for byte, what the
reference does in two steps through the file system: ``img.crop(box)`` saved as PNG by
object_fidelity/crop_object.py:38-46, re-opened and ``transforms.Resize((299, 299))``-d by the O-FID / O-IS
datasets (fid_score.py:210, object_centric_inception_score.py:89)."""
import io

import numpy as np
import pytest
import torch
from PIL import Image

from tise_toolbox_b200 import _lib, synth

from test_png_host import png_bytes

pytestmark = pytest.mark.gpu


def _reference(img, box):
    crop = Image.fromarray(img).crop(tuple(box))  # PIL rounds the box with int(round(v))
    b = io.BytesIO()
    crop.save(b, "PNG")  # the PNG round trip of crop_object.py is lossless; kept to follow the reference literally
    return np.asarray(Image.open(io.BytesIO(b.getvalue())).convert("RGB").resize((299, 299), Image.BILINEAR))


def _boxes(rng, n, h, w):
    out = []
    for _ in range(n):
        x0, y0 = rng.uniform(0, w - 2), rng.uniform(0, h - 2)
        x1, y1 = rng.uniform(x0 + 1.0, w), rng.uniform(y0 + 1.0, h)
        out.append((x0, y0, x1, y1))
    out[0] = (0.0, 0.0, float(w), float(h))  # the whole image
    out[1] = (w - 1.0, h - 1.0, float(w), float(h))  # a single pixel
    out[2] = (0.5, 1.5, 2.5, 3.5)  # halves: Python rounds to even -> (0, 2, 2, 4)
    return np.array(out, dtype=np.float32)


def test_crop_resize_uniform_sources(engine):
    rng = np.random.default_rng(0)
    imgs = synth.images(1, 0, 0, 6)
    boxes = _boxes(rng, 40, 256, 256)
    idx = rng.integers(0, 6, 40)
    got = engine.resize_crops(imgs.cuda(), idx, boxes).cpu().numpy()
    for i in range(40):
        ref = _reference(imgs[idx[i]].numpy(), boxes[i])
        assert np.array_equal(got[i], ref), (i, boxes[i])


def test_crop_resize_from_decoded_pack(engine):
    """decode (PNG on the device) -> crop + resize, no host round trip: big sources exercise the down-scaling path."""
    rng = np.random.default_rng(1)
    srcs = [synth.images(5, 1, i, 1, h, w)[0].numpy() for i, (h, w) in enumerate([(480, 640), (333, 500), (64, 48), (900, 700)])]
    packed = engine.decode_images([png_bytes(a) for a in srcs])
    idx, boxes = [], []
    for k, a in enumerate(srcs):
        b = _boxes(rng, 8, a.shape[0], a.shape[1])
        boxes.append(b)
        idx += [k] * 8
    boxes = np.concatenate(boxes)
    got = engine.resize_crops(packed, idx, boxes).cpu().numpy()
    for i in range(len(idx)):
        assert np.array_equal(got[i], _reference(srcs[idx[i]], boxes[i])), (i, boxes[i])
    feats = engine.features_from_resized(torch.from_numpy(got).cuda())
    assert tuple(feats.shape) == (32, 2048) and bool(torch.isfinite(feats).all())


def test_crop_box_errors(engine):
    imgs = synth.images(1, 0, 0, 1).cuda()
    for bad in [(0, 0, 257, 10), (-1, 0, 10, 10), (5, 5, 5, 9), (10, 20, 5, 30)]:
        with pytest.raises(_lib.TiseError):
            engine.resize_crops(imgs, [0], [bad])
    assert tuple(engine.resize_crops(imgs, [], np.zeros((0, 4))).shape) == (0, 299, 299, 3)
