"""-m gpu: the reference-facing IS* module (inception_score_star.py <-> image_realism/IS/{coco,bird}/inception_score_star_*.py)
through folders of files, against the oracle's restatement of the scripts' own lines (TensorFlow cannot run here,
SURVEY.md §8c): preprocess coco :115-121 / bird :64-71, heads coco :104-108 / bird :181-194, splits + KL :53-60."""
import os

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import restate
from tise_toolbox_b200 import inception_score_star as iss
from tise_toolbox_b200 import synth
from tise_toolbox_b200.engine import T_BIRD, T_COCO, Engine

pytestmark = pytest.mark.gpu


def _folder(tmp_path, n=36):
    imgs = torch.cat([synth.images(1, 0, 700, n // 2), synth.images(2, 1, 700, n - n // 2)])
    for i, im in enumerate(imgs):
        a = im.numpy()
        name = str(tmp_path / ("im_%03d.png" % i))
        if i == 3:
            Image.fromarray(a[..., 0]).save(name)  # grey: the scripts tile it with np.resize (a quirk, kept)
        elif i == 5:
            Image.fromarray(np.dstack([a, np.full(a.shape[:2], 77, np.uint8)]), "RGBA").save(name)  # alpha is cut off
        elif i == 7:
            Image.fromarray(a).save(str(tmp_path / "im_007.jpg"), quality=95)
        else:
            Image.fromarray(a).save(name)
    return str(tmp_path)


def _oracle_arrays(folder):
    names = iss.load_data(folder, "coco")  # os.walk order, "jpg" / "png" substring filter
    return names, [iss.to_rgb_u8(iss.imread(n)) for n in names]


def test_coco_folder_of_files(tmp_path, monkeypatch):
    folder = _folder(tmp_path)
    names, arrays = _oracle_arrays(folder)
    assert len(names) == 36
    sd = synth.inception_state_dict(4, num_classes=1008)
    trunk = restate.Trunk(sd, num_classes=1008)
    x = restate.preprocess(arrays, "coco")
    sd = synth.calibrate_head(sd, trunk.pool3(x), T_COCO, seed=1008, class_offset=0)
    trunk = restate.Trunk(sd, num_classes=1008)
    ref = restate.isstar_inception_score(restate.coco_preds(trunk.pool3(x), sd["fc.weight"]), 4)
    assert ref[0] > 1.5
    eng = Engine(sd, device="cuda:0", variant="coco", micro_batch=32)
    got = iss.inception_score(folder, variant="coco", engine=eng, splits=4, batch_size=8)
    assert abs(got[0] - ref[0]) <= 1e-3 * ref[0], (got, ref)
    assert abs(got[1] - ref[1]) <= 2e-2 * ref[1] + 1e-4
    monkeypatch.setenv("TISE_GPU_DECODE", "0")  # PIL decode of every file on the host: same images, same score
    host = iss.get_inception_score(names, splits=4, variant="coco", engine=eng, batch_size=8)
    assert abs(host[0] - got[0]) <= 1e-6 * got[0] and abs(host[1] - got[1]) <= 1e-6


def test_bird_arrays_keep_full_batches_only(tmp_path):
    folder = _folder(tmp_path, 21)
    arrays = [a for a in iss.load_data(folder, "bird")]
    arrays = [iss.to_rgb_u8(a) for a in arrays]
    sd = synth.inception_state_dict(5, num_classes=51)
    trunk = restate.Trunk(sd, num_classes=51)
    kept = arrays[:16]  # floor(21 / 8) full batches (bird :83-97), identity order
    x = restate.preprocess(kept, "bird")
    sd = synth.calibrate_head(sd, trunk.pool3(x), T_BIRD, seed=51, class_offset=1)
    trunk = restate.Trunk(sd, num_classes=51)
    ref = restate.isstar_inception_score(restate.bird_preds(trunk.net.fc(trunk.pool3(x))), 2)
    eng = Engine(sd, device="cuda:0", variant="bird", micro_batch=32)
    got = iss.get_inception_score(arrays, splits=2, variant="bird", engine=eng, batch_size=8, shuffle=False)
    assert abs(got[0] - ref[0]) <= 1e-3 * ref[0], (got, ref)
    with pytest.raises(AssertionError):
        iss.get_inception_score([a.astype(np.float32) / 255 for a in arrays], variant="bird", engine=eng, batch_size=8)
