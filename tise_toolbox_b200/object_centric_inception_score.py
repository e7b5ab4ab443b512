"""Drop-in for object_fidelity/O-IS/object_centric_inception_score.py on B200.

``inception_score(imgs, cuda, batch_size, resize, splits)`` and ``IgnoreLabelDataset`` keep the
reference's signatures.  Unlike the reference module this one has no import-time side effects
(the reference parses argv and runs at import, :107-129); run it with ``python -m``.
"""
import argparse
import os

import numpy as np
import torch
import torch.utils.data
from PIL import Image

from . import inception_score as _is
from .engine import Engine, ISPartials
from .img_data import DeviceImageLoader
from .inception import load_state_dict_file

WEIGHTS = "weights/inceptionv3_fine_to_with_80_coco_classes.pth"  # object_centric_inception_score.py:46


def inception_score(imgs, cuda=True, batch_size=32, resize=False, splits=1, state_dict=None, engine=None):
    """object_centric_inception_score.py:17-81.  ``imgs``: dataset of 3x299x299 float images in [-1, 1]
    (the reference's IgnoreLabelDataset) or, for the fused GPU path, a dataset whose items are uint8
    HWC tensors of any size (``IgnoreLabelDataset(path, raw_u8=True)``)."""
    N = len(imgs)
    assert batch_size > 0
    assert N > batch_size
    if engine is None:
        if state_dict is None:
            state_dict = load_state_dict_file(os.environ.get("TISE_OIS_WEIGHTS", WEIGHTS))
        engine = Engine(state_dict, variant="ois", micro_batch=1024)
    engine.set_variant("ois")
    raw = bool(getattr(imgs, "raw_u8", False))
    on_device = raw and hasattr(imgs, "imgspath") and os.environ.get("TISE_GPU_DECODE", "1") != "0"
    if on_device:  # host threads read bytes; PNG decode, resize and normalisation run on the GPU
        files = [os.path.join(imgs.imgspath, name) for name in imgs.namelist]
        dataloader = DeviceImageLoader(files, engine, batch_size, drop_last=False)
    else:
        dataloader = torch.utils.data.DataLoader(imgs, batch_size=batch_size,
                                                 collate_fn=(lambda b: list(b)) if raw else None)
    parts = ISPartials(engine, engine.num_classes, splits)
    start = 0
    for batch in (dataloader.groups() if on_device else dataloader):
        if on_device:  # whole decode groups: the split of an image depends on its global index, not on the batching
            feat = engine.features(batch)
        elif raw:
            feat = engine.features([t.to(engine.device) for t in batch])
        else:
            x = batch.to(engine.device, torch.float32)
            if resize:  # nn.Upsample(size=(299, 299), mode="bilinear"), :49
                x = engine.upsample_float(x, align_corners=False)
            feat = engine.features_from_float(x)
        logits = engine.logits(feat, use_bias=True)
        _is.update_from_logits(engine, parts, logits, "ois", N, start)
        start += int(logits.shape[0])
    mean, std, _ = parts.finalize(nan_on_zero=False)
    return mean, std


class IgnoreLabelDataset(torch.utils.data.Dataset):
    """object_centric_inception_score.py:84-104: flat os.listdir, no extension filter."""

    def __init__(self, imgspath, raw_u8=False):
        from torchvision import transforms

        self.imgspath = imgspath
        self.raw_u8 = raw_u8
        self.transform = transforms.Compose([
            transforms.Resize((299, 299)),
            transforms.ToTensor(),
            transforms.Normalize((0.5, 0.5, 0.5), (0.5, 0.5, 0.5)),
        ])
        self.namelist = os.listdir(self.imgspath)

    def __getitem__(self, index):
        img = Image.open(os.path.join(self.imgspath, self.namelist[index])).convert("RGB")
        if self.raw_u8:
            return torch.from_numpy(np.array(img, dtype=np.uint8))
        return self.transform(img)

    def __len__(self):
        return len(self.namelist)


def parse_args():
    parser = argparse.ArgumentParser()
    parser.add_argument("--image_dir", default="", type=str)
    parser.add_argument("--saved_file", default="", type=str)
    parser.add_argument("--gpu_id", default=0, type=int)
    return parser.parse_args()


if __name__ == "__main__":
    args = parse_args()
    print("Load images from: ", args.image_dir)
    imgs = IgnoreLabelDataset(args.image_dir, raw_u8=True)
    print("Calculating Inception Score...")
    torch.cuda.set_device(args.gpu_id)
    IS_mean, IS_std = inception_score(imgs, cuda=True, batch_size=32, resize=False, splits=10)
    with open(args.saved_file, "w") as f:
        f.write(f"O-IS: {IS_mean} +-  {IS_std}")
    print(f"O-IS: {IS_mean} +- {IS_std}")
