"""Drop-in for the two TensorFlow IS* scripts on B200:
  image_realism/IS/coco/inception_score_star_coco.py  (variant "coco")
  image_realism/IS/bird/inception_score_star_bird.py  (variant "bird")

The reference builds TF-1.x graphs (a downloaded frozen Inception for coco, a slim Inception-v3 fine-tuned
on CUB for bird).  Both are the Inception-v3 topology; here the trunk is the same libtise_b200 network as
for FID, with the caller's state_dict (1008-way fc without bias for coco, 51-way fc with bias for bird).
``preprocess``, ``load_data``, ``get_inception_score`` and ``inception_score`` keep the reference's names
and argument meaning; images are file paths (coco) or HWC uint8 arrays (bird), exactly as there.
"""
import math
import os
import sys
from argparse import ArgumentDefaultsHelpFormatter, ArgumentParser

import numpy as np
import torch
from PIL import Image

from . import inception_score as _is
from .engine import Engine, ISPartials
from .inception import load_state_dict_file


def imread(path):
    """scipy.misc.imread(path): the file's own mode, as a uint8 array."""
    return np.array(Image.open(path))


def to_rgb_u8(img):
    """inception_score_star_coco.py:116-117 / _bird.py:65-66: a 2-D (grey) image becomes HxWx3 through
    ``np.resize`` (which tiles the flattened data -- a quirk kept on purpose); RGBA loses its alpha."""
    img = np.asarray(img)
    if len(img.shape) == 2:
        img = np.resize(img, (img.shape[0], img.shape[1], 3))
    if img.shape[2] > 3:
        img = img[:, :, :3]
    return np.ascontiguousarray(img, dtype=np.uint8)


def preprocess(img, variant="coco"):
    """CPU statement of the reference's preprocess (coco :115-121, bird :64-71) for callers that want the
    array; the scoring path below does the same resize + scaling on the GPU."""
    img = to_rgb_u8(img)
    img = np.asarray(Image.fromarray(img).resize((299, 299), Image.BILINEAR)).astype(np.float32)
    if variant == "bird":
        img = img / 127.5 - 1.0
    return np.expand_dims(img, 0)


def load_data(fullpath, variant="coco"):
    """coco :124-135 returns file names; bird :114-125 returns the decoded arrays."""
    print("[Data] Read data from " + fullpath)
    images = []
    for path, subdirs, files in os.walk(fullpath):
        for name in files:
            if name.rfind("jpg") != -1 or name.rfind("png") != -1:
                filename = os.path.join(path, name)
                if os.path.isfile(filename):
                    images.append(imread(filename) if variant == "bird" else filename)
    return images


def get_inception_score(images, splits=10, variant="coco", engine=None, state_dict=None, batch_size=64,
                        shuffle=None, rng=None):
    """coco :32-60 / bird :74-111.  ``images``: list of paths or of HWC uint8 arrays.
    bird keeps floor(N / batch_size) full batches of a shuffled order (``shuffle`` defaults to True for
    bird as in the reference, pass False or an ``rng`` for reproducibility); coco scores every image."""
    if engine is None:
        if state_dict is None:
            state_dict = load_state_dict_file(os.environ["TISE_IS_WEIGHTS"])
        engine = Engine(state_dict, variant=variant)
    engine.set_variant(variant)
    _, off, use_bias, nan_on_zero = _is.VARIANTS[variant]
    n = len(images)
    order = np.arange(n)
    if variant == "bird":
        assert type(images[0]) == np.ndarray and len(images[0].shape) == 3
        assert np.max(images[0]) > 10 and np.min(images[0]) >= 0.0
        if shuffle is None or shuffle:
            (rng or np.random).shuffle(order)
        n = _is.kept_images(n, "bird", batch_size)
        order = order[:n]
    parts = ISPartials(engine, engine.num_classes - off, splits)
    if n and all(isinstance(images[i], str) for i in order) and os.environ.get("TISE_GPU_DECODE", "1") != "0":
        # file paths (coco): bytes go file -> pinned memory -> HBM; RGB / RGBA PNGs and YCbCr JPEGs are decoded on the device,
        # 4096 files per call (the split of an image depends on its global index only, so the batching is free).
        # scipy.misc.imread keeps the file's own mode, and the scripts' grey -> RGB step is np.resize (a tiling quirk,
        # :116-117), so every other kind of file takes exactly that host path.
        import io
        from concurrent.futures import ThreadPoolExecutor

        def host_decode(blob):
            return to_rgb_u8(np.array(Image.open(io.BytesIO(blob))))

        group = max(1, 4096 // batch_size) * batch_size
        with ThreadPoolExecutor(max_workers=8) as readers:
            for start in range(0, n, group):
                paths = [images[i] for i in order[start:start + group]]
                staged = engine.stage_files(paths, readers)
                for k in range(staged["n"]):  # only files whose own mode is RGB / RGBA may skip the scripts' host path
                    if staged["kind"][k] == 1 and staged["infos"][k].color_type not in (2, 6):
                        staged["supported"][k] = False
                    if staged["kind"][k] == 2 and staged["jinfos"][k].components != 3:
                        staged["supported"][k] = False
                pack = engine.decode_staged(staged, host_decoder=host_decode)
                logits = engine.logits(engine.features(pack), use_bias=use_bias)
                _is.update_from_logits(engine, parts, logits, variant, n, start)
        mean, std, _ = parts.finalize(nan_on_zero=nan_on_zero)
        return mean, std
    for start in range(0, n, batch_size):
        idx = order[start:start + batch_size]
        batch = [torch.from_numpy(to_rgb_u8(imread(images[i]) if isinstance(images[i], str) else images[i]))
                 .to(engine.device) for i in idx]
        feat = engine.features(batch)
        logits = engine.logits(feat, use_bias=use_bias)
        _is.update_from_logits(engine, parts, logits, variant, n, start)
    mean, std, _ = parts.finalize(nan_on_zero=nan_on_zero)
    return mean, std


def inception_score(path, variant="coco", **kw):
    images = load_data(path, variant)
    return get_inception_score(images, variant=variant, **kw)


if __name__ == "__main__":
    parser = ArgumentParser(formatter_class=ArgumentDefaultsHelpFormatter)
    parser.add_argument("--image_folder", type=str, default="")
    parser.add_argument("--saved_file", type=str, default="")
    parser.add_argument("--gpu", type=int, default=0)
    parser.add_argument("--variant", type=str, default="coco", choices=["coco", "bird"])
    parser.add_argument("--weights", type=str, default=None)
    parser.add_argument("--splits", type=int, default=10)
    parser.add_argument("--batch_size", type=int, default=64)
    args = parser.parse_args()
    os.environ["CUDA_VISIBLE_DEVICES"] = str(args.gpu)
    if args.weights:
        os.environ["TISE_IS_WEIGHTS"] = args.weights
    images = load_data(args.image_folder, args.variant)
    print(".......")
    mean, std = get_inception_score(images, splits=args.splits, variant=args.variant, batch_size=args.batch_size)
    if args.saved_file:
        with open(args.saved_file, "w") as f:
            if args.variant == "coco":
                f.write("[Inception Score] mean: {:.5f} std: {:.5f}".format(mean, std))
            else:
                f.write(f"IS = {mean}  +-  {std}")
    print("[Inception Score] mean: {:.2f} std: {:.2f}".format(mean, std))
