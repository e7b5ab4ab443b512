"""CPU restatement of the reference's image-realism hot path (numpy / torch-CPU / scipy).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Every function cites the reference
file:line (relative to /root/reference) it follows.  The arithmetic of the path lives in
third-party libraries that the reference pins in ``requirements.txt`` (torchvision 0.9.1
``Inception3``, Pillow 8.3.2 ``Image.resize``, numpy ``mean``/``cov``, scipy ``linalg.sqrtm``
and ``stats.entropy``); the same libraries (newer versions, same algorithms) are present in
this image, and this module calls them exactly where the reference does.  The only thing
restated from a published algorithm is Pillow's fixed-point resampler
(``pillow_resize_bilinear``), which tests pin bit-for-bit against ``PIL.Image.resize``.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# A2  Pillow Image.resize(BILINEAR): fid_score.py:210, object_centric_inception_score.py:89,
#     inception_score_star_coco.py:118 (scipy.misc.imresize -> PIL), ..._bird.py:67
#     Algorithm: Pillow src/libImaging/Resample.c (precompute_coeffs, normalize_coeffs_8bpc,
#     ImagingResampleHorizontal_8bpc / Vertical_8bpc), PRECISION_BITS = 32 - 8 - 2.
# --------------------------------------------------------------------------------------
PRECISION_BITS = 32 - 8 - 2


def pillow_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray]:
    """bounds [out,2] (xmin, count) and fixed-point weights [out, ksize] of the triangle filter."""
    scale = float(in_size) / float(out_size)
    filterscale = scale if scale >= 1.0 else 1.0
    support = 1.0 * filterscale  # bilinear filter support is 1.0
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = np.zeros(ksize, dtype=np.float64)
        ww = 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            if a < 0.0:
                a = -a
            v = 1.0 - a if a < 1.0 else 0.0
            w[x] = v
            ww += v
        for x in range(xmax):
            if ww != 0.0:
                w[x] /= ww
        for x in range(ksize):
            if w[x] < 0:
                kk[xx, x] = int(-0.5 + w[x] * (1 << PRECISION_BITS))
            else:
                kk[xx, x] = int(0.5 + w[x] * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _resample_axis0(img: np.ndarray, out_size: int) -> np.ndarray:
    """Resample along axis 0 of a [L, ...] uint8 array with Pillow's 8bpc integer kernel."""
    in_size = img.shape[0]
    bounds, kk = pillow_coeffs(in_size, out_size)
    src = img.astype(np.int64)
    out = np.empty((out_size,) + img.shape[1:], dtype=np.uint8)
    for xx in range(out_size):
        xmin, cnt = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(cnt):
            acc += src[xmin + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def pillow_resize_bilinear(img: np.ndarray, out_h: int = 299, out_w: int = 299) -> np.ndarray:
    """uint8 [H,W,C] -> uint8 [out_h,out_w,C]; horizontal pass first (into uint8), then vertical;
    a pass whose size does not change is skipped, exactly as ImagingResample does."""
    assert img.dtype == np.uint8 and img.ndim == 3
    h, w, _ = img.shape
    tmp = img
    if w != out_w:
        tmp = np.ascontiguousarray(_resample_axis0(np.ascontiguousarray(tmp.transpose(1, 0, 2)), out_w)
                                   .transpose(1, 0, 2))
    if h != out_h:
        tmp = _resample_axis0(tmp, out_h)
    return tmp


# --------------------------------------------------------------------------------------
# A3/A4  u8 -> float input of the network, the four variants of the path
# --------------------------------------------------------------------------------------
INPUT_VARIANTS = ("fid", "ois", "bird", "coco")


def input_lut(variant: str) -> torch.Tensor:
    """float32 [3,256]: value fed to Conv2d_1a for every (channel, uint8 level), computed with the
    very op sequence the reference uses so that the table is bit-exact.

    fid : ToTensor (x/255, fid_score.py:211) then InceptionV3.forward's affine
          (image_realism/FID/inception.py:120-124) -- torchvision transform_input on [0,1] data.
    ois : ToTensor then Normalize(.5,.5) (object_centric_inception_score.py:90-91),
          transform_input=False (:42).
    bird: img.astype(float32) / 127.5 - 1.0 (inception_score_star_bird.py:68-70).
    coco: float32 0..255 into the 2015 frozen graph, whose first ops are Sub(128) and
          Mul(1/128) (inception_score_star_coco.py:119; graph 'Sub', 'Mul' nodes).
    """
    lv = torch.arange(256, dtype=torch.float32)
    if variant == "fid":
        x = (lv / 255).view(1, 256).repeat(3, 1)
        x[0] = x[0] * (0.229 / 0.5) + (0.485 - 0.5) / 0.5
        x[1] = x[1] * (0.224 / 0.5) + (0.456 - 0.5) / 0.5
        x[2] = x[2] * (0.225 / 0.5) + (0.406 - 0.5) / 0.5
        return x
    if variant == "ois":
        x = (lv / 255).view(1, 256).repeat(3, 1)
        return (x - 0.5) / 0.5
    if variant == "bird":
        x = torch.from_numpy(np.arange(256).astype(np.float32) / 127.5 - 1.0).to(torch.float32)
        return x.view(1, 256).repeat(3, 1)
    if variant == "coco":
        x = (lv - 128.0) * (1.0 / 128.0)
        return x.view(1, 256).repeat(3, 1)
    raise ValueError(variant)


def preprocess(imgs: Sequence[np.ndarray], variant: str) -> torch.Tensor:
    """list of uint8 HWC images -> float32 [B,3,299,299] network input (A2+A3+A4)."""
    lut = input_lut(variant)
    out = torch.empty(len(imgs), 3, 299, 299, dtype=torch.float32)
    for i, im in enumerate(imgs):
        r = torch.from_numpy(pillow_resize_bilinear(np.asarray(im))).long()  # [299,299,3]
        for c in range(3):
            out[i, c] = lut[c][r[:, :, c]]
    return out


# --------------------------------------------------------------------------------------
# A5  the Inception-v3 trunk (torchvision Inception3 -- the third-party module the reference
#     instantiates at image_realism/FID/inception.py:57-95 and
#     object_centric_inception_score.py:42-47)
# --------------------------------------------------------------------------------------
class Trunk:
    def __init__(self, state_dict: Dict[str, torch.Tensor], num_classes: int = 1000):
        from torchvision.models import Inception3

        net = Inception3(num_classes=num_classes, aux_logits=True, transform_input=False,
                         init_weights=False)
        net.load_state_dict(state_dict)
        net.eval()
        self.net = net

    @torch.no_grad()
    def blocks(self, x: torch.Tensor) -> List[torch.Tensor]:
        """Outputs of the four FID blocks (inception.py:59-95): maxpool1, maxpool2, Mixed_6e, pool3."""
        n = self.net
        outs = []
        y = n.Conv2d_1a_3x3(x); y = n.Conv2d_2a_3x3(y); y = n.Conv2d_2b_3x3(y)
        y = F.max_pool2d(y, kernel_size=3, stride=2); outs.append(y)
        y = n.Conv2d_3b_1x1(y); y = n.Conv2d_4a_3x3(y)
        y = F.max_pool2d(y, kernel_size=3, stride=2); outs.append(y)
        for m in (n.Mixed_5b, n.Mixed_5c, n.Mixed_5d, n.Mixed_6a, n.Mixed_6b, n.Mixed_6c, n.Mixed_6d, n.Mixed_6e):
            y = m(y)
        outs.append(y)
        for m in (n.Mixed_7a, n.Mixed_7b, n.Mixed_7c):
            y = m(y)
        outs.append(F.adaptive_avg_pool2d(y, (1, 1)))
        return outs

    @torch.no_grad()
    def pool3(self, x: torch.Tensor) -> torch.Tensor:
        return self.blocks(x)[3].reshape(x.shape[0], -1)

    @torch.no_grad()
    def logits(self, x: torch.Tensor) -> torch.Tensor:
        """Inception3.forward in eval mode: pool3 -> dropout(identity) -> fc (with bias)."""
        return self.net.fc(self.pool3(x))


# --------------------------------------------------------------------------------------
# A6-A9  FID statistics (image_realism/FID/fid_score.py)
# --------------------------------------------------------------------------------------
def get_activations(batches: Iterable[torch.Tensor], trunk: Trunk, batch_size: int, dims: int = 2048) -> np.ndarray:
    """fid_score.py:67-118 for dims=2048 with float batches already preprocessed ('fid' variant).
    ``batches`` must yield exactly-batch_size tensors (DataLoader(drop_last=True), :215-217)."""
    rows = []
    for b in batches:
        assert b.shape[0] == batch_size
        rows.append(trunk.pool3(b).cpu().numpy().reshape(batch_size, -1).astype(np.float64))
    return np.concatenate(rows, 0) if rows else np.empty((0, dims))


def calculate_activation_statistics(act: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """fid_score.py:194-195."""
    return np.mean(act, axis=0), np.cov(act, rowvar=False)


def calculate_frechet_distance(mu1, sigma1, mu2, sigma2, eps=1e-6):
    """fid_score.py:121-171 (scipy >= 1.13 dropped sqrtm's ``disp`` kwarg; the result is the same array)."""
    from scipy import linalg

    mu1, mu2 = np.atleast_1d(mu1), np.atleast_1d(mu2)
    sigma1, sigma2 = np.atleast_2d(sigma1), np.atleast_2d(sigma2)
    assert mu1.shape == mu2.shape, "Training and test mean vectors have different lengths"
    assert sigma1.shape == sigma2.shape, "Training and test covariances have different dimensions"
    diff = mu1 - mu2
    covmean = linalg.sqrtm(sigma1.dot(sigma2))
    if not np.isfinite(covmean).all():
        offset = np.eye(sigma1.shape[0]) * eps
        covmean = linalg.sqrtm((sigma1 + offset).dot(sigma2 + offset))
    if np.iscomplexobj(covmean):
        if not np.allclose(np.diagonal(covmean).imag, 0, atol=1e-3):
            raise ValueError("Imaginary component {}".format(np.max(np.abs(covmean.imag))))
        covmean = covmean.real
    return diff.dot(diff) + np.trace(sigma1) + np.trace(sigma2) - 2 * np.trace(covmean)


# --------------------------------------------------------------------------------------
# A11-A14  inception scores
# --------------------------------------------------------------------------------------
T_COCO = 0.9091363549232483   # inception_score_star_coco.py:107
T_BIRD = 0.5980541706085205   # inception_score_star_bird.py:192
T_OIS = 2.1737587451934814    # object_centric_inception_score.py:55


def ois_preds(logits: torch.Tensor) -> np.ndarray:
    """object_centric_inception_score.py:54-57: softmax(logits / T) -> float64 rows of preds."""
    x = logits.detach() / T_OIS
    return F.softmax(x, dim=1).data.cpu().numpy().astype(np.float64)


def ois_inception_score(preds: np.ndarray, splits: int) -> Tuple[float, float]:
    """object_centric_inception_score.py:69-81 (scipy.stats.entropy per image; tail N % splits dropped)."""
    from scipy.stats import entropy

    n = preds.shape[0]
    split_scores = []
    for k in range(splits):
        part = preds[k * (n // splits): (k + 1) * (n // splits), :]
        py = np.mean(part, axis=0)
        scores = [entropy(part[i, :], py) for i in range(part.shape[0])]
        split_scores.append(np.exp(np.mean(scores)))
    return float(np.mean(split_scores)), float(np.std(split_scores))


def coco_preds(pool3: torch.Tensor, w: torch.Tensor) -> np.ndarray:
    """inception_score_star_coco.py:104-108: matmul(pool3, W) (no bias), / T, softmax, float32."""
    logits = pool3.detach() @ w.detach().t()
    logits = logits / torch.tensor(T_COCO, dtype=torch.float32)
    return F.softmax(logits, dim=1).numpy()


def bird_preds(logits51: torch.Tensor) -> np.ndarray:
    """inception_score_star_bird.py:187-194: drop background class 0, / T, softmax, float32."""
    known = logits51.detach()[:, 1:]
    known = known / torch.tensor(T_BIRD, dtype=torch.float32)
    return F.softmax(known, dim=1).numpy()


def isstar_inception_score(preds: np.ndarray, splits: int = 10) -> Tuple[float, float]:
    """inception_score_star_coco.py:53-60 == inception_score_star_bird.py:98-108 (float32 numpy)."""
    scores = []
    for i in range(splits):
        part = preds[(i * preds.shape[0] // splits): ((i + 1) * preds.shape[0] // splits), :]
        kl = part * (np.log(part) - np.log(np.expand_dims(np.mean(part, 0), 0)))
        kl = np.mean(np.sum(kl, 1))
        scores.append(np.exp(kl))
    return np.mean(scores).item(), np.std(scores).item()
