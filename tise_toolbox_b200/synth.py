"""Synthetic inputs for the image-realism hot path (harness, not product).

The reference ships neither images nor weights, so both sides of every
parity test -- the CPU oracle and the CUDA path -- are fed from here:

* ``images(seed, set_id, start, count, h, w)``: counter-based uint8 HWC images.  All the
  arithmetic is integer (int64 tensors, 32-bit mixing hash), so the same call produces
  bit-identical images on the CPU and on a CUDA device; any rank can materialise any
  index range of its shard.  Images are low-frequency colour fields (a 5x5 control grid,
  fixed-point bilinear upsampled) with per-image gain, per-channel offset and noise --
  white noise alone gives almost no feature variance between images.
  ``set_id`` 0/1 are two visibly different distributions for the two sides of an FID.
* ``inception_state_dict(seed, num_classes, fc_bias)``: a seeded, variance-preserving
  torchvision ``Inception3`` state_dict.  torchvision's own default init makes pool3
  explode to ~1e10 and the reference then raises ``ValueError``.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

_M32 = 0xFFFFFFFF


def _mix(h: torch.Tensor) -> torch.Tensor:
    """32-bit avalanche hash on int64 tensors (values kept in [0, 2^32))."""
    h = h & _M32
    h = h ^ (h >> 15)
    h = (h * 0x85EBCA77) & _M32
    h = h ^ (h >> 13)
    h = (h * 0xC2B2AE3D) & _M32
    h = h ^ (h >> 16)
    return h


def _hash(*parts) -> torch.Tensor:
    h = None
    for p in parts:
        p = p if torch.is_tensor(p) else torch.tensor(int(p), dtype=torch.int64)
        h = _mix(p + 0x9E3779B1) if h is None else _mix(h * 0x9E3779B1 + p + 0x7F4A7C15)
    return h


def images(seed: int, set_id: int, start: int, count: int, h: int = 256, w: int = 256,
           device="cpu") -> torch.Tensor:
    """uint8 [count, h, w, 3] images with global indices start..start+count-1."""
    dev = torch.device(device)
    i64 = dict(dtype=torch.int64, device=dev)
    idx = torch.arange(start, start + count, **i64).view(-1, 1, 1, 1)
    base = _hash(torch.tensor(seed, **i64), torch.tensor(set_id, **i64), idx)  # [n,1,1,1]
    c = torch.arange(3, **i64).view(1, 1, 1, 3)

    # 5x5 control grid of colours
    gy = torch.arange(5, **i64).view(1, 5, 1, 1)
    gx = torch.arange(5, **i64).view(1, 1, 5, 1)
    grid = _hash(base, gy * 5 + gx, c) & 255  # [n,5,5,3]

    # 8.8 fixed point positions in grid space (4 cells across the image)
    fy = (torch.arange(h, **i64) * 4 * 256) // h
    fx = (torch.arange(w, **i64) * 4 * 256) // w
    iy, ay = (fy >> 8).clamp(max=3), None
    ay = fy - (iy << 8)
    ix = (fx >> 8).clamp(max=3)
    ax = fx - (ix << 8)
    g00 = grid[:, iy][:, :, ix]
    g01 = grid[:, iy][:, :, ix + 1]
    g10 = grid[:, iy + 1][:, :, ix]
    g11 = grid[:, iy + 1][:, :, ix + 1]
    ax_ = ax.view(1, 1, w, 1)
    ay_ = ay.view(1, h, 1, 1)
    top = g00 * (256 - ax_) + g01 * ax_
    bot = g10 * (256 - ax_) + g11 * ax_
    v = (top * (256 - ay_) + bot * ay_) >> 16  # 0..255

    gain = 128 + (_hash(base, 1001) & 255)  # 0.5 .. 1.5 (x256)
    off = (_hash(base, 2002, c) & 63) - 32
    if set_id:
        off = off + ((_hash(base, 3003, c) & 255) * 77 >> 8) * (1 if set_id % 2 else -1)
        gain = gain + 24 * set_id

    y = torch.arange(h, **i64).view(1, h, 1, 1)
    x = torch.arange(w, **i64).view(1, 1, w, 1)
    nh = _hash(base, (y * w + x) * 3 + c + 4004)
    s = (nh & 255) + ((nh >> 8) & 255) + ((nh >> 16) & 255) + ((nh >> 24) & 255) - 510
    noise = (s * 22) >> 8  # std ~ 12.7 grey levels

    out = ((v * gain) >> 8) + off + noise
    return out.clamp_(0, 255).to(torch.uint8)


def crop_sizes(seed: int, start: int, count: int, lo: int = 16, hi: int = 480):
    """Deterministic log-uniform (h, w) pairs for the O-FID / O-IS variable-size crops."""
    out = []
    span = math.log(hi / lo)
    for i in range(start, start + count):
        hh = int(_hash(seed, 5005, i)) / 2 ** 32
        ww = int(_hash(seed, 6006, i)) / 2 ** 32
        out.append((int(lo * math.exp(span * hh)), int(lo * math.exp(span * ww))))
    return out


def inception_state_dict(seed: int = 0, num_classes: int = 1000, fc_std: float = 0.3,
                         aux_logits: bool = True) -> Dict[str, torch.Tensor]:
    """Seeded torchvision ``Inception3`` state_dict (CPU fp32).

    conv: He-normal sigma = sqrt(2 / fan_in); BN: gamma ~ U(0.5, 1.5), beta ~ N(0, 0.2^2),
    running_mean ~ N(0, 0.1^2), running_var ~ U(0.5, 1.5); fc ~ N(0, fc_std^2) so that the
    inception score of the synthetic sets is well away from 1 (SURVEY.md §7 'hard parts').
    """
    from torchvision.models import Inception3

    g = torch.Generator(device="cpu")
    g.manual_seed(1000003 * seed + 17)
    net = Inception3(num_classes=num_classes, aux_logits=aux_logits, transform_input=False,
                     init_weights=False)
    sd = net.state_dict()
    out = {}
    for name, t in sd.items():
        if name.endswith("num_batches_tracked"):
            out[name] = torch.zeros_like(t)
            continue
        shape = tuple(t.shape)
        if name.endswith("conv.weight"):
            fan_in = shape[1] * shape[2] * shape[3]
            v = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
        elif name.endswith("bn.weight"):
            v = torch.rand(shape, generator=g) + 0.5
        elif name.endswith("bn.bias"):
            v = torch.randn(shape, generator=g) * 0.2
        elif name.endswith("running_mean"):
            v = torch.randn(shape, generator=g) * 0.1
        elif name.endswith("running_var"):
            v = torch.rand(shape, generator=g) + 0.5
        elif name.endswith("fc.weight"):
            v = torch.randn(shape, generator=g) * fc_std
        elif name.endswith("fc.bias"):
            v = torch.randn(shape, generator=g) * 0.5
        else:  # pragma: no cover - Inception3 has no other parameter kinds
            v = torch.randn(shape, generator=g) * 0.01
        out[name] = v.to(torch.float32).contiguous()
    return out


def calibrate_head(state_dict: Dict[str, torch.Tensor], feats: torch.Tensor, temperature: float,
                   target_std: float = 2.5, seed: int = 0, class_offset: int = 0) -> Dict[str, torch.Tensor]:
    """Returns a copy of ``state_dict`` whose fc head makes the inception score informative.

    With a random head the image-independent part of the logits, W . mean(pool3), dwarfs the
    image-dependent part (post-ReLU features have mean >> std), every image gets the same arg-max and
    IS ~ 1.0000x -- a vacuous parity target.  Given pool3 features ``feats`` [n, 2048] of
    some calibration images, the fc rows are redrawn orthogonal to the mean feature and scaled so that
    logits / temperature vary with ``target_std`` across images; the bias is small noise."""
    g = torch.Generator(device="cpu")
    g.manual_seed(7919 * seed + 13)
    c = int(state_dict["fc.weight"].shape[0])
    f = feats.detach().cpu().double()
    m = f.mean(0)
    w = torch.randn(c, f.shape[1], generator=g, dtype=torch.float64)
    w = w - (w @ m)[:, None] * m[None, :] / (m @ m)
    z = (f - m) @ w.t()
    w = w * (target_std * temperature / z[:, class_offset:].std(0).mean().clamp_min(1e-12))
    out = dict(state_dict)
    out["fc.weight"] = w.float().contiguous()
    out["fc.bias"] = (torch.randn(c, generator=g) * 0.1).float()
    return out


def calibration_logits(seed: int, n: int, num_classes: int, sharpness: float = 3.0, label_noise: float = 0.25):
    """Seeded validation logits / labels for the temperature-calibration path: an over-confident classifier
    (logit gap ``sharpness`` towards a class that is the true label only with probability 1 - label_noise), so
    that the optimal temperature is well above 1 and NLL(T) has an interior minimum."""
    rng = np.random.RandomState(seed)
    labels = rng.randint(0, num_classes, size=n).astype(np.int64)
    guess = np.where(rng.rand(n) < label_noise, rng.randint(0, num_classes, size=n), labels)
    logits = rng.randn(n, num_classes).astype(np.float32)
    logits[np.arange(n), guess] += np.float32(sharpness) * (1.0 + rng.rand(n).astype(np.float32))
    return torch.from_numpy(logits), torch.from_numpy(labels)
