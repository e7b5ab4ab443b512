"""Inception-score host logic shared by the three variants of the reference:

  ois  -- object_fidelity/O-IS/object_centric_inception_score.py:17-81 (80-way fc with bias, T = 2.1737...,
          split k = preds[k*(N//S):(k+1)*(N//S)], tail dropped, scipy.stats.entropy in float64)
  coco -- image_realism/IS/coco/inception_score_star_coco.py:32-60,104-108 (1008-way matmul, no bias,
          T = 0.9091..., split i = preds[i*N//S:(i+1)*N//S], float32 numpy log => NaN on p == 0)
  bird -- image_realism/IS/bird/inception_score_star_bird.py:74-111,187-194 (51-way fc with bias, class 0
          dropped before the softmax, T = 0.5980..., floor(N/64)*64 images kept, same splits as coco)

The device side (libtise_b200 ``tise_is_update``) produces per-split partial sums that are additive over
batches and ranks; this module only decides which split each global image index belongs to and
finalises exp(mean KL).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from .engine import Engine, ISPartials, T_BIRD, T_COCO, T_OIS

VARIANTS = {
    #        temperature, class_offset, use_bias, nan_on_zero
    "ois": (T_OIS, 0, True, False),
    "coco": (T_COCO, 0, False, True),
    "bird": (T_BIRD, 1, True, True),
}


def split_ids(n_total: int, splits: int, variant: str, start: int = 0, count: Optional[int] = None) -> np.ndarray:
    """int32 split id of global image indices start..start+count-1 out of n_total (-1 = not scored)."""
    count = n_total - start if count is None else count
    idx = np.arange(start, start + count, dtype=np.int64)
    if variant == "ois":
        per = n_total // splits  # object_centric_inception_score.py:73 -- the N % splits tail is dropped
        sid = np.where(per > 0, idx // max(per, 1), -1)
        sid = np.where(sid < splits, sid, -1)
    else:
        # inception_score_star_coco.py:55: split i covers [i*N//S, (i+1)*N//S)
        bounds = (np.arange(1, splits + 1, dtype=np.int64) * n_total) // splits
        sid = np.searchsorted(bounds, idx, side="right")
        sid = np.where(idx < n_total, sid, -1)
    return sid.astype(np.int32)


def split_ids_device(n_total: int, splits: int, variant: str, start: int, count: int, device) -> torch.Tensor:
    """:func:`split_ids` evaluated on ``device`` (integer index bookkeeping only): no host array crosses to the GPU, so
    the call is asynchronous -- a pageable host-to-device copy would synchronise the stream first and stall the host
    behind the whole trunk of the batch."""
    idx = torch.arange(start, start + count, dtype=torch.int64, device=device)
    if variant == "ois":
        per = n_total // splits
        if per <= 0:
            return torch.full((count,), -1, dtype=torch.int32, device=device)
        sid = idx // per
        sid = torch.where(sid < splits, sid, torch.full_like(sid, -1))
    else:
        bounds = (torch.arange(1, splits + 1, dtype=torch.int64, device=device) * n_total) // splits
        sid = torch.searchsorted(bounds, idx, right=True)
        sid = torch.where(idx < n_total, sid, torch.full_like(sid, -1))
    return sid.to(torch.int32)


def kept_images(n: int, variant: str, batch_size: int = 64) -> int:
    """How many of n images enter the score: bird keeps floor(n / 64) full batches
    (inception_score_star_bird.py:83), the others keep all."""
    return (n // batch_size) * batch_size if variant == "bird" else n


def update_from_logits(engine: Engine, parts: ISPartials, logits: torch.Tensor, variant: str, n_total: int,
                       start: int):
    t, off, _, _ = VARIANTS[variant]
    sid = split_ids_device(n_total, parts.s, variant, start, int(logits.shape[0]), engine.device)
    parts.update(logits, sid, off, t)


def score_from_logits(engine: Engine, logits: torch.Tensor, variant: str, splits: int = 10,
                      n_total: Optional[int] = None, start: int = 0, group=None) -> Tuple[float, float]:
    """(mean, std) of the inception score of ``logits`` [n, C] (device float32); with ``group`` the partial
    sums of all ranks are combined first (each rank passes its own ``start`` and the global ``n_total``)."""
    t, off, _, nan_on_zero = VARIANTS[variant]
    n_total = int(logits.shape[0]) if n_total is None else n_total
    parts = ISPartials(engine, int(logits.shape[1]) - off, splits)
    update_from_logits(engine, parts, logits, variant, n_total, start)
    if group is not None:
        parts.allreduce(group)
    mean, std, _ = parts.finalize(nan_on_zero=nan_on_zero)
    return mean, std
