"""Bulk helpers around ``oracle/restate.py`` for the parity tests at BASELINE sizes -- TEST INFRASTRUCTURE ONLY
(see ``oracle/__init__.py``).  Nothing here is restated arithmetic: the resize is the ``PIL.Image.resize`` call the
reference itself makes (fid_score.py:210 via torchvision ``transforms.Resize``; ``restate.pillow_resize_bilinear`` is
pinned bit-for-bit against it in tests/test_oracle.py), the network is torchvision ``Inception3`` in fp32.
"""
from __future__ import annotations

import os
from typing import Dict, Sequence

import numpy as np
import torch
from PIL import Image

from . import restate


def use_all_cores() -> int:
    """torchrun exports OMP_NUM_THREADS=1; the reference's CPU path uses every core (torch default)."""
    n = os.cpu_count() or 1
    if torch.get_num_threads() != n:
        torch.set_num_threads(n)
    return n


def pil_preprocess(imgs: Sequence[np.ndarray], variant: str) -> torch.Tensor:
    """uint8 HWC images -> float32 [B, 3, 299, 299]: PIL bilinear resize (A2) + the variant's input map (A3/A4)."""
    lut = restate.input_lut(variant)
    out = torch.empty(len(imgs), 3, 299, 299, dtype=torch.float32)
    for i, im in enumerate(imgs):
        r = torch.from_numpy(np.asarray(Image.fromarray(np.asarray(im)).resize((299, 299), Image.BILINEAR))).long()
        for c in range(3):
            out[i, c] = lut[c][r[:, :, c]]
    return out


def oracle_pool3(state_dict: Dict[str, torch.Tensor], imgs, variant: str = "fid", num_classes: int = 1000,
                 batch: int = 50, trunk: "restate.Trunk | None" = None) -> np.ndarray:
    """pool3 features [N, 2048] float32 of the CPU path, in the reference's batches of ``batch`` (fid_score.py:99-113)."""
    use_all_cores()
    trunk = trunk or restate.Trunk(state_dict, num_classes=num_classes)
    rows = []
    for i0 in range(0, len(imgs), batch):
        chunk = imgs[i0:i0 + batch]
        chunk = [np.asarray(c) for c in (chunk.numpy() if torch.is_tensor(chunk) else chunk)]
        rows.append(trunk.pool3(pil_preprocess(chunk, variant)).numpy())
    return np.concatenate(rows, 0)


def fid_from_features(fa: np.ndarray, fb: np.ndarray) -> float:
    """fid_score.py:194-195 + :121-171 on two feature matrices (float64 like ``get_activations`` returns)."""
    sa = restate.calculate_activation_statistics(fa.astype(np.float64))
    sb = restate.calculate_activation_statistics(fb.astype(np.float64))
    return float(restate.calculate_frechet_distance(sa[0], sa[1], sb[0], sb[1]))
