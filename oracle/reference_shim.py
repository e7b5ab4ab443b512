"""Runs the reference's OWN source files (unmodified, from /root/reference) on the CPU.

TEST INFRASTRUCTURE ONLY.  In the build container the files are read from /root/reference; on the GPU
box (where that tree does not exist) from the unmodified copies under git-ignored ``oracle/_ref/``
(``oracle/build_ref.py``, run by ``__graft_entry__.build()``).  ``tools/make_golden.py`` uses this module to (a) validate ``oracle/restate.py`` against
the real reference and (b) write the committed fixtures under ``tests/golden/``.

Shims -- none of them touches the arithmetic of the path:
  1. ``torchvision.models.inception_v3(pretrained=True)`` would download ImageNet weights
     (image_realism/FID/inception.py:57, object_centric_inception_score.py:42): replaced by a
     constructor that builds the same ``Inception3`` and loads the harness state_dict.
  2. ``scipy.linalg.sqrtm(..., disp=False)`` (fid_score.py:155): scipy >= 1.13 dropped ``disp`` and
     returns the array only; wrapped to return ``(sqrtm(a), 0.0)`` as scipy 1.1.0 did.
  3. ``torch.load("weights/...pth")`` (object_centric_inception_score.py:46) returns the harness
     state_dict; the module-level argparse/main of that script (:107-129) is not executed.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root() -> str:
    """The reference tree: $TISE_REFERENCE_ROOT, else /root/reference (build container), else the unmodified copies
    of the path's nine files that ``oracle/build_ref.py`` placed under git-ignored ``oracle/_ref/`` (GPU box)."""
    for cand in (os.environ.get("TISE_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "image_realism", "FID")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "image_realism", "FID"))


def _patched_inception_v3(state_dict, num_classes):
    from torchvision.models import Inception3

    def ctor(pretrained=True, transform_input=False, **kw):
        net = Inception3(num_classes=num_classes, aux_logits=True, transform_input=transform_input,
                         init_weights=False)
        net.load_state_dict(state_dict)
        return net

    return ctor


def load_fid_module(state_dict, num_classes=1000):
    """Returns the reference's ``fid_score`` module (image_realism/FID) ready to run on the CPU."""
    import scipy.linalg
    import torchvision.models as tvm

    fid_dir = os.path.join(REFERENCE_ROOT, "image_realism", "FID")
    tvm.inception_v3 = _patched_inception_v3(state_dict, num_classes)
    for name in ("fid_score", "inception", "img_data"):
        sys.modules.pop(name, None)
    sys.path.insert(0, fid_dir)
    argv = sys.argv
    sys.argv = [argv[0]]
    try:
        mod = importlib.import_module("fid_score")
    finally:
        sys.argv = argv
        sys.path.remove(fid_dir)
    real_sqrtm = scipy.linalg.sqrtm

    def sqrtm(a, disp=True, blocksize=64):
        r = real_sqrtm(a)
        return r if disp else (r, 0.0)

    mod.linalg = types.SimpleNamespace(sqrtm=sqrtm)
    return mod


def load_ois_function(state_dict, num_classes=80):
    """Returns the reference's O-IS ``inception_score`` function (object_centric_inception_score.py:17-81),
    compiled from the source lines themselves, with the weight loading patched."""
    import torch
    import torchvision.models as tvm

    path = os.path.join(REFERENCE_ROOT, "object_fidelity", "O-IS", "object_centric_inception_score.py")
    lines = open(path).read().split("\n")
    src = "\n".join(lines[1:15] + lines[16:104])  # imports + inception_score + IgnoreLabelDataset, no main
    ns: dict = {}
    tvm.inception_v3 = _patched_inception_v3(state_dict, num_classes)
    real_load = torch.load
    torch.load = lambda *a, **k: state_dict
    try:
        exec(compile(src, path, "exec"), ns)
    finally:
        torch.load = real_load

    def run(imgs, batch_size=32, splits=10):
        torch.load = lambda *a, **k: state_dict
        try:
            return ns["inception_score"](imgs, cuda=False, batch_size=batch_size, resize=False, splits=splits)
        finally:
            torch.load = real_load

    return run, ns["IgnoreLabelDataset"]
