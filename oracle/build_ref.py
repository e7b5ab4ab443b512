#!/usr/bin/env python3
"""Recipe for ``oracle/_ref/``: the reference's OWN source files of the hot path, copied UNMODIFIED from
``/root/reference`` (TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``).

The reference is pure Python: there is nothing to compile, so "building" the real reference for the
GPU box means placing the nine files the path consists of where ``oracle/reference_shim.py`` can
import them.  ``oracle/_ref/`` is git-ignored (no reference source ever enters the history) but travels with the
``gpurun`` snapshot like the built ``.so``; ``__graft_entry__.build()`` runs this whenever ``/root/reference`` is
present.  ``bench.py --impl reference`` and the ``cpu_baseline`` leg then time the reference's own
``get_activations`` / ``calculate_activation_statistics`` / ``calculate_frechet_distance`` (``kind: "reference"``);
without the directory they fall back to ``oracle/restate.py`` (``kind: "port"``).

    python oracle/build_ref.py [reference_root]
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

FILES = [  # relative to the reference root; SURVEY.md §8a rows A1-A13
    "image_realism/FID/fid_score.py",
    "image_realism/FID/inception.py",
    "image_realism/FID/img_data.py",
    "object_fidelity/O-FID/fid_score.py",
    "object_fidelity/O-FID/inception.py",
    "object_fidelity/O-FID/img_data.py",
    "object_fidelity/O-IS/object_centric_inception_score.py",
    "image_realism/IS/coco/inception_score_star_coco.py",
    "image_realism/IS/bird/inception_score_star_bird.py",
]


def build(reference_root: str = "/root/reference") -> str:
    if not os.path.isdir(os.path.join(reference_root, "image_realism", "FID")):
        raise FileNotFoundError("reference tree not found at %s" % reference_root)
    manifest = {}
    for rel in FILES:
        src = os.path.join(reference_root, rel)
        dst = os.path.join(REF_DIR, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(REF_DIR, "MANIFEST.json"), "w") as f:
        json.dump({"source": reference_root, "sha256": manifest}, f, indent=1)
    return REF_DIR


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "image_realism", "FID", "fid_score.py"))


if __name__ == "__main__":
    print(build(sys.argv[1] if len(sys.argv) > 1 else "/root/reference"))
