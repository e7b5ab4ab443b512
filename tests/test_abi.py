"""CPU (-m "not gpu"): the C-ABI library builds for sm_100a, loads, and exports every symbol
include/tise_b200.h declares.  No compute call is made (there is no GPU here)."""
import ctypes
import os
import re

import pytest

from tise_toolbox_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "tise_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tise_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.EXPORTS)


def test_library_builds_and_exports_every_symbol():
    path = build.build_library()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    loaded = _lib.load()
    assert loaded.tise_version() >= 100
    assert loaded.tise_status_string(-3) == b"workspace too small"


def test_no_device_means_loud_failure():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _lib.load()
    ctx = ctypes.c_void_p()
    assert lib.tise_ctx_create(0, ctypes.byref(ctx)) == _lib.STATUS_CUDA  # no CPU fallback
    from tise_toolbox_b200.engine import Engine

    with pytest.raises(RuntimeError):
        Engine(None)


def test_sass_contains_blackwell_instructions():
    import shutil
    import subprocess

    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not found")
    sass = subprocess.run(["cuobjdump", "-sass", build.build_library()], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass       # tcgen05.mma
    assert "UTMALDG" in sass and "IM2COL" in sass  # TMA (im2col mode)
    assert "LDTM" in sass          # tcgen05.ld
    assert "HGMMA" not in sass
