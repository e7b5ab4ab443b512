"""ctypes binding of libtise_b200.so (include/tise_b200.h).  There is no fallback: if the
library is missing and cannot be built, importing a compute entry point raises."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_LIB = None

c_ctx_p = C.c_void_p
c_stream = C.c_void_p


class WeightDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.POINTER(C.c_float)), ("numel", C.c_int64)]


class PngInfo(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("bit_depth", C.c_int32), ("color_type", C.c_int32),
                ("interlace", C.c_int32), ("channels", C.c_int32), ("n_idat", C.c_int32), ("supported", C.c_int32),
                ("idat_bytes", C.c_int64)]


class JpegInfo(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("components", C.c_int32), ("chroma_mode", C.c_int32),
                ("restart_interval", C.c_int32), ("supported", C.c_int32), ("progressive", C.c_int32), ("scans", C.c_int32), ("huffman_tables", C.c_int32), ("pad", C.c_int32),
                ("scan_bytes", C.c_int64),
                ("blocks", C.c_int64)]


class TiseError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__("libtise_b200: %s (status %d)" % (message, status))
        self.status = status


STATUS_INVALID, STATUS_CUDA, STATUS_WORKSPACE, STATUS_STATE, STATUS_NUMERIC = -1, -2, -3, -4, -5

_SIGNATURES = {
    "tise_version": (C.c_int, []),
    "tise_status_string": (C.c_char_p, [C.c_int]),
    "tise_ctx_create": (C.c_int, [C.c_int, C.POINTER(c_ctx_p)]),
    "tise_ctx_destroy": (None, [c_ctx_p]),
    "tise_last_error": (C.c_char_p, [c_ctx_p]),
    "tise_set_conv_mode": (C.c_int, [c_ctx_p, C.c_int, C.c_int]),
    "tise_set_pass_policy": (C.c_int, [c_ctx_p, C.c_int]),
    "tise_calibrate_single_pass": (C.c_int, [c_ctx_p, C.c_int, C.c_void_p, C.c_size_t, c_stream]),
    "tise_load_inception": (C.c_int, [c_ctx_p, C.POINTER(WeightDesc), C.c_int, C.c_int]),
    "tise_resize_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "tise_resize_u8": (C.c_int, [c_ctx_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                 C.c_void_p, C.c_size_t, c_stream]),
    "tise_crop_resize_u8": (C.c_int, [c_ctx_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_size_t, c_stream]),
    "tise_png_probe": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(PngInfo)]),
    "tise_png_decode_workspace_bytes": (C.c_size_t, [C.POINTER(PngInfo), C.c_int]),
    "tise_png_decode": (C.c_int, [c_ctx_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_size_t, c_stream]),
    "tise_jpeg_probe": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(JpegInfo)]),
    "tise_jpeg_decode_workspace_bytes": (C.c_size_t, [C.POINTER(JpegInfo), C.c_int]),
    "tise_jpeg_decode": (C.c_int, [c_ctx_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_size_t, c_stream]),
    "tise_file_sizes": (C.c_int, [C.POINTER(C.c_char_p), C.c_int, C.c_void_p, C.c_int]),
    "tise_read_files": (C.c_int, [C.POINTER(C.c_char_p), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                  C.POINTER(C.c_int)]),
    "tise_probe_many": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(PngInfo), C.POINTER(JpegInfo),
                                  C.c_void_p, C.c_void_p, C.c_int]),
    "tise_upsample_bilinear_f32": (C.c_int, [c_ctx_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                             c_stream]),
    "tise_channel_affine_f32": (C.c_int, [c_ctx_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                          c_stream]),
    "tise_set_progress_event": (C.c_int, [c_ctx_p, C.c_void_p, C.c_int]),
    "tise_set_input_lut": (C.c_int, [c_ctx_p, C.c_void_p]),
    "tise_inception_workspace_bytes": (C.c_size_t, [c_ctx_p, C.c_int]),
    "tise_inception_features": (C.c_int, [c_ctx_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_size_t, c_stream]),
    "tise_logits": (C.c_int, [c_ctx_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, c_stream]),
    "tise_moments_update": (C.c_int, [c_ctx_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, c_stream]),
    "tise_moments_finalize": (C.c_int, [c_ctx_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                        c_stream]),
    "tise_is_update": (C.c_int, [c_ctx_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, c_stream]),
    "tise_calibration_eval": (C.c_int, [c_ctx_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_int,
                                        C.c_void_p, c_stream]),
    "tise_frechet_workspace_bytes": (C.c_size_t, [C.c_int]),
    "tise_frechet": (C.c_int, [c_ctx_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                               C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p, C.c_size_t, c_stream]),
    "tise_debug_conv": (C.c_int, [c_ctx_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_int, C.c_int, C.c_void_p]),
    "tise_profile_enable": (C.c_int, [c_ctx_p, C.c_int]),
    "tise_profile_read": (C.c_int, [c_ctx_p, C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.POINTER(C.c_double)]),
    "tise_launch_count": (C.c_longlong, [c_ctx_p]),
}

EXPORTS = tuple(_SIGNATURES)


def lib_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True):
    """Returns the loaded CDLL; builds it with nvcc first if it is missing or stale."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB
    if build_if_missing:
        try:
            path = _build.build_library()
        except RuntimeError:
            if not os.path.exists(path):
                raise
    if not os.path.exists(path):
        raise RuntimeError("libtise_b200.so is missing (%s); run `python -m tise_toolbox_b200.build`. "
                           "There is no CPU fallback." % path)
    lib = C.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = the .so does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(ctx, status: int):
    if status == 0:
        return
    lib = load()
    msg = lib.tise_last_error(ctx).decode() if ctx else ""
    if not msg:
        msg = lib.tise_status_string(status).decode()
    raise TiseError(status, msg)
