"""Host-side driver of the CUDA hot path: owns a libtise_b200 context, device workspaces and the
running FID / IS statistics.  PyTorch is used for device memory, streams and
``torch.distributed`` only; every arithmetic step is a kernel of libtise_b200.so.

The reference-facing modules (``fid_score``, ``inception``, ``object_centric_inception_score``,
``inception_score_star``) are thin wrappers over :class:`Engine`.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib

TAP_BY_DIM = {64: 0, 192: 1, 768: 2, 2048: 3}  # image_realism/FID/inception.py:14-19
INPUT_VARIANTS = ("fid", "ois", "bird", "coco")

T_COCO = 0.9091363549232483   # image_realism/IS/coco/inception_score_star_coco.py:107
T_BIRD = 0.5980541706085205   # image_realism/IS/bird/inception_score_star_bird.py:192
T_OIS = 2.1737587451934814    # object_fidelity/O-IS/object_centric_inception_score.py:55

# Operand policy of the convolution stack.  Weights are split offline into fp16 hi + lo planes (~22 bits).
#   strict  (single_pass_max_hw = 0): every layer multiplies its fp16 activations with W_hi AND W_lo, two tcgen05 passes.
#   default (single_pass_max_hw = SINGLE_PASS_ALL): the tcgen05 implicit-GEMM layers issue the W_hi pass only, and the
#           image-independent part of the dropped product, W_lo . mean(activation), is folded into the layer's shift
#           by a one-off calibration at load time (tise_calibrate_single_pass: synthetic images generated on the
#           device; the correction is a property of the weights, not of the caller's data).  What remains of the W_lo
#           term is zero-mean noise of the size of the fp16 activation rounding.
# Measured on B200 at BASELINE.json configs[0] size: see DESIGN.md §2 / profiles/.  TISE_PASS1_MAXHW overrides.
SINGLE_PASS_ALL = 1 << 20
DEFAULT_SINGLE_PASS_MAX_HW = SINGLE_PASS_ALL
CALIBRATION_IMAGES = 32
POOL3_REL_L2_BOUND = {"strict": 5e-4, "default": 5e-4, "uncorrected": 1e-3}


def input_lut(variant: str) -> torch.Tensor:
    """float32 [3, 256] table: network input for each (channel, uint8 level), built with the op
    sequence of the reference so the fused uint8 path is bit-identical to its float path:
    fid  -- ToTensor (fid_score.py:211) + affine of image_realism/FID/inception.py:120-124;
    ois  -- ToTensor + Normalize(.5, .5) (object_centric_inception_score.py:90-91);
    bird -- x / 127.5 - 1 (inception_score_star_bird.py:70);
    coco -- raw 0..255 into a graph that starts with (x - 128) / 128 (inception_score_star_coco.py:119)."""
    lv = torch.arange(256, dtype=torch.float32)
    if variant == "fid":
        x = (lv / 255).view(1, 256).repeat(3, 1)
        x[0] = x[0] * (0.229 / 0.5) + (0.485 - 0.5) / 0.5
        x[1] = x[1] * (0.224 / 0.5) + (0.456 - 0.5) / 0.5
        x[2] = x[2] * (0.225 / 0.5) + (0.406 - 0.5) / 0.5
        return x.contiguous()
    if variant == "ois":
        x = (lv / 255).view(1, 256).repeat(3, 1)
        return ((x - 0.5) / 0.5).contiguous()
    if variant == "bird":
        x = torch.from_numpy(np.arange(256).astype(np.float32) / 127.5 - 1.0).to(torch.float32)
        return x.view(1, 256).repeat(3, 1).contiguous()
    if variant == "coco":
        return ((lv - 128.0) * (1.0 / 128.0)).view(1, 256).repeat(3, 1).contiguous()
    raise ValueError("unknown input variant %r" % (variant,))


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class PackedImages:
    """n uint8 HWC RGB images of arbitrary sizes in ONE device buffer: image i is ``buf[offsets[i]:][:h*w*3]``.
    This is the layout tise_png_decode writes and tise_resize_u8 reads, so decode -> resize moves no bytes."""

    def __init__(self, buf: torch.Tensor, offsets: np.ndarray, hw: np.ndarray):
        self.buf = buf
        self.offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self.hw = np.ascontiguousarray(hw, dtype=np.int32).reshape(-1, 2)

    def __len__(self):
        return int(self.hw.shape[0])

    def __getitem__(self, i) -> torch.Tensor:
        h, w = int(self.hw[i, 0]), int(self.hw[i, 1])
        o = int(self.offsets[i])
        return self.buf[o:o + h * w * 3].view(h, w, 3)

    def tensors(self) -> List[torch.Tensor]:
        return [self[i] for i in range(len(self))]


class ImageDecodeError(ValueError):
    """A file the device decoder accepted turned out to be corrupt (status codes: csrc/png_core.h)."""


class Engine:
    """One libtise_b200 context on one GPU with an Inception-v3 ``state_dict`` loaded."""

    def __init__(self, state_dict: Optional[Dict[str, torch.Tensor]] = None,
                 device: Union[int, str, torch.device, None] = None, variant: str = "fid", micro_batch: int = 128,
                 conv_impl: str = "umma", passes: int = 2, single_pass_max_hw: Optional[int] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("tise_toolbox_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("tise_toolbox_b200 runs on CUDA devices only, got %s" % dev)
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self.micro_batch = int(micro_batch)
        ctx = _lib.c_ctx_p()
        _lib.check(None, self.lib.tise_ctx_create(dev.index, C.byref(ctx)))
        self.ctx = ctx
        self._ws: Optional[torch.Tensor] = None
        self._rs_ws: Optional[torch.Tensor] = None
        self._u8: Optional[torch.Tensor] = None
        self._fr_ws: Optional[torch.Tensor] = None
        self._png_ws: Optional[torch.Tensor] = None
        self._jpg_ws: Optional[torch.Tensor] = None
        self._d_files: Optional[torch.Tensor] = None
        self.num_classes = 0
        self.variant = None
        self.set_conv_mode(conv_impl, passes)
        if single_pass_max_hw is None:
            single_pass_max_hw = int(os.environ.get("TISE_PASS1_MAXHW", DEFAULT_SINGLE_PASS_MAX_HW))
        self.set_pass_policy(single_pass_max_hw)
        self.calibrated = False
        self._calibrate = os.environ.get("TISE_LO_CALIBRATION", "1") != "0"
        self.set_variant(variant)
        if state_dict is not None:  # statistics-only use (moments, Frechet distance) needs no weights
            self.load_state_dict(state_dict)

    # ------------------------------------------------------------------ setup
    def close(self):
        if getattr(self, "ctx", None):
            self.lib.tise_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _check(self, status: int):
        _lib.check(self.ctx, status)

    def set_conv_mode(self, conv_impl: str = "umma", passes: int = 2):
        self._check(self.lib.tise_set_conv_mode(self.ctx, {"umma": 0, "simt": 1}[conv_impl], int(passes)))
        self.conv_impl, self.passes = conv_impl, int(passes)

    def set_pass_policy(self, single_pass_max_hw: int = 0):
        """Convolutions whose input map is at most this many pixels high skip the W_lo pass (0: none = strict)."""
        self._check(self.lib.tise_set_pass_policy(self.ctx, int(single_pass_max_hw)))
        self.pass1_max_hw = int(single_pass_max_hw)

    def calibrate_single_pass(self, n_images: int = CALIBRATION_IMAGES):
        """Mean-field W_lo correction of the single-pass layers (tise_calibrate_single_pass); needs weights and LUT."""
        need = int(self.lib.tise_inception_workspace_bytes(self.ctx, n_images))
        ws = self._buf("_ws", need)
        with torch.cuda.device(self.device):
            self._check(self.lib.tise_calibrate_single_pass(self.ctx, int(n_images), ws.data_ptr(), ws.numel(), _stream()))
        self.calibrated = True

    def load_state_dict(self, state_dict: Dict[str, torch.Tensor]):
        keep, descs = [], []
        for name, t in state_dict.items():
            if name.endswith("num_batches_tracked") or name.startswith("AuxLogits"):
                continue  # the aux head is never evaluated in eval mode (torchvision Inception3._forward)
            a = np.ascontiguousarray(t.detach().cpu().to(torch.float32).numpy())
            keep.append(a)
            descs.append(_lib.WeightDesc(name.encode(), a.ctypes.data_as(C.POINTER(C.c_float)), a.size))
        num_classes = int(state_dict["fc.weight"].shape[0]) if "fc.weight" in state_dict else 0
        arr = (_lib.WeightDesc * len(descs))(*descs)
        with torch.cuda.device(self.device):
            self._check(self.lib.tise_load_inception(self.ctx, arr, len(descs), num_classes))
        self.num_classes = num_classes
        self.calibrated = False
        if self._calibrate and self.conv_impl == "umma":  # (cheap: one 32-image forward + ~90 tiny kernels)
            self.calibrate_single_pass()

    def set_variant(self, variant: str):
        if variant == self.variant:
            return
        lut = input_lut(variant).numpy()
        self._check(self.lib.tise_set_input_lut(self.ctx, lut.ctypes.data))
        self.variant = variant

    @property
    def launches(self) -> int:
        return int(self.lib.tise_launch_count(self.ctx))

    def profile(self, on: bool):
        self._check(self.lib.tise_profile_enable(self.ctx, int(bool(on))))

    def profile_read(self):
        """{kind: (ms, launches, algorithmic flops)} for kinds stem / conv / maxpool / avgpool / gap / avgpool_bn."""
        ms = (C.c_double * 8)()
        ln = (C.c_longlong * 8)()
        fl = (C.c_double * 8)()
        self._check(self.lib.tise_profile_read(self.ctx, ms, ln, fl))
        names = ["stem", "conv", "maxpool", "avgpool", "gap", "avgpool_bn"]
        return {n: (ms[i], int(ln[i]), fl[i]) for i, n in enumerate(names)}

    # ------------------------------------------------------------------ buffers
    def _buf(self, name: str, nbytes: int) -> torch.Tensor:
        cur = getattr(self, name)
        if cur is None or cur.numel() < nbytes:
            cur = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=self.device)
            setattr(self, name, cur)
        return cur

    # ------------------------------------------------------------------ A1 decode
    def png_probe(self, blob: bytes) -> Optional[_lib.PngInfo]:
        """IHDR / chunk walk of one file on the host; None if the bytes are not a PNG container."""
        info = _lib.PngInfo()
        rc = self.lib.tise_png_probe(blob, len(blob), C.byref(info))
        return info if rc == 0 else None

    def _pinned(self, nbytes: int) -> torch.Tensor:
        """Two grow-only pinned staging buffers used alternately (allocating pinned memory costs milliseconds)."""
        if not hasattr(self, "_pin"):
            self._pin, self._pin_next = [None, None], 0
        k = self._pin_next
        self._pin_next = 1 - k
        if self._pin[k] is None or self._pin[k].numel() < nbytes:
            self._pin[k] = torch.empty(max(int(nbytes * 1.25), 1 << 20), dtype=torch.uint8, pin_memory=True)
        return self._pin[k]

    def _probe_staged(self, pinned: torch.Tensor, f_off: np.ndarray, sizes: np.ndarray, threads: int = 4) -> dict:
        """Chunk / marker walk of every staged file on the host (ONE tise_probe_many call, the interpreter lock released
        meanwhile): kind 1 = PNG, 2 = JPEG the device decoders handle, 0 = anything else (decoded by the caller's host
        decoder)."""
        m = int(sizes.shape[0])
        infos = (_lib.PngInfo * max(m, 1))()
        jinfos = (_lib.JpegInfo * max(m, 1))()
        kind = np.zeros(max(m, 1), dtype=np.int8)
        hw = np.zeros((max(m, 1), 2), dtype=np.int32)
        use_jpeg = os.environ.get("TISE_GPU_JPEG", "1") != "0"
        f_off = np.ascontiguousarray(f_off, dtype=np.int64)
        sizes = np.ascontiguousarray(sizes, dtype=np.int64)
        if m:
            _lib.check(None, self.lib.tise_probe_many(pinned.data_ptr(), f_off.ctypes.data, sizes.ctypes.data, m, int(use_jpeg),
                                                      infos, jinfos, kind.ctypes.data, hw.ctypes.data, int(threads)))
        kind, hw = kind[:m], hw[:m]
        return {"pinned": pinned, "f_off": f_off, "sizes": sizes, "infos": infos, "jinfos": jinfos, "kind": kind, "hw": hw,
                "supported": kind != 0, "n": m, "total": int(f_off[-1] + sizes[-1]) + 16 if m else 16}

    @staticmethod
    def _layout(sizes: np.ndarray):
        f_off = np.zeros(sizes.shape[0], dtype=np.int64)
        if sizes.shape[0]:
            f_off[1:] = np.cumsum((sizes + 15) // 16 * 16)[:-1]
        return f_off, (int(f_off[-1] + sizes[-1]) + 16 if sizes.shape[0] else 16)

    def stage_blobs(self, blobs: Sequence[bytes]) -> dict:
        """Encoded files (bytes objects) -> one pinned host buffer + their chunk-level description (host work only)."""
        sizes = np.array([len(b) for b in blobs], dtype=np.int64)
        f_off, total = self._layout(sizes)
        pinned = self._pinned(total)
        hbuf = pinned.numpy()

        def pack(lo, hi):  # plain memcpys: numpy releases the GIL, so 8 threads reach the host's copy bandwidth
            for k in range(lo, hi):
                hbuf[f_off[k]:f_off[k] + sizes[k]] = np.frombuffer(blobs[k], dtype=np.uint8)

        m = len(blobs)
        if m >= 64:
            if not hasattr(self, "_io_pool"):
                from concurrent.futures import ThreadPoolExecutor

                self._io_pool = ThreadPoolExecutor(max_workers=8)
            step = (m + 7) // 8
            list(self._io_pool.map(lambda lo: pack(lo, min(m, lo + step)), range(0, m, step)))
        else:
            pack(0, m)
        return self._probe_staged(pinned, f_off, sizes)

    def stage_files(self, paths: Sequence[str], pool=None) -> dict:
        """Reads files straight into pinned memory: one tise_file_sizes and one tise_read_files call per group (POSIX
        reads on host threads inside the library, the interpreter lock released meanwhile -- per-file Python reads cost
        ~45 us each and scaled badly over threads), then one tise_probe_many.  No intermediate bytes objects, no second
        host copy.  Host work only -- safe to run on a worker thread.  ``pool``: its worker count sets the thread count."""
        m = len(paths)
        threads = int(getattr(pool, "_max_workers", 8)) if pool is not None else 8
        arr = (C.c_char_p * max(m, 1))(*[os.fsencode(p) for p in paths])
        sizes = np.zeros(max(m, 1), dtype=np.int64)
        if m and self.lib.tise_file_sizes(arr, m, sizes.ctypes.data, threads) != 0:
            bad = int(np.nonzero(sizes[:m] < 0)[0][0])
            raise FileNotFoundError("cannot stat %r" % (paths[bad],))
        sizes = sizes[:m]
        f_off, total = self._layout(sizes)
        pinned = self._pinned(total)
        if m:
            failed = C.c_int(-1)
            rc = self.lib.tise_read_files(arr, f_off.ctypes.data, sizes.ctypes.data, m, pinned.data_ptr(), threads,
                                          C.byref(failed))
            if rc != 0:
                raise IOError("short read of %s" % (paths[failed.value] if failed.value >= 0 else "a file of the group"))
        staged = self._probe_staged(pinned, f_off, sizes, threads=min(threads, 8))
        staged["names"] = list(paths)  # for error messages
        return staged

    def decode_staged_png(self, staged: dict, d_files: torch.Tensor, index: np.ndarray, out: torch.Tensor,
                          out_off: np.ndarray, jpeg: bool = False) -> torch.Tensor:
        """tise_png_decode (or tise_jpeg_decode with ``jpeg``) on the staged files ``index`` (already uploaded as
        ``d_files``) -> images at ``out[out_off[k]:]``; returns the device status vector (asynchronous: nothing is read
        back here)."""
        m = int(index.shape[0])
        cls, key = (_lib.JpegInfo, "jinfos") if jpeg else (_lib.PngInfo, "infos")
        if m == int(staged["n"]) and m > 0 and bool((index == np.arange(m)).all()):
            infos = staged[key]  # the whole group, in order: no copy
        else:  # gather the selected records with one numpy take instead of a Python loop over ctypes structures
            rec = np.frombuffer(staged[key], dtype=np.uint8).reshape(-1, C.sizeof(cls))
            infos = (cls * max(m, 1)).from_buffer_copy(np.ascontiguousarray(rec[index]).tobytes() if m else bytes(C.sizeof(cls)))
        if jpeg:
            ws_bytes, decode, ws_name = self.lib.tise_jpeg_decode_workspace_bytes, self.lib.tise_jpeg_decode, "_jpg_ws"
        else:
            ws_bytes, decode, ws_name = self.lib.tise_png_decode_workspace_bytes, self.lib.tise_png_decode, "_png_ws"
        f_off = np.ascontiguousarray(staged["f_off"][index])
        sizes = np.ascontiguousarray(staged["sizes"][index])
        out_off = np.ascontiguousarray(out_off, dtype=np.int64)
        ws = self._buf(ws_name, int(ws_bytes(infos, m)))
        status = torch.empty(m, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(decode(self.ctx, staged["pinned"].data_ptr(), d_files.data_ptr(), f_off.ctypes.data,
                               sizes.ctypes.data, m, out.data_ptr(), out_off.ctypes.data, status.data_ptr(), ws.data_ptr(),
                               ws.numel(), _stream()))
        return status

    def decode_staged(self, staged: dict, host_fallback: bool = True, host_decoder=None) -> PackedImages:
        """Staged files -> RGB8 HWC images in HBM (what ``PIL.Image.open(f).convert("RGB")`` returns, img_data.py:21).
        Non-interlaced PNGs are inflated / unfiltered on the device (tise_png_decode), baseline JPEGs entropy-decoded,
        inverse-transformed and colour-converted there (tise_jpeg_decode); any other file (progressive JPEG, interlaced
        PNG, other formats) is outside the device decoders and, with ``host_fallback``, is decoded by PIL exactly like
        the reference does and uploaded into the same pack.  ``host_decoder`` (bytes -> HWC uint8 array) replaces
        that PIL call for callers whose reference decodes differently (the IS* scripts, inception_score_star.py)."""
        n = staged["n"]
        kind = np.where(staged["supported"], staged["kind"], 0)  # (callers may veto files by clearing "supported")
        on_png, on_jpg, on_host = np.nonzero(kind == 1)[0], np.nonzero(kind == 2)[0], np.nonzero(kind == 0)[0]
        if len(on_host) and not host_fallback:
            raise ValueError("decode: file %d is neither a PNG nor a baseline JPEG the device decoders handle" % int(on_host[0]))
        hbuf = staged["pinned"].numpy()
        host_imgs = {}
        if len(on_host):
            import io

            from PIL import Image

            for i in on_host:
                o, sz = int(staged["f_off"][i]), int(staged["sizes"][i])
                blob = hbuf[o:o + sz].tobytes()
                host_imgs[int(i)] = (np.ascontiguousarray(host_decoder(blob), dtype=np.uint8) if host_decoder else
                                     np.array(Image.open(io.BytesIO(blob)).convert("RGB"), dtype=np.uint8))
        hw = staged["hw"].copy()
        for i, a in host_imgs.items():
            hw[i] = a.shape[:2]
        nbytes = hw[:, 0].astype(np.int64) * hw[:, 1] * 3
        out_off = np.zeros(n, dtype=np.int64)
        if n:
            out_off[1:] = np.cumsum((nbytes + 15) // 16 * 16)[:-1]
        total = int(out_off[-1] + nbytes[-1]) if n else 0
        out = torch.empty(max(total, 16), dtype=torch.uint8, device=self.device)
        if len(on_png) or len(on_jpg):
            # file bytes -> HBM on a private upload stream into a persistent buffer: the copy does not queue behind the
            # trunk kernels of the previous group on the compute stream (the buffer is free: the previous decode was
            # awaited below before this call)
            total = int(staged["total"])
            if not hasattr(self, "_upload_stream"):
                self._upload_stream = torch.cuda.Stream(device=self.device)
            before = self._d_files
            d_files = self._buf("_d_files", total)[:total]
            if self._d_files is not before:
                # a grown buffer comes from the compute stream's pool: kernels queued there may still read the block under
                # its previous identity, so the upload stream waits for them once
                self._upload_stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self._upload_stream):
                d_files.copy_(staged["pinned"][:total], non_blocking=True)
                uploaded = torch.cuda.Event()
                uploaded.record(self._upload_stream)
            torch.cuda.current_stream(self.device).wait_event(uploaded)
            runs = [(idx, self.decode_staged_png(staged, d_files, idx, out, out_off[idx], jpeg=j), "JPEG" if j else "PNG")
                    for idx, j in ((on_png, False), (on_jpg, True)) if len(idx)]
            for idx, status, what in runs:
                st = status.cpu().numpy()  # synchronises: the pinned staging buffer may be reused afterwards
                if st.any():
                    k = int(np.nonzero(st)[0][0])
                    names = staged.get("names")
                    where = "%r (file %d of the group)" % (names[int(idx[k])], int(idx[k])) if names else "file %d" % int(idx[k])
                    raise ImageDecodeError("corrupt %s stream in %s (decoder status %d)" % (what, where, int(st[k])))
        for i, a in host_imgs.items():
            o = int(out_off[i])
            out[o:o + a.size].copy_(torch.from_numpy(a.reshape(-1)))
        return PackedImages(out, out_off, hw)

    def decode_images(self, blobs: Sequence[bytes], host_fallback: bool = True) -> PackedImages:
        """Encoded image files as bytes objects -> :class:`PackedImages` (see :meth:`decode_staged`)."""
        return self.decode_staged(self.stage_blobs(blobs), host_fallback)

    def decode_files(self, paths: Sequence[str], host_fallback: bool = True, pool=None) -> PackedImages:
        """Image files on disk -> :class:`PackedImages`; bytes go file -> pinned memory -> HBM."""
        return self.decode_staged(self.stage_files(paths, pool), host_fallback)

    # ------------------------------------------------------------------ A2 resize
    def resize(self, images: Union[torch.Tensor, Sequence[torch.Tensor], "PackedImages"],
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """uint8 HWC RGB images on the device -> uint8 [n, 299, 299, 3] (Pillow-exact bilinear).
        ``images``: one [n, H, W, 3] tensor, a sequence of [H_i, W_i, 3] tensors (ragged crops), or a
        :class:`PackedImages` (the decoder's output, consumed in place)."""
        if isinstance(images, PackedImages):
            n = len(images)
            hw = images.hw.reshape(-1)
            offs = images.offsets
            src = images.buf
            offs_p, uniform = offs.ctypes.data, 0
            max_h, max_w = (int(images.hw[:, 0].max()), int(images.hw[:, 1].max())) if n else (1, 1)
        elif torch.is_tensor(images):
            assert images.dtype == torch.uint8 and images.dim() == 4 and images.shape[3] == 3
            imgs = images.contiguous()
            n, h, w = int(imgs.shape[0]), int(imgs.shape[1]), int(imgs.shape[2])
            hw = np.array([h, w], dtype=np.int32)
            offs_p, uniform, max_h, max_w = None, 1, h, w
            src = imgs
        else:
            n = len(images)
            hw = np.array([[int(t.shape[0]), int(t.shape[1])] for t in images], dtype=np.int32).reshape(-1)
            sizes = hw.reshape(-1, 2).astype(np.int64)
            nbytes = sizes[:, 0] * sizes[:, 1] * 3
            offs = np.zeros(n, dtype=np.int64)
            offs[1:] = np.cumsum(nbytes)[:-1]
            src = torch.cat([t.reshape(-1) for t in images]) if n else torch.empty(0, dtype=torch.uint8, device=self.device)
            assert src.dtype == torch.uint8
            offs_p, uniform = offs.ctypes.data, 0
            max_h, max_w = (int(sizes[:, 0].max()), int(sizes[:, 1].max())) if n else (1, 1)
        if out is None:
            out = torch.empty((n, 299, 299, 3), dtype=torch.uint8, device=self.device)
        if n == 0:
            return out
        need = int(self.lib.tise_resize_workspace_bytes(n, max_h, max_w))
        ws = self._buf("_rs_ws", need)
        with torch.cuda.device(self.device):
            self._check(self.lib.tise_resize_u8(self.ctx, src.data_ptr(), offs_p, hw.ctypes.data, uniform, n,
                                                out.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
        return out

    def resize_crops(self, images: Union[torch.Tensor, "PackedImages"], image_index: Sequence[int], boxes,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Crop hand-off (object_fidelity/crop_object.py:38-46 + the Resize of the O-FID / O-IS datasets):
        ``boxes`` [n, 4] = (x0, y0, x1, y1) in pixels (floats are rounded like ``PIL.Image.crop`` rounds);
        crop i is cut from ``images[image_index[i]]`` and resized to 299 x 299 in one kernel ->
        uint8 [n, 299, 299, 3], byte-identical to ``img.crop(box).resize((299, 299), BILINEAR)``."""
        b = np.asarray(boxes, dtype=np.float64).reshape(-1, 4)
        box = np.array([[int(round(float(v))) for v in row] for row in b], dtype=np.int32).reshape(-1, 4)
        idx = np.asarray(image_index, dtype=np.int64).reshape(-1)
        n = int(box.shape[0])
        assert idx.shape[0] == n
        if isinstance(images, PackedImages):
            src, offs_all, hw_all = images.buf, images.offsets, images.hw
        else:
            assert images.dtype == torch.uint8 and images.dim() == 4 and images.shape[3] == 3
            src = images.contiguous()
            m, h, w = int(src.shape[0]), int(src.shape[1]), int(src.shape[2])
            offs_all = np.arange(m, dtype=np.int64) * (h * w * 3)
            hw_all = np.tile(np.array([[h, w]], dtype=np.int32), (m, 1))
        if out is None:
            out = torch.empty((n, 299, 299, 3), dtype=torch.uint8, device=self.device)
        if n == 0:
            return out
        offs = np.ascontiguousarray(offs_all[idx])
        hw = np.ascontiguousarray(hw_all[idx]).astype(np.int32)
        box = np.ascontiguousarray(box)
        max_h = int(max(1, (box[:, 3] - box[:, 1]).max()))
        max_w = int(max(1, (box[:, 2] - box[:, 0]).max()))
        ws = self._buf("_rs_ws", int(self.lib.tise_resize_workspace_bytes(n, max_h, max_w)))
        with torch.cuda.device(self.device):
            self._check(self.lib.tise_crop_resize_u8(self.ctx, src.data_ptr(), offs.ctypes.data, hw.ctypes.data,
                                                     box.ctypes.data, n, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                                     _stream()))
        return out

    def upsample_float(self, x: torch.Tensor, align_corners: bool) -> torch.Tensor:
        """float32 [n, c, h, w] -> [n, c, 299, 299], bilinear (the reference's F.upsample / nn.Upsample on float
        inputs of another size, inception.py:117-118 / object_centric_inception_score.py:49)."""
        assert x.dtype == torch.float32 and x.dim() == 4
        x = x.to(self.device).contiguous()
        n, c, h, w = (int(v) for v in x.shape)
        out = torch.empty((n, c, 299, 299), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.tise_upsample_bilinear_f32(self.ctx, x.data_ptr(), n, c, h, w, int(bool(align_corners)),
                                                            out.data_ptr(), _stream()))
        return out

    def channel_affine(self, x: torch.Tensor, scale: Sequence[float], bias: Sequence[float]) -> torch.Tensor:
        """float32 [n, c, h, w] -> x * scale[c] + bias[c] (rounded multiply, then rounded add: inception.py:120-124)."""
        assert x.dtype == torch.float32 and x.dim() == 4 and len(scale) == len(bias) == int(x.shape[1])
        x = x.to(self.device).contiguous()
        n, c, h, w = (int(v) for v in x.shape)
        out = torch.empty_like(x)
        sc = np.asarray(scale, dtype=np.float32)
        bs = np.asarray(bias, dtype=np.float32)
        with torch.cuda.device(self.device):
            self._check(self.lib.tise_channel_affine_f32(self.ctx, x.data_ptr(), n, c, h * w, sc.ctypes.data, bs.ctypes.data,
                                                         out.data_ptr(), _stream()))
        return out

    # ------------------------------------------------------------------ A5 trunk
    def set_progress_event(self, event: Optional[torch.cuda.Event], max_hw: int = 17):
        """tise_set_progress_event: ``event`` is recorded in every trunk call when the forward pass leaves the HBM-bound
        front of the network (first convolution on a map of at most ``max_hw`` pixels); None clears it."""
        handle = None
        if event is not None:
            with torch.cuda.device(self.device):
                event.record()  # (materialises the cudaEvent_t)
            handle = C.c_void_p(event.cuda_event)
        self._check(self.lib.tise_set_progress_event(self.ctx, handle, int(max_hw)))
        self._progress_event = event  # keep it alive while the context refers to it

    def progress_event(self, max_hw: int = 17) -> torch.cuda.Event:
        """The engine's shared progress event (created and registered on first use): every pacing caller waits on this
        one, so jobs that share an engine do not unregister each other's events."""
        ev = getattr(self, "_progress_event", None)
        if ev is None:
            ev = torch.cuda.Event()
            self.set_progress_event(ev, max_hw)
        return ev

    def _trunk(self, u8: Optional[torch.Tensor], f32: Optional[torch.Tensor], n: int, tap: int, feat: torch.Tensor):
        need = int(self.lib.tise_inception_workspace_bytes(self.ctx, self.micro_batch))
        ws = self._buf("_ws", need)
        hook = getattr(self, "on_micro_batch", None)  # callable(index, count): host-side, after each micro-batch is enqueued
        count = (n + self.micro_batch - 1) // self.micro_batch
        with torch.cuda.device(self.device):
            for k, i0 in enumerate(range(0, n, self.micro_batch)):
                m = min(self.micro_batch, n - i0)
                self._check(self.lib.tise_inception_features(
                    self.ctx, u8[i0:].data_ptr() if u8 is not None else None,
                    f32[i0:].data_ptr() if f32 is not None else None, m, tap, feat[i0:].data_ptr(),
                    ws.data_ptr(), ws.numel(), _stream()))
                if hook is not None:
                    hook(k, count)

    def features_from_resized(self, u8_299: torch.Tensor, dims: int = 2048) -> torch.Tensor:
        """uint8 [n, 299, 299, 3] (already resized) -> float32 [n, dims] block features."""
        assert u8_299.dtype == torch.uint8 and tuple(u8_299.shape[1:]) == (299, 299, 3)
        n = int(u8_299.shape[0])
        feat = torch.empty((n, dims), dtype=torch.float32, device=self.device)
        self._trunk(u8_299.contiguous(), None, n, TAP_BY_DIM[dims], feat)
        return feat

    def features(self, images: Union[torch.Tensor, Sequence[torch.Tensor]], dims: int = 2048) -> torch.Tensor:
        """uint8 HWC images (any size) -> float32 [n, dims]: resize -> LUT -> Inception-v3 -> pooled block output."""
        return self.features_from_resized(self.resize(images), dims)

    def features_from_float(self, x: torch.Tensor, dims: int = 2048) -> torch.Tensor:
        """float32 [n, 3, 299, 299], already normalised (what Conv2d_1a_3x3 sees) -> float32 [n, dims]."""
        assert x.dtype == torch.float32 and tuple(x.shape[1:]) == (3, 299, 299)
        x = x.contiguous()
        n = int(x.shape[0])
        feat = torch.empty((n, dims), dtype=torch.float32, device=self.device)
        self._trunk(None, x, n, TAP_BY_DIM[dims], feat)
        return feat

    def logits(self, feat: torch.Tensor, use_bias: bool = True) -> torch.Tensor:
        assert feat.dtype == torch.float32 and feat.shape[1] == 2048
        feat = feat.contiguous()
        out = torch.empty((feat.shape[0], self.num_classes), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.tise_logits(self.ctx, feat.data_ptr(), int(feat.shape[0]), int(bool(use_bias)),
                                             out.data_ptr(), _stream()))
        return out

    # ------------------------------------------------------------------ A9 Frechet distance
    def frechet(self, mu1, sigma1, mu2, sigma2, eps: float = 1e-6) -> Tuple[float, int]:
        d = int(np.asarray(mu1).shape[0]) if not torch.is_tensor(mu1) else int(mu1.shape[0])
        to = lambda a: (a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))).to(
            self.device, torch.float64).contiguous()
        m1, s1, m2, s2 = to(mu1), to(sigma1), to(mu2), to(sigma2)
        ws = self._buf("_fr_ws", int(self.lib.tise_frechet_workspace_bytes(d)))
        out, flags = C.c_double(0.0), C.c_int(0)
        with torch.cuda.device(self.device):
            self._check(self.lib.tise_frechet(self.ctx, m1.data_ptr(), s1.data_ptr(), m2.data_ptr(), s2.data_ptr(), d,
                                              float(eps), C.byref(out), C.byref(flags), ws.data_ptr(), ws.numel(),
                                              _stream()))
        return float(out.value), int(flags.value)


_DEFAULT_ENGINES: Dict[int, Engine] = {}


def default_engine(device: Union[int, str, torch.device, None] = None) -> Engine:
    """A weight-less engine per device for the statistics entry points (moments, Frechet distance)."""
    if not torch.cuda.is_available():
        raise RuntimeError("tise_toolbox_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    idx = torch.cuda.current_device() if device is None else (torch.device(device).index or 0)
    if idx not in _DEFAULT_ENGINES:
        _DEFAULT_ENGINES[idx] = Engine(None, device=torch.device("cuda", idx))
    return _DEFAULT_ENGINES[idx]


class Moments:
    """Running sum_i x_i and sum_i x_i x_i^T (fp64, on the device) -> (mu, sigma) of np.mean / np.cov
    (fid_score.py:194-195).  Additive across batches and across ranks.  ``buf`` (1 + d + d*d float64, zeroed) lets a
    caller place the sums inside a larger packed buffer (:class:`StatsPack`: one all-reduce for everything)."""

    def __init__(self, engine: Engine, d: int = 2048, buf: Optional[torch.Tensor] = None):
        self.e, self.d = engine, d
        if buf is None:
            buf = torch.zeros(1 + d + d * d, dtype=torch.float64, device=engine.device)  # [count | sum | sxx]
        assert buf.dtype == torch.float64 and buf.numel() == 1 + d + d * d and buf.is_contiguous()
        self.buf = buf
        self.count = 0

    @staticmethod
    def numel(d: int = 2048) -> int:
        return 1 + d + d * d

    @property
    def sum(self):
        return self.buf[1:1 + self.d]

    @property
    def sxx(self):
        return self.buf[1 + self.d:]

    def update(self, feat: torch.Tensor):
        assert feat.dtype == torch.float32 and feat.shape[1] == self.d
        feat = feat.contiguous()
        n = int(feat.shape[0])
        with torch.cuda.device(self.e.device):
            self.e._check(self.e.lib.tise_moments_update(self.e.ctx, feat.data_ptr(), n, self.d, self.sum.data_ptr(),
                                                         self.sxx.data_ptr(), _stream()))
        self.count += n

    def before_reduce(self):
        self.buf[0] = float(self.count)

    def after_reduce(self):
        self.count = int(round(float(self.buf[0].item())))

    def allreduce(self, group=None):
        """The path's one exchange step: sum the packed moments over all ranks."""
        import torch.distributed as dist

        self.before_reduce()
        dist.all_reduce(self.buf, op=dist.ReduceOp.SUM, group=group)
        self.after_reduce()

    def finalize_device(self) -> Tuple[torch.Tensor, torch.Tensor]:
        mu = torch.empty(self.d, dtype=torch.float64, device=self.e.device)
        sigma = torch.empty((self.d, self.d), dtype=torch.float64, device=self.e.device)
        with torch.cuda.device(self.e.device):
            self.e._check(self.e.lib.tise_moments_finalize(self.e.ctx, self.sum.data_ptr(), self.sxx.data_ptr(),
                                                           self.count, self.d, mu.data_ptr(), sigma.data_ptr(),
                                                           _stream()))
        return mu, sigma

    def finalize(self) -> Tuple[np.ndarray, np.ndarray]:
        mu, sigma = self.finalize_device()
        return mu.cpu().numpy(), sigma.cpu().numpy()


class ISPartials:
    """Per-split partial sums of the inception score: sum_i sum_c p log p, sum_i p (marginal), count.
    mean_i KL(p_i || m) = S / n - sum_c m_c log m_c with m = marg / n  (one-pass form).
    ``fbuf`` (``numel(C, S)`` float64, zeroed) places the sums -- and, for the exchange step, a float64 copy of the
    integer counters -- inside a larger packed buffer (:class:`StatsPack`)."""

    def __init__(self, engine: Engine, n_classes: int, n_splits: int, fbuf: Optional[torch.Tensor] = None):
        self.e, self.cs, self.s = engine, n_classes, n_splits
        nf = n_splits + n_splits * n_classes
        if fbuf is None:
            fbuf = torch.zeros(self.numel(n_classes, n_splits), dtype=torch.float64, device=engine.device)
        assert fbuf.dtype == torch.float64 and fbuf.numel() == self.numel(n_classes, n_splits) and fbuf.is_contiguous()
        self.fbuf = fbuf
        self.f = fbuf[:nf]            # [splogp S | marg S*Cs]
        self.icopy = fbuf[nf:]        # [cnt S | zeros 1] as float64, only meaningful around the all-reduce
        self.i = torch.zeros(n_splits + 1, dtype=torch.int64, device=engine.device)  # the kernel's integer counters

    @staticmethod
    def numel(n_classes: int, n_splits: int) -> int:
        return n_splits + n_splits * n_classes + n_splits + 1

    def update(self, logits: torch.Tensor, split_id: torch.Tensor, class_offset: int, temperature: float):
        assert logits.dtype == torch.float32 and split_id.dtype == torch.int32
        logits = logits.contiguous()
        n, c = int(logits.shape[0]), int(logits.shape[1])
        assert c - class_offset == self.cs and split_id.numel() == n
        e = self.e
        with torch.cuda.device(e.device):
            e._check(e.lib.tise_is_update(e.ctx, logits.data_ptr(), n, c, int(class_offset), float(temperature),
                                          split_id.contiguous().data_ptr(), self.s, self.f.data_ptr(),
                                          self.f[self.s:].data_ptr(), self.i.data_ptr(), self.i[self.s:].data_ptr(),
                                          _stream()))

    def before_reduce(self):
        self.icopy.copy_(self.i)  # counts < 2^53: exact in float64

    def after_reduce(self):
        self.i.copy_(self.icopy.round())

    def allreduce(self, group=None):
        import torch.distributed as dist

        self.before_reduce()
        dist.all_reduce(self.fbuf, op=dist.ReduceOp.SUM, group=group)
        self.after_reduce()

    def finalize(self, nan_on_zero: bool = False) -> Tuple[float, float, np.ndarray]:
        """(mean, std (ddof=0), per-split scores).  ``nan_on_zero`` reproduces the IS* scripts' explicit
        log: a probability that underflowed to 0 makes the score NaN (inception_score_star_coco.py:56),
        whereas scipy.stats.entropy (O-IS) treats 0 log 0 as 0."""
        f = self.f.cpu().numpy()
        i = self.i.cpu().numpy()
        splogp, marg = f[:self.s], f[self.s:].reshape(self.s, self.cs)
        cnt, zeros = i[:self.s].astype(np.float64), int(i[self.s])
        scores = np.full(self.s, np.nan)
        for k in range(self.s):
            if cnt[k] <= 0:
                continue
            m = marg[k] / cnt[k]
            nz = m > 0
            scores[k] = np.exp(splogp[k] / cnt[k] - float(np.sum(m[nz] * np.log(m[nz]))))
        if nan_on_zero and zeros > 0:
            scores[:] = np.nan
        return float(np.mean(scores)), float(np.std(scores)), scores


class StatsPack:
    """ONE contiguous fp64 device buffer ``[count | sum x | sum x x^T | IS sums | IS counters]`` holding every additive
    statistic of a set, so that the path's single exchange step is a single ``all_reduce`` of
    33.6 MB instead of one collective per statistic."""

    def __init__(self, engine: Engine, d: int = 2048, is_classes: int = 0, is_splits: int = 0, want_moments: bool = True):
        nm = Moments.numel(d) if want_moments else 0
        ni = ISPartials.numel(is_classes, is_splits) if is_classes > 0 else 0
        self.buf = torch.zeros(max(nm + ni, 1), dtype=torch.float64, device=engine.device)
        self.moments = Moments(engine, d, self.buf[:nm]) if want_moments else None
        self.parts = ISPartials(engine, is_classes, is_splits, self.buf[nm:nm + ni]) if is_classes > 0 else None

    def allreduce(self, group=None):
        import torch.distributed as dist

        for st in (self.moments, self.parts):
            if st is not None:
                st.before_reduce()
        dist.all_reduce(self.buf, op=dist.ReduceOp.SUM, group=group)
        for st in (self.moments, self.parts):
            if st is not None:
                st.after_reduce()


def warm_collective(device, numel: int = 1 + 2048 + 2048 * 2048, group=None):
    """First use of a NCCL communicator at a given message size costs tens of milliseconds (channel set-up, buffer
    registration: 24.6 ms at 8 ranks in round 1).  Evaluation jobs call this once during set-up, next to the weight
    upload, so that the end-of-run exchange step costs what the wire costs."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return
    t = torch.zeros(numel, dtype=torch.float64, device=device)
    for _ in range(2):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    torch.cuda.synchronize(device)
