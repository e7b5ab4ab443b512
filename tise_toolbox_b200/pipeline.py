"""The fused, sharded form of the hot path: uint8 images in HBM -> resize -> Inception-v3 -> FID moments
and IS partial sums, one process per GPU, a single end-of-run all-reduce.

``RealismJob`` is what bench.py times and what a multi-GPU evaluation script would drive; the
reference-facing modules (fid_score, ...) give the same numbers through the reference's call surface.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import inception_score as _is
from .engine import Engine, ISPartials, Moments, StatsPack


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous global index range [lo, hi) of ``rank``: rank r gets [r n / R, (r + 1) n / R)."""
    return (rank * n) // world, ((rank + 1) * n) // world


def shard_batches(n_files: int, batch_size: int, rank: int, world: int) -> Tuple[int, int]:
    """File index range [lo, hi) of ``rank`` when a folder is evaluated by ``world`` processes: the reference's
    ``drop_last=True`` (fid_score.py:216) is applied to the whole set first, then whole batches are dealt out in
    contiguous runs, so every rank sees full batches only and the union is exactly what one process would use."""
    lo, hi = shard_range(n_files // batch_size if batch_size > 0 else 0, rank, world)
    return lo * batch_size, hi * batch_size


def format_result(kind: str, a: float, b: Optional[float] = None) -> str:
    """The one-line result files of the reference, byte for byte."""
    if kind == "fid":
        return f"FID: {a}"  # image_realism/FID/fid_score.py:251-252
    if kind == "ofid":
        return f"O-FID: {a}"  # object_fidelity/O-FID/fid_score.py:219-222
    if kind == "coco":
        return "[Inception Score] mean: {:.5f} std: {:.5f}".format(a, b)  # inception_score_star_coco.py:153
    if kind == "bird":
        return f"IS = {a}  +-  {b}"  # inception_score_star_bird.py:209
    if kind == "ois":
        return f"O-IS: {a} +-  {b}"  # object_centric_inception_score.py:127
    raise ValueError(kind)


class RealismJob:
    """Streams batches of uint8 images through the trunk and keeps the additive statistics.

    n_total / start: global number of images of the set and the global index of this rank's first
    image -- they fix the split of every image independently of the number of ranks.
    is_variant: None (FID only) or "ois" / "coco" / "bird"."""

    def __init__(self, engine: Engine, n_total: int, start: int = 0, is_variant: Optional[str] = None,
                 splits: int = 10, input_variant: str = "fid", want_moments: bool = True):
        self.e = engine
        self.n_total, self.pos = int(n_total), int(start)
        self.is_variant = is_variant
        self.input_variant = input_variant
        cls = 0
        if is_variant is not None:
            _, off, _, _ = _is.VARIANTS[is_variant]
            cls = engine.num_classes - off
        # every additive statistic of the set lives in ONE packed fp64 buffer: the exchange step is one all-reduce
        import os

        self.pace_h2d = os.environ.get("TISE_PACE_H2D", "1") != "0"  # step_from_host: see there
        self.pack = StatsPack(engine, 2048, cls, splits if cls else 0, want_moments)
        self.moments, self.parts = self.pack.moments, self.pack.parts

    def step(self, images: Union[torch.Tensor, Sequence[torch.Tensor]]) -> torch.Tensor:
        """One batch (device uint8 [B,H,W,3] or list of [H,W,3]); returns the batch's pool3 features."""
        e = self.e
        e.set_variant(self.input_variant)
        feat = e.features(images)
        if self.moments is not None:
            self.moments.update(feat)
        if self.parts is not None:
            _, _, use_bias, _ = _is.VARIANTS[self.is_variant]
            logits = e.logits(feat, use_bias=use_bias)
            _is.update_from_logits(e, self.parts, logits, self.is_variant, self.n_total, self.pos)
        self.pos += int(feat.shape[0])
        return feat

    def step_from_host(self, pinned: torch.Tensor, next_pinned: Optional[torch.Tensor] = None) -> torch.Tensor:
        """End-to-end step from pinned host memory.  The H2D copy runs on a private copy stream into one of two
        staging buffers; passing the following batch as ``next_pinned`` lets its copy overlap this step's compute (paced:
        one chunk per micro-batch, released when the trunk leaves its HBM-bound front).  The caller keeps both host
        batches alive and unmodified until the step that consumes them has been issued."""
        dev = self.e.device
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._staging = [None, None]
            self._slot_done = [None, None]  # event after the last step that read each staging buffer
            self._pending = {}
            self._slot = 0
            self._progress = None

        def staging(slot, like):
            """Device buffer of slot `slot` shaped like the host batch `like`: grow-only storage, a view per batch.
            Growing allocates from the COMPUTE stream's pool, whose free blocks may still be read by kernels queued
            there under their previous identity -- the copy stream, which writes the buffer, must wait for them
            (without this a batch of a new size could be overwritten in flight: caught by
            tests/test_gpu_e2e.py::test_step_from_host_equals_resident_steps)."""
            need = like.numel()
            cur = self._staging[slot]
            if cur is None or cur.numel() < need:
                cur = self._staging[slot] = torch.empty(need, dtype=torch.uint8, device=dev)
                self._copy_stream.wait_stream(torch.cuda.current_stream(dev))
            return cur[:need].view(like.shape)

        def start_copy(src, slot):
            buf = staging(slot, src)
            if self._slot_done[slot] is not None:  # only the last reader of THIS buffer has to be finished
                self._copy_stream.wait_event(self._slot_done[slot])
            with torch.cuda.stream(self._copy_stream):
                buf.copy_(src, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            self._pending[src.data_ptr()] = (slot, ev, buf)

        key = pinned.data_ptr()
        if key not in self._pending:
            start_copy(pinned, self._slot)
        slot, ev, cur_buf = self._pending.pop(key)
        torch.cuda.current_stream(dev).wait_event(ev)
        self._slot = 1 - slot
        paced = self.pace_h2d and next_pinned is not None and next_pinned.data_ptr() not in self._pending
        if paced:
            # The next batch crosses PCIe in one chunk per micro-batch of THIS step, each released by the trunk's progress
            # event (tise_set_progress_event: the forward pass has left the HBM-bound front of the network), so the DMA
            # traffic runs beside the compute-bound layers instead of beside the ones at the memory limit
            # (tools/e2e_probe.py: an H2D transfer beside the trunk costs 3-5 % wherever it lands by default).
            nslot = self._slot
            buf = staging(nslot, next_pinned)
            if self._slot_done[nslot] is not None:
                self._copy_stream.wait_event(self._slot_done[nslot])
            self._progress = self.e.progress_event()
            total = int(next_pinned.shape[0])
            state = {"done": 0}

            def hook(k, count):
                hi = total if k == count - 1 else (total * (k + 1)) // count
                lo = state["done"]
                if hi > lo:
                    self._copy_stream.wait_event(self._progress)  # the record of micro-batch k has just been issued
                    with torch.cuda.stream(self._copy_stream):
                        buf[lo:hi].copy_(next_pinned[lo:hi], non_blocking=True)
                    state["done"] = hi

            self.e.on_micro_batch = hook
        try:
            out = self.step(cur_buf)
        finally:
            self.e.on_micro_batch = None
        self._slot_done[slot] = torch.cuda.Event()
        self._slot_done[slot].record(torch.cuda.current_stream(dev))
        if paced:
            if state["done"] < total:  # (a step without micro-batches: nothing was released)
                with torch.cuda.stream(self._copy_stream):
                    buf[state["done"]:].copy_(next_pinned[state["done"]:], non_blocking=True)
            with torch.cuda.stream(self._copy_stream):
                ev2 = torch.cuda.Event()
                ev2.record(self._copy_stream)
            self._pending[next_pinned.data_ptr()] = (nslot, ev2, buf)
        elif next_pinned is not None and next_pinned.data_ptr() not in self._pending:
            start_copy(next_pinned, self._slot)
        return out

    def read_partials_async(self):
        """Device -> host read of the running IS partial sums as they stand after the steps issued so far, without
        stalling the GPU: the sums are snapshotted on the compute stream (a 4 KB device copy), the snapshot crosses PCIe
        on a side stream into pinned memory, and the returned handle's ``wait()`` blocks the host only -- call it after
        the NEXT step has been issued and the GPU never idles between steps."""
        assert self.parts is not None
        dev = self.e.device
        if not hasattr(self, "_rb_stream"):
            self._rb_stream = torch.cuda.Stream(device=dev)
            self._rb_slots = [(torch.empty_like(self.parts.f), torch.empty(self.parts.f.numel(), dtype=torch.float64).pin_memory())
                              for _ in range(2)]
            self._rb_next = 0
        snap, host = self._rb_slots[self._rb_next]
        self._rb_next = 1 - self._rb_next
        snap.copy_(self.parts.f)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(dev))
        self._rb_stream.wait_event(ready)
        with torch.cuda.stream(self._rb_stream):
            host.copy_(snap, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self._rb_stream)

        class _Handle:
            def wait(_self):
                done.synchronize()
                return host

        return _Handle()

    def finish(self, group=None) -> Dict[str, object]:
        """All-reduce (if ``group`` / torch.distributed is initialised) and finalise on every rank."""
        import torch.distributed as dist

        use_dist = dist.is_available() and dist.is_initialized()
        out: Dict[str, object] = {}
        if use_dist:
            self.pack.allreduce(group)
        if self.moments is not None:
            out["count"] = self.moments.count
            out["mu"], out["sigma"] = self.moments.finalize_device()
        if self.parts is not None:
            _, _, _, nan_on_zero = _is.VARIANTS[self.is_variant]
            m, s, scores = self.parts.finalize(nan_on_zero=nan_on_zero)
            out["is_mean"], out["is_std"], out["is_scores"] = m, s, scores
        return out


def fid_from_stats(engine: Engine, a: Dict[str, object], b: Dict[str, object], eps: float = 1e-6) -> float:
    value, flags = engine.frechet(a["mu"], a["sigma"], b["mu"], b["sigma"], eps)
    if flags & 2:
        raise ValueError("Imaginary component: sqrt(sigma1 sigma2) has eigenvalues with |Im| > 1e-3")
    return value
