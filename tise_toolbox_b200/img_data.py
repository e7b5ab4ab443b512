"""Drop-in for image_realism/FID/img_data.py (and the O-FID copy): same ``Dataset`` class.

With ``transform=None`` the reference returns the PIL image; here ``raw_u8=True`` returns the decoded
uint8 HWC tensor instead, which is what the GPU resize kernel consumes (the CPU ``transforms.Resize``
of fid_score.py:208-213 then never runs).

``DeviceImageLoader`` replaces ``DataLoader(Dataset(path, transform), num_workers=8)`` of fid_score.py:206-217
altogether: host threads only read file BYTES; PNG inflate / unfilter, the resize and
ToTensor all run on the GPU.
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch
from PIL import Image
from torch.utils import data


class Dataset(data.Dataset):
    "Characterizes a dataset for PyTorch"

    def __init__(self, path, transform=None, raw_u8=False):
        self.file_names = self.get_filenames(path)
        self.transform = transform
        self.raw_u8 = raw_u8

    def __len__(self):
        return len(self.file_names)

    def __getitem__(self, index):
        img = Image.open(self.file_names[index]).convert("RGB")  # img_data.py:21
        if self.raw_u8:
            return torch.from_numpy(np.array(img, dtype=np.uint8))
        if self.transform is not None:
            img = self.transform(img)
        return img

    def get_filenames(self, data_path):
        # img_data.py:27-35: recursive walk, keep names *containing* "jpg" or "png", filesystem order
        images = []
        for path, subdirs, files in os.walk(data_path):
            for name in files:
                if name.rfind("jpg") != -1 or name.rfind("png") != -1:
                    filename = os.path.join(path, name)
                    if os.path.isfile(filename):
                        images.append(filename)
        return images


def collate_u8(batch):
    """uint8 HWC tensors of one size stack to [B,H,W,3]; ragged batches stay a list."""
    if all(b.shape == batch[0].shape for b in batch):
        return torch.stack(batch)
    return list(batch)


class DeviceImageLoader:
    """Iterable of batches for ``calculate_activation_statistics`` / ``get_activations``: same order, same
    ``batch_size`` / ``drop_last`` semantics as the reference's DataLoader(shuffle=False) (fid_score.py:214-217),
    but every batch is a :class:`~tise_toolbox_b200.engine.PackedImages` decoded on the device
    (``Engine.decode_images``).

    One warp inflates one image, so the decoder needs thousands of images in flight to fill the GPU: files are read
    and decoded ``decode_batch`` at a time (default 4096, whatever the consumer's ``batch_size`` is -- the
    reference's 50 / 64 would leave 97 % of the SMs idle) and handed out as ``batch_size`` slices of that pack.
    File reads of the next group run on ``io_threads`` host threads while the GPU works on the current one."""

    def __init__(self, file_names, engine, batch_size, drop_last=True, io_threads=None, decode_batch=4096):
        self.file_names = list(file_names)
        self.engine = engine
        self.batch_size = int(batch_size)
        self.drop_last = bool(drop_last)
        self.io_threads = int(io_threads) if io_threads else max(4, min(16, os.cpu_count() or 8))
        per = max(1, int(decode_batch) // self.batch_size)
        self.group = per * self.batch_size  # files per decode call: a whole number of consumer batches

    def __len__(self):
        n, b = len(self.file_names), self.batch_size
        return n // b if self.drop_last else (n + b - 1) // b

    def _group_lists(self):
        b = self.batch_size
        n_used = len(self) * b if self.drop_last else len(self.file_names)
        names = self.file_names[:n_used]
        # ramp-up: a short first group (~1024 files) lets the GPU start after ~30 ms of file reads instead of ~110 ms;
        # every later group is read while its predecessor is on the GPU
        first = min(self.group, max(b, 1024 // b * b))
        if len(names) <= 2 * first:
            first = self.group
        cuts = [0] + list(range(first, len(names), self.group)) if len(names) > first else [0]
        return [names[lo:hi] for lo, hi in zip(cuts, cuts[1:] + [len(names)]) if hi > lo]

    def prestage(self):
        """Starts reading this loader's first group in the background (idempotent).  A consumer that walks several
        loaders back to back (fid_score: generated set, then real set) calls it while the previous loader's last group
        is on the GPU, so the second set does not pay the pipeline fill again."""
        if getattr(self, "_pre", None) is None:
            self._lists = self._group_lists()
            self._readers = ThreadPoolExecutor(max_workers=self.io_threads)
            self._stager = ThreadPoolExecutor(max_workers=1)
            self._pre = self._stager.submit(self.engine.stage_files, self._lists[0], self._readers) if self._lists else None

    def groups(self, next_loader=None):
        """The decoded packs themselves, ``decode_batch`` files at a time (a whole number of consumer batches each).
        Consumers whose result does not depend on how the set is cut into batches (the FID moments, the IS partial
        sums) take these instead of the ``batch_size`` slices: 50 images per trunk call would leave the GPU idle.
        ``next_loader``: a loader whose first group is pre-staged while this one's last group is being processed."""
        self.prestage()
        groups, nxt = self._lists, self._pre
        self._pre = None
        try:
            # one staging thread reads group k + 1 into the engine's other pinned buffer (itself fanning the reads out
            # over `io_threads`) while this thread uploads / decodes group k and the consumer runs the trunk on it
            for gi, g in enumerate(groups):
                staged = nxt.result()
                # the next group goes into the engine's OTHER pinned buffer (last used by group gi - 1, whose upload was
                # awaited inside its decode_staged), so its file reads overlap this group's upload, decode and trunk
                if gi + 1 < len(groups):
                    nxt = self._stager.submit(self.engine.stage_files, groups[gi + 1], self._readers)
                elif next_loader is not None:
                    next_loader.prestage()
                pack = self.engine.decode_staged(staged)  # returns after the H2D + decode of this group completed
                yield pack
        finally:
            self._stager.shutdown(wait=True)
            self._readers.shutdown(wait=True)

    def __iter__(self):
        from .engine import PackedImages

        b = self.batch_size
        for pack in self.groups():
            for i in range(0, len(pack), b):
                yield PackedImages(pack.buf, pack.offsets[i:i + b], pack.hw[i:i + b])
