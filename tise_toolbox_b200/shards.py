"""Raw uint8 shard files: the zero-decode ingest format of SURVEY.md §8f row 1 ("PNG/JPEG decode ... or raw-u8 shard
format").  A generator (e.g. AttnGAN++/coco_gen_image_from_caption.py:148-156, which saves one PNG per image) or a
one-off conversion of a folder writes its images as packed HWC RGB bytes; evaluation then moves them file -> pinned
memory -> HBM without touching a codec, in exactly the :class:`~tise_toolbox_b200.engine.PackedImages` layout the resize
kernel reads.

File layout (little endian):
    0   8 bytes   magic  b"TISEU8\\x00\\x01"
    8   u64       n      number of images
    16  n x (u64 offset, u32 height, u32 width)      offset of the pixels from the start of the file, 16-byte aligned
    ..  pixels    height x width x 3 uint8 per image, each image padded to a multiple of 16 bytes

    python -m tise_toolbox_b200.shards pack  <image folder> <out.shard>     (same file walk and order as img_data.Dataset)
    python -m tise_toolbox_b200.shards info  <file.shard>
"""
import os
import struct
import sys

import numpy as np

MAGIC = b"TISEU8\x00\x01"
_ENTRY = struct.Struct("<QII")


def _align16(x):
    return (x + 15) // 16 * 16


def write_shard(path, images):
    """``images``: sequence of HWC uint8 RGB arrays (any sizes).  Returns the number of bytes written."""
    n = len(images)
    header = 16 + n * _ENTRY.size
    off = _align16(header)
    table = []
    for im in images:
        a = np.asarray(im)
        if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
            raise ValueError("write_shard: images must be HWC uint8 RGB, got %s %s" % (a.dtype, a.shape))
        table.append((off, a.shape[0], a.shape[1]))
        off = _align16(off + a.size)
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<Q", n))
        for e in table:
            f.write(_ENTRY.pack(*e))
        f.write(b"\x00" * (_align16(header) - header))
        for im, (o, h, w) in zip(images, table):
            a = np.ascontiguousarray(im)
            f.write(a.tobytes())
            f.write(b"\x00" * (_align16(a.size) - a.size))
    return off


def read_index(path):
    """(offsets int64 [n], hw int32 [n, 2], file size) of a shard; validates the header and the table."""
    size = os.path.getsize(path)
    with open(path, "rb") as f:
        head = f.read(16)
        if len(head) < 16 or head[:8] != MAGIC:
            raise ValueError("%s is not a TISE uint8 shard" % path)
        n = struct.unpack("<Q", head[8:])[0]
        if 16 + n * _ENTRY.size > size:
            raise ValueError("%s: truncated index (%d images)" % (path, n))
        raw = f.read(n * _ENTRY.size)
    tab = np.frombuffer(raw, dtype=np.dtype([("off", "<u8"), ("h", "<u4"), ("w", "<u4")]))
    offs = tab["off"].astype(np.int64)
    hw = np.stack([tab["h"], tab["w"]], 1).astype(np.int32) if n else np.zeros((0, 2), np.int32)
    ends = offs + hw[:, 0].astype(np.int64) * hw[:, 1] * 3
    if n and (ends.max() > size or (offs % 16).any() or (hw <= 0).any() or offs.min() < 16 + n * _ENTRY.size):
        raise ValueError("%s: corrupt index" % path)
    # ShardLoader uploads [offs[g0], ends[g1 - 1]) and addresses the images relative to offs[g0]: entries must be in
    # file order and must not overlap, or the device would read outside the uploaded span
    if n > 1 and (ends[:-1] > offs[1:]).any():
        raise ValueError("%s: corrupt index (entries out of order or overlapping)" % path)
    return offs, hw, size


def read_images(path, indices=None):
    """Host-side reader (tests, tools): list of HWC uint8 arrays."""
    offs, hw, _ = read_index(path)
    idx = range(len(offs)) if indices is None else indices
    out = []
    with open(path, "rb") as f:
        for i in idx:
            f.seek(int(offs[i]))
            h, w = int(hw[i, 0]), int(hw[i, 1])
            out.append(np.frombuffer(f.read(h * w * 3), dtype=np.uint8).reshape(h, w, 3))
    return out


class ShardLoader:
    """Batches of a shard as :class:`PackedImages` on the device, ``batch_size`` / ``drop_last`` as in the reference's
    DataLoader (fid_score.py:214-217).  ``groups()`` yields ``group`` images per H2D copy for the batch-invariant
    consumers, exactly like :class:`~tise_toolbox_b200.img_data.DeviceImageLoader`."""

    def __init__(self, path, engine, batch_size, drop_last=True, group=4096, rank=0, world=1):
        self.path, self.engine = path, engine
        self.batch_size, self.drop_last = int(batch_size), bool(drop_last)
        self.offsets, self.hw, self.size = read_index(path)
        if world > 1:  # one process per GPU: a contiguous run of whole batches each, like the folder branch
            from .pipeline import shard_batches

            lo, hi = shard_batches(len(self.offsets), self.batch_size, rank, world)
            self.offsets, self.hw = self.offsets[lo:hi], self.hw[lo:hi]
        self.group = max(1, int(group) // self.batch_size) * self.batch_size

    def __len__(self):
        n, b = len(self.offsets), self.batch_size
        return n // b if self.drop_last else (n + b - 1) // b

    def groups(self):
        import torch

        from .engine import PackedImages

        b = self.batch_size
        n_used = len(self) * b if self.drop_last else len(self.offsets)
        with open(self.path, "rb", buffering=0) as f:
            for g0 in range(0, n_used, self.group):
                g1 = min(n_used, g0 + self.group)
                lo = int(self.offsets[g0])
                hi = int(self.offsets[g1 - 1] + int(self.hw[g1 - 1, 0]) * int(self.hw[g1 - 1, 1]) * 3)
                pinned = self.engine._pinned(hi - lo)
                view = memoryview(pinned.numpy())[:hi - lo]
                f.seek(lo)
                got = 0
                while got < hi - lo:
                    r = f.readinto(view[got:])
                    if not r:
                        raise IOError("short read of %s" % self.path)
                    got += r
                dev = pinned[:hi - lo].to(self.engine.device, non_blocking=True)
                torch.cuda.current_stream(self.engine.device).synchronize()  # the pinned buffer is reused by the next group
                yield PackedImages(dev, self.offsets[g0:g1] - lo, self.hw[g0:g1])

    def __iter__(self):
        from .engine import PackedImages

        b = self.batch_size
        for pack in self.groups():
            for i in range(0, len(pack), b):
                yield PackedImages(pack.buf, pack.offsets[i:i + b], pack.hw[i:i + b])


def pack_folder(folder, out):
    """Folder of image files -> shard, same walk / order / convert("RGB") as img_data.Dataset (decoded with PIL on the
    host: a one-off conversion)."""
    from PIL import Image

    from .img_data import Dataset

    names = Dataset(folder).file_names
    images = [np.asarray(Image.open(n).convert("RGB")) for n in names]
    return write_shard(out, images), len(images)


if __name__ == "__main__":
    if len(sys.argv) == 4 and sys.argv[1] == "pack":
        nbytes, n = pack_folder(sys.argv[2], sys.argv[3])
        print("%s: %d images, %d bytes" % (sys.argv[3], n, nbytes))
    elif len(sys.argv) == 3 and sys.argv[1] == "info":
        offs, hw, size = read_index(sys.argv[2])
        print("%s: %d images, %d bytes, sizes %s .. %s" % (sys.argv[2], len(offs), size, hw.min(0).tolist() if len(hw) else "-",
                                                              hw.max(0).tolist() if len(hw) else "-"))
    else:
        print(__doc__)
        sys.exit(2)
