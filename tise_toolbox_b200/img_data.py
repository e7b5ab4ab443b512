"""Drop-in for image_realism/FID/img_data.py (and the O-FID copy): same ``Dataset`` class.

With ``transform=None`` the reference returns the PIL image; here ``raw_u8=True`` returns the decoded
uint8 HWC tensor instead, which is what the GPU resize kernel consumes (the CPU ``transforms.Resize``
of fid_score.py:208-213 then never runs).
"""
import os

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
