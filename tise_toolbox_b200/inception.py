"""Drop-in for image_realism/FID/inception.py and object_fidelity/O-FID/inception.py: ``InceptionV3``
with the reference's constructor arguments and ``forward`` contract, running on libtise_b200.

Differences a caller can observe (both forced by the environment, neither changes FID):
  * weights: the reference downloads torchvision's ImageNet checkpoint (inception.py:57) or loads
    ``weights/inceptionv3_fine_to_with_80_coco_classes.pth`` (O-FID/inception.py:58-64).  Here the
    torchvision ``Inception3`` state_dict comes from ``state_dict=`` / ``weights_path=`` or the
    ``TISE_INCEPTION_WEIGHTS`` environment variable; nothing is downloaded.
  * blocks 0-2 return the spatially averaged map [B, C, 1, 1] instead of [B, C, H, W]; the reference's
    only consumer applies ``adaptive_avg_pool2d`` to it straight away (fid_score.py:108-111).
"""
import os

import torch
import torch.nn as nn

from .engine import Engine, PackedImages

_DIMS = {0: 64, 1: 192, 2: 768, 3: 2048}


def load_state_dict_file(path):
    sd = torch.load(path, map_location="cpu")
    if isinstance(sd, dict) and "state_dict" in sd and not any(k.endswith("conv.weight") for k in sd):
        sd = sd["state_dict"]
    return sd


class InceptionV3(nn.Module):
    """Inception-v3 feature extractor (same interface as the reference class)."""

    DEFAULT_BLOCK_INDEX = 3
    BLOCK_INDEX_BY_DIM = {64: 0, 192: 1, 768: 2, 2048: 3}

    def __init__(self, output_blocks=[DEFAULT_BLOCK_INDEX], resize_input=True, normalize_input=True,
                 requires_grad=False, state_dict=None, weights_path=None, device=None, micro_batch=1024, engine=None,
                 single_pass_max_hw=None):
        super(InceptionV3, self).__init__()
        self.resize_input = resize_input
        self.normalize_input = normalize_input
        self.output_blocks = sorted(output_blocks)
        self.last_needed_block = max(output_blocks)
        assert self.last_needed_block <= 3, "Last possible output block index is 3"
        if requires_grad:
            raise NotImplementedError("libtise_b200 is an inference path: requires_grad=True is not supported")
        if engine is not None:  # share an existing context (weights already on the device)
            self.engine = engine
            return
        if state_dict is None:
            weights_path = weights_path or os.environ.get("TISE_INCEPTION_WEIGHTS")
            if not weights_path:
                raise RuntimeError("InceptionV3: pass state_dict= / weights_path= or set TISE_INCEPTION_WEIGHTS to a "
                                   "torchvision Inception3 state_dict (no download is attempted)")
            state_dict = load_state_dict_file(weights_path)
        self.engine = Engine(state_dict, device=device, variant="fid", micro_batch=micro_batch,
                             single_pass_max_hw=single_pass_max_hw)

    # the reference calls model.cuda() / model.eval(); both are no-ops here (always CUDA, always eval)
    def cuda(self, device=None):
        return self

    def eval(self):
        return self

    def forward(self, inp):
        """inp: float Bx3xHxW in (0, 1) (inception.py:100-134).  Returns one [B, C, 1, 1] tensor per
        requested block."""
        x = inp.to(self.engine.device, torch.float32)
        if self.resize_input and tuple(x.shape[2:]) != (299, 299):
            # identity at 299x299 (the Dataset already resized with Pillow); live only for odd callers
            x = self.engine.upsample_float(x, align_corners=True)
        if self.normalize_input:
            # inception.py:119-124, torchvision's transform_input applied to [0,1] data: one libtise kernel with the
            # reference's rounding (multiply, then add)
            x = self.engine.channel_affine(x, (0.229 / 0.5, 0.224 / 0.5, 0.225 / 0.5),
                                           ((0.485 - 0.5) / 0.5, (0.456 - 0.5) / 0.5, (0.406 - 0.5) / 0.5))
        outp = []
        for idx in self.output_blocks:
            f = self.engine.features_from_float(x, dims=_DIMS[idx])
            outp.append(f.view(f.shape[0], f.shape[1], 1, 1))
        return outp

    def features_u8(self, images, dims=2048):
        """Fast path: uint8 HWC images of any size (tensor [B,H,W,3] or list of [H,W,3]) -> [B, dims]
        float32; Pillow-exact resize, ToTensor and the affine above all happen on the GPU."""
        self.engine.set_variant("fid")
        if isinstance(images, PackedImages):
            pass  # already on the device (Engine.decode_images)
        elif torch.is_tensor(images):
            images = images.to(self.engine.device, non_blocking=True)
        else:
            images = [t.to(self.engine.device, non_blocking=True) for t in images]
        return self.engine.features(images, dims=dims)
