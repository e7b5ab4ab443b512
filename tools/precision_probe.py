"""CPU experiment: how much FID moves when the conv stack's operands are rounded the way the
tensor-core path rounds them.  Decides the operand format of the tcgen05 kernels.

  python tools/precision_probe.py [N]   (N images per set, default 500)
"""
import sys
import time

import numpy as np
import torch
from scipy import linalg
from torchvision.models import Inception3

sys.path.insert(0, ".")
from tise_toolbox_b200 import synth  # noqa: E402


def build(sd):
    net = Inception3(num_classes=1000, aux_logits=True, transform_input=False, init_weights=False)
    net.load_state_dict(sd)
    net.eval()
    return net


def features(net, imgs_u8, bs=50):
    feats = []
    with torch.no_grad():
        for i in range(0, imgs_u8.shape[0], bs):
            x = imgs_u8[i:i + bs].permute(0, 3, 1, 2).float().div(255)
            x = torch.nn.functional.interpolate(x, size=(299, 299), mode="bilinear", align_corners=False)
            x = x.clone()
            x[:, 0] = x[:, 0] * (0.229 / 0.5) + (0.485 - 0.5) / 0.5
            x[:, 1] = x[:, 1] * (0.224 / 0.5) + (0.456 - 0.5) / 0.5
            x[:, 2] = x[:, 2] * (0.225 / 0.5) + (0.406 - 0.5) / 0.5
            f = net._forward(x)[0] if False else None
            # manual trunk: everything up to avgpool
            y = net.Conv2d_1a_3x3(x); y = net.Conv2d_2a_3x3(y); y = net.Conv2d_2b_3x3(y); y = net.maxpool1(y)
            y = net.Conv2d_3b_1x1(y); y = net.Conv2d_4a_3x3(y); y = net.maxpool2(y)
            for m in (net.Mixed_5b, net.Mixed_5c, net.Mixed_5d, net.Mixed_6a, net.Mixed_6b, net.Mixed_6c,
                      net.Mixed_6d, net.Mixed_6e, net.Mixed_7a, net.Mixed_7b, net.Mixed_7c):
                y = m(y)
            feats.append(y.mean(dim=(2, 3)).double().numpy())
    return np.concatenate(feats)


def fid(a, b):
    m1, m2 = a.mean(0), b.mean(0)
    s1, s2 = np.cov(a, rowvar=False), np.cov(b, rowvar=False)
    cm = linalg.sqrtm(s1.dot(s2))
    cm = cm.real
    return float((m1 - m2).dot(m1 - m2) + np.trace(s1) + np.trace(s2) - 2 * np.trace(cm))


def install(net, mode):
    """mode: 'f32' | 'f16a' (fp16 activations, hi+lo fp16 weights) | 'bf16x3' | 'f16' | 'bf16'."""
    first = net.Conv2d_1a_3x3.conv
    stem = {id(net.Conv2d_2a_3x3.conv), id(net.Conv2d_2b_3x3.conv), id(net.Conv2d_3b_1x1.conv), id(net.Conv2d_4a_3x3.conv)}
    if mode == "f16in":  # the product path with the stem conv on the tensor cores too: its input rounded to fp16 as well
        w = first.weight.data
        hi = w.half().float(); lo = (w - hi).half().float(); first.weight.data = hi + lo
        first.register_forward_pre_hook(lambda m, inp: (inp[0].half().float(),))
        mode = "f16a"
    if mode in ("f16a_lo8", "f16a_lo8e4"):
        # candidate for the next round: W_hi pass in fp16 as now, W_lo pass as kind::f8f6f4 -- activations as e5m2 (the
        # top byte of the fp16 value) or e4m3, W_lo (|lo| <= 1 after the per-channel power-of-two scaling) as e4m3
        act8 = torch.float8_e5m2 if mode == "f16a_lo8" else torch.float8_e4m3fn
        for mod in net.modules():
            if isinstance(mod, torch.nn.Conv2d) and mod is not first:
                w = mod.weight.data
                amax = w.abs().flatten(1).max(1).values.clamp_min(1e-30)
                e = torch.floor(torch.log2(4095.0 / amax))
                sc = torch.pow(2.0, e).view(-1, 1, 1, 1)
                ws = w * sc
                hi = ws.half().float()
                lo8 = (ws - hi).to(torch.float8_e4m3fn).float()
                mod.w_hi = hi / sc
                mod.w_lo = lo8 / sc

                def fwd(x, m=mod):
                    a16 = x.half().float()
                    a8 = a16.to(act8).float()
                    return (torch.nn.functional.conv2d(a16, m.w_hi, None, m.stride, m.padding) +
                            torch.nn.functional.conv2d(a8, m.w_lo, None, m.stride, m.padding))
                mod.forward = fwd
        return
    for mod in net.modules():
        if isinstance(mod, torch.nn.Conv2d) and mod is not first:
            w = mod.weight.data
            if mode in ("f16stem", "f16nostem"):  # single fp16 pass only in (f16stem) / except in (f16nostem) the stem convs
                single = (id(mod) in stem) == (mode == "f16stem")
                if single:
                    mod.weight.data = w.half().float()
                else:
                    hi = w.half().float(); lo = (w - hi).half().float(); mod.weight.data = hi + lo
                mod.register_forward_pre_hook(lambda m, inp: (inp[0].half().float(),))
            elif mode == "f16a":
                hi = w.half().float(); lo = (w - hi).half().float(); mod.weight.data = hi + lo
                mod.register_forward_pre_hook(lambda m, inp: (inp[0].half().float(),))
            elif mode == "f16":
                mod.weight.data = w.half().float()
                mod.register_forward_pre_hook(lambda m, inp: (inp[0].half().float(),))
            elif mode == "bf16":
                mod.weight.data = w.bfloat16().float()
                mod.register_forward_pre_hook(lambda m, inp: (inp[0].bfloat16().float(),))
            elif mode == "bf16x3":
                hi = w.bfloat16().float(); lo = (w - hi).bfloat16().float(); mod.weight.data = hi + lo

                def hook(m, inp):
                    x = inp[0]; h = x.bfloat16().float(); l = (x - h).bfloat16().float()
                    return (h + l,)
                mod.register_forward_pre_hook(hook)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
    modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["f32", "f16a", "f16"]
    sd = synth.inception_state_dict(0)
    a = synth.images(1, 0, 0, n)
    b = synth.images(2, 1, 0, n)
    ref = None
    for mode in modes:
        net = build(sd)
        install(net, mode)
        t = time.time()
        fa, fb = features(net, a), features(net, b)
        v = fid(fa, fb)
        if mode == "f32":
            ref = (fa, fb, v)
            print(f"pool3 mean {fa.mean():.4f} std-across-images {fa.std(0).mean():.4f} max {fa.max():.2f}")
        rel = abs(v - ref[2]) / ref[2] if ref else float("nan")
        ferr = np.linalg.norm(fa - ref[0]) / np.linalg.norm(ref[0]) if ref else float("nan")
        print(f"{mode:8s} FID {v:.6f} rel_err {rel:.3e} pool3 relL2 {ferr:.3e}  ({time.time() - t:.0f}s)", flush=True)


if __name__ == "__main__":
    main()
