"""CPU experiment (torch fp32 emulation): single-pass layers (weights rounded to fp16, "W_hi only") with and without
the mean-field W_lo correction  shift += W_lo . mean(activations)  calibrated on a few images.

  python tools/bias_correction_probe.py [N images] [calibration images]

Prints pool3 relative L2 and |d mu| / |mu| against the fp32 network for several single-pass policies.
"""
import sys
import time

import numpy as np
import torch
from torchvision.models import Inception3

sys.path.insert(0, ".")
from oracle import harness  # noqa: E402
from tise_toolbox_b200 import synth  # noqa: E402


def build(sd):
    net = Inception3(num_classes=1000, aux_logits=True, transform_input=False, init_weights=False)
    net.load_state_dict(sd)
    net.eval()
    return net


def convs_with_input_hw(net, x):
    """{conv module: input height} from one forward pass."""
    hw = {}
    hooks = [m.register_forward_pre_hook(lambda mod, inp: hw.__setitem__(mod, int(inp[0].shape[2])))
             for m in net.modules() if isinstance(m, torch.nn.Conv2d)]
    with torch.no_grad():
        net(x)
    for h in hooks:
        h.remove()
    return hw


@torch.no_grad()
def pool3(net, x, bs=50):
    out = []
    for i in range(0, x.shape[0], bs):
        y = x[i:i + bs]
        n = net
        y = n.Conv2d_1a_3x3(y); y = n.Conv2d_2a_3x3(y); y = n.Conv2d_2b_3x3(y); y = n.maxpool1(y)
        y = n.Conv2d_3b_1x1(y); y = n.Conv2d_4a_3x3(y); y = n.maxpool2(y)
        for m in (n.Mixed_5b, n.Mixed_5c, n.Mixed_5d, n.Mixed_6a, n.Mixed_6b, n.Mixed_6c, n.Mixed_6d, n.Mixed_6e,
                  n.Mixed_7a, n.Mixed_7b, n.Mixed_7c):
            y = m(y)
        out.append(y.mean(dim=(2, 3)))
    return torch.cat(out).double().numpy()


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n = int(args[0]) if len(args) > 0 else 200
    ncal = int(args[1]) if len(args) > 1 else 64
    torch.set_num_threads(8)
    sd = synth.inception_state_dict(0, num_classes=1000)
    imgs = torch.cat([synth.images(1, 0, 0, n // 2), synth.images(2, 1, 0, n // 2)])
    x = harness.pil_preprocess([i.numpy() for i in imgs], "fid")
    cal_sets = {"same-dist": torch.cat([synth.images(1, 0, 5000, ncal // 2), synth.images(2, 1, 5000, ncal // 2)]),
                "other-dist": synth.images(7, 3, 9000, ncal),
                "white-noise": torch.randint(0, 256, (ncal, 256, 256, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(1)),
                "constant-grey": torch.full((ncal, 256, 256, 3), 128, dtype=torch.uint8)}
    ref_net = build(sd)
    hw = convs_with_input_hw(ref_net, x[:2])
    t0 = time.time()
    ref = pool3(ref_net, x)
    print("fp32 reference: %.1f s" % (time.time() - t0), flush=True)

    def run(max_hw, cal=None, act16=True):
        net = build(sd)
        name_of = {m: k for k, m in net.named_modules()}
        hw2 = convs_with_input_hw(net, x[:2])
        single = [m for m, h in hw2.items() if h <= max_hw and name_of[m] != "Conv2d_1a_3x3.conv"]
        lo = {}
        for m in single:
            w = m.weight.data
            amax = w.abs().flatten(1).max(1).values.clamp_min(1e-30)
            sc = torch.pow(2.0, 11 - torch.floor(torch.log2(amax))).view(-1, 1, 1, 1)
            hi = (w * sc).half().float() / sc
            lo[m] = w - hi
            m.weight.data = hi
        hooks = []
        if act16:  # fp16 activations at every conv input, as on the device
            for m in net.modules():
                if isinstance(m, torch.nn.Conv2d) and name_of[m] != "Conv2d_1a_3x3.conv":
                    hooks.append(m.register_forward_pre_hook(lambda mod, inp: (inp[0].half().float(),)))
        if cal is not None:
            means = {}
            ch = [m.register_forward_pre_hook(lambda mod, inp: means.__setitem__(mod, inp[0].mean(dim=(0, 2, 3))))
                  for m in single]
            xc = harness.pil_preprocess([i.numpy() for i in cal], "fid")
            # calibration pass with the exact weights (what a strict forward would see)
            for m in single:
                m.weight.data = m.weight.data + lo[m]
            with torch.no_grad():
                net(xc)
            for m in single:
                m.weight.data = m.weight.data - lo[m]
            for h in ch:
                h.remove()
            for m in single:
                corr = (lo[m].sum(dim=(2, 3)) * means[m][None, :]).sum(1)  # all taps valid (borders ignored)
                hooks.append(m.register_forward_hook(lambda mod, inp, out, c=corr: out + c.view(1, -1, 1, 1)))
        f = pool3(net, x)
        rel = float(np.linalg.norm(f - ref) / np.linalg.norm(ref))
        mu = float(np.linalg.norm(f.mean(0) - ref.mean(0)) / np.linalg.norm(ref.mean(0)))
        return rel, mu, len(single)

    print("%-34s %10s %10s" % ("policy", "pool3 L2", "|dmu|/|mu|"))
    r = run(0)
    print("%-34s %10.2e %10.2e  (%d single-pass convs)" % ("strict (fp16 activations only)", *r), flush=True)
    for max_hw in ((1000,) if "--robust" in sys.argv else (17, 35, 1000)):
        r = run(max_hw)
        print("%-34s %10.2e %10.2e  (%d)" % ("1-pass hw<=%d" % max_hw, *r), flush=True)
        for cname, cal in cal_sets.items():
            r = run(max_hw, cal)
            print("%-34s %10.2e %10.2e" % ("  + correction (%s)" % cname, r[0], r[1]), flush=True)


if __name__ == "__main__":
    main()
