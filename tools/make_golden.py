"""Generates tests/golden/* by running the reference's own code (oracle/reference_shim.py) in the build
container, and checks oracle/restate.py against it on the way.  Run once; commit the outputs.

    python tools/make_golden.py
"""
import hashlib
import json
import os
import sys
import textwrap

import numpy as np
import torch
from PIL import Image

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import reference_shim, restate  # noqa: E402
from tise_toolbox_b200 import synth  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    assert reference_shim.available(), "needs /root/reference"
    os.makedirs(GOLD, exist_ok=True)
    meta = {"generated_with": {"torch": torch.__version__, "numpy": np.__version__}}

    # ---- A2: real Pillow vs restatement, on synthetic images of assorted sizes
    resize = []
    for i, (h, w) in enumerate([(256, 256), (64, 48), (500, 375), (299, 299), (17, 333), (480, 16), (299, 100), (33, 700)]):
        im = synth.images(11, i % 2, i, 1, h, w)[0].numpy()
        pil = np.asarray(Image.fromarray(im).resize((299, 299), Image.BILINEAR))
        assert np.array_equal(pil, restate.pillow_resize_bilinear(im)), (h, w)
        resize.append({"seed": 11, "set": i % 2, "idx": i, "h": h, "w": w, "in_sha256": sha(im), "out_sha256": sha(pil)})
    meta["resize"] = resize

    # ---- weights fingerprint
    sd = synth.inception_state_dict(0, num_classes=1000)
    meta["weights"] = {k: sha(sd[k].numpy()) for k in ("Conv2d_1a_3x3.conv.weight", "Mixed_7c.branch_pool.bn.running_var", "fc.weight")}

    # ---- A3-A9: the reference's fid_score module on 2 x 16 synthetic images, batch 8
    fid = reference_shim.load_fid_module(sd)
    model = fid.InceptionV3([3])
    sets = [synth.images(1, 0, 0, 16), synth.images(2, 1, 0, 16)]

    class Batches(list):  # a DataLoader stand-in: iterable of [8,3,299,299] float batches with __len__
        pass

    acts, stats = [], []
    for s in sets:
        # Dataset.__getitem__ of the reference: PIL -> Resize((299,299)) -> ToTensor (fid_score.py:208-213)
        t = torch.stack([torch.from_numpy(np.asarray(Image.fromarray(im.numpy()).resize((299, 299), Image.BILINEAR)))
                         .permute(2, 0, 1).float().div(255) for im in s])
        batches = Batches([t[0:8], t[8:16]])
        act = fid.get_activations(batches, model, batch_size=8, dims=2048, cuda=False, verbose=False)
        acts.append(act)
        stats.append((np.mean(act, axis=0), np.cov(act, rowvar=False)))
    fid_ref = float(fid.calculate_frechet_distance(stats[0][0], stats[0][1], stats[1][0], stats[1][1]))
    np.save(os.path.join(GOLD, "pool3_ref.npy"), np.concatenate([acts[0][:4], acts[1][:4]]).astype(np.float32))
    meta["fid_2x16"] = {"value": fid_ref, "batch_size": 8, "n_per_set": 16}

    # restatement on the same inputs
    trunk = restate.Trunk(sd)
    racts = []
    for s in sets:
        x = restate.preprocess([im.numpy() for im in s], "fid")
        racts.append(restate.get_activations([x[0:8], x[8:16]], trunk, 8))
    for a, b in zip(acts, racts):
        assert np.abs(a - b).max() <= 1e-5 * np.abs(a).max(), np.abs(a - b).max()
    rs = [restate.calculate_activation_statistics(a) for a in racts]
    fid_res = float(restate.calculate_frechet_distance(rs[0][0], rs[0][1], rs[1][0], rs[1][1]))
    assert abs(fid_res - fid_ref) <= 1e-6 * abs(fid_ref), (fid_res, fid_ref)
    print("FID 2x16: reference %.9f restatement %.9f" % (fid_ref, fid_res))

    # ---- A9 on well-conditioned seeded matrices (d = 192), reference function
    rng = np.random.RandomState(5)
    d = 192
    xa, xb = rng.randn(400, d) * rng.rand(d) + rng.randn(d), rng.randn(500, d) * rng.rand(d) * 1.3 + rng.randn(d)
    ma, sa, mb, sb = xa.mean(0), np.cov(xa, rowvar=False), xb.mean(0), np.cov(xb, rowvar=False)
    v = float(fid.calculate_frechet_distance(ma, sa, mb, sb))
    assert abs(v - float(restate.calculate_frechet_distance(ma, sa, mb, sb))) <= 1e-9 * v
    meta["frechet_d192"] = {"seed": 5, "value": v, "self": float(fid.calculate_frechet_distance(ma, sa, ma, sa))}

    # ---- A11: the reference's O-IS function, 40 images, 80 classes, 4 splits
    sd80 = synth.inception_state_dict(3, num_classes=80)
    imgs = torch.cat([synth.images(1, 0, 100, 20), synth.images(2, 1, 100, 20)])
    x = restate.preprocess([im.numpy() for im in imgs], "ois")
    sd80 = synth.calibrate_head(sd80, restate.Trunk(sd80, num_classes=80).pool3(x), restate.T_OIS, seed=3)
    run, _ = reference_shim.load_ois_function(sd80, 80)
    ois_ref = run(list(x), batch_size=8, splits=4)
    t80 = restate.Trunk(sd80, num_classes=80)
    np.save(os.path.join(GOLD, "ois_head.npy"), np.concatenate([sd80["fc.weight"].numpy(), sd80["fc.bias"].numpy()[:, None]], 1))
    preds = restate.ois_preds(t80.logits(x))
    ois_res = restate.ois_inception_score(preds, 4)
    assert abs(ois_res[0] - ois_ref[0]) <= 1e-6 * ois_ref[0], (ois_res, ois_ref)
    meta["ois_40"] = {"mean": float(ois_ref[0]), "std": float(ois_ref[1]), "splits": 4, "weights_seed": 3}
    print("O-IS 40: reference %s restatement %s" % (ois_ref, ois_res))

    # ---- A12/A13: the split/KL lines of the coco script itself (inception_score_star_coco.py:52-59)
    src = open(os.path.join(reference_shim.REFERENCE_ROOT, "image_realism/IS/coco/inception_score_star_coco.py")).read().split("\n")
    body = textwrap.dedent("\n".join(src[52:59]))
    rng = np.random.RandomState(9)
    logits = torch.from_numpy((rng.randn(157, 1008) * 3 + rng.randn(1, 1008) * 2).astype(np.float32))
    p = torch.softmax(logits / torch.tensor(restate.T_COCO), dim=1).numpy()
    ns = {"np": np, "preds": p, "splits": 10}
    exec(body, ns)
    star_ref = (float(np.mean(ns["scores"])), float(np.std(ns["scores"])))
    star_res = restate.isstar_inception_score(p, 10)
    assert abs(star_res[0] - star_ref[0]) <= 1e-7 * star_ref[0]
    meta["isstar_157"] = {"seed": 9, "mean": star_ref[0], "std": star_ref[1]}
    print("IS* 157: reference lines %s restatement %s" % (star_ref, star_res))

    with open(os.path.join(GOLD, "golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote", GOLD)


if __name__ == "__main__":
    main()
