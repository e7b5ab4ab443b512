"""CPU (-m "not gpu"): the oracle restatement against the committed golden vectors, which
tools/make_golden.py produced by running the reference's own source (oracle/reference_shim.py)."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import reference_shim, restate
from tise_toolbox_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(GOLD, "golden.json")) as f:
        return json.load(f)


def test_synthetic_images_are_reproducible(golden):
    for g in golden["resize"]:
        im = synth.images(g["seed"], g["set"], g["idx"], 1, g["h"], g["w"])[0].numpy()
        assert sha(im) == g["in_sha256"]


def test_pillow_restatement_matches_golden_and_pillow(golden):
    for g in golden["resize"]:
        im = synth.images(g["seed"], g["set"], g["idx"], 1, g["h"], g["w"])[0].numpy()
        out = restate.pillow_resize_bilinear(im)
        assert sha(out) == g["out_sha256"], (g["h"], g["w"])
        assert np.array_equal(out, np.asarray(Image.fromarray(im).resize((299, 299), Image.BILINEAR)))


def test_resize_edge_cases():
    one = np.full((1, 1, 3), 200, dtype=np.uint8)
    assert (restate.pillow_resize_bilinear(one) == 200).all()
    im = synth.images(1, 0, 0, 1, 299, 299)[0].numpy()
    assert np.array_equal(restate.pillow_resize_bilinear(im), im)  # both passes skipped


def test_input_luts_match_reference_ops():
    lut = restate.input_lut("fid")
    x = torch.arange(256, dtype=torch.uint8).view(1, 256, 1).repeat(1, 1, 3)  # HWC with every level
    t = x.permute(2, 0, 1).float().div(255).unsqueeze(0).clone()  # ToTensor
    t[:, 0] = t[:, 0] * (0.229 / 0.5) + (0.485 - 0.5) / 0.5
    t[:, 1] = t[:, 1] * (0.224 / 0.5) + (0.456 - 0.5) / 0.5
    t[:, 2] = t[:, 2] * (0.225 / 0.5) + (0.406 - 0.5) / 0.5
    assert torch.equal(lut, t[0, :, 0, :])
    from torchvision import transforms

    pil = Image.fromarray(np.arange(256, dtype=np.uint8).reshape(1, 256, 1).repeat(3, 2))
    o = transforms.Normalize((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))(transforms.ToTensor()(pil))
    assert torch.equal(restate.input_lut("ois"), o[:, 0, :])
    from tise_toolbox_b200.engine import input_lut

    for v in restate.INPUT_VARIANTS:  # the product builds the same tables
        assert torch.equal(input_lut(v), restate.input_lut(v))


def test_weights_are_reproducible(golden, state_dict):
    for k, h in golden["weights"].items():
        assert sha(state_dict[k].numpy()) == h, k


def test_pool3_and_fid_match_reference_golden(golden, state_dict):
    trunk = restate.Trunk(state_dict)
    sets = [synth.images(1, 0, 0, 16), synth.images(2, 1, 0, 16)]
    acts = []
    for s in sets:
        x = restate.preprocess([im.numpy() for im in s], "fid")
        acts.append(restate.get_activations([x[0:8], x[8:16]], trunk, 8))
    ref = np.load(os.path.join(GOLD, "pool3_ref.npy"))
    got = np.concatenate([acts[0][:4], acts[1][:4]])
    assert np.abs(got - ref).max() <= 2e-5 * np.abs(ref).max()
    st = [restate.calculate_activation_statistics(a) for a in acts]
    v = restate.calculate_frechet_distance(st[0][0], st[0][1], st[1][0], st[1][1])
    assert abs(v - golden["fid_2x16"]["value"]) <= 1e-5 * golden["fid_2x16"]["value"]


def _d192():
    rng = np.random.RandomState(5)
    d = 192
    xa, xb = rng.randn(400, d) * rng.rand(d) + rng.randn(d), rng.randn(500, d) * rng.rand(d) * 1.3 + rng.randn(d)
    return xa.mean(0), np.cov(xa, rowvar=False), xb.mean(0), np.cov(xb, rowvar=False)


def test_frechet_restatement_matches_golden(golden):
    ma, sa, mb, sb = _d192()
    v = restate.calculate_frechet_distance(ma, sa, mb, sb)
    assert abs(v - golden["frechet_d192"]["value"]) <= 1e-9 * v
    # the reference's one transferable known answer: FID(X, X) ~ 0 (O-FID/real_images.txt = -2.6e-12)
    assert abs(restate.calculate_frechet_distance(ma, sa, ma, sa)) <= 1e-7


def test_frechet_error_behaviour():
    ma, sa, mb, sb = _d192()
    with pytest.raises(AssertionError):
        restate.calculate_frechet_distance(ma[:-1], sa, mb, sb)
    with pytest.raises(AssertionError):
        restate.calculate_frechet_distance(ma, sa[:-1, :-1], mb, sb)


def test_isstar_restatement_matches_golden(golden):
    rng = np.random.RandomState(9)
    logits = torch.from_numpy((rng.randn(157, 1008) * 3 + rng.randn(1, 1008) * 2).astype(np.float32))
    p = torch.softmax(logits / torch.tensor(restate.T_COCO), dim=1).numpy()
    m, s = restate.isstar_inception_score(p, 10)
    assert abs(m - golden["isstar_157"]["mean"]) <= 1e-6 * m and abs(s - golden["isstar_157"]["std"]) <= 1e-4 * s


def test_ois_restatement_matches_reference_golden(golden):
    sd80 = synth.inception_state_dict(3, num_classes=80)
    head = np.load(os.path.join(GOLD, "ois_head.npy"))
    sd80["fc.weight"] = torch.from_numpy(head[:, :2048].copy())
    sd80["fc.bias"] = torch.from_numpy(head[:, 2048].copy())
    imgs = torch.cat([synth.images(1, 0, 100, 20), synth.images(2, 1, 100, 20)])
    x = restate.preprocess([im.numpy() for im in imgs], "ois")
    preds = restate.ois_preds(restate.Trunk(sd80, num_classes=80).logits(x))
    m, s = restate.ois_inception_score(preds, 4)
    assert abs(m - golden["ois_40"]["mean"]) <= 1e-5 * m
    assert abs(s - golden["ois_40"]["std"]) <= 1e-3 * s
    # one-pass form used by the CUDA path: mean KL = S/n - sum m log m
    n = 10
    part = preds[:n]
    mm = part.mean(0)
    one_pass = np.exp((part * np.log(part)).sum() / n - (mm * np.log(mm)).sum())
    two_pass = restate.ois_inception_score(part, 1)[0]
    assert abs(one_pass - two_pass) <= 1e-7 * two_pass


@pytest.mark.skipif(not reference_shim.available(), reason="/root/reference is only present in the build container")
def test_reference_frechet_function_itself(golden, state_dict):
    fid = reference_shim.load_fid_module(state_dict)
    ma, sa, mb, sb = _d192()
    assert abs(float(fid.calculate_frechet_distance(ma, sa, mb, sb)) - golden["frechet_d192"]["value"]) <= 1e-9
    with pytest.raises(RuntimeError):
        fid.calculate_fid_given_paths(["/nonexistent/a", "/nonexistent/b"], 50, "", 2048)
