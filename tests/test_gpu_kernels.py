"""-m gpu: kernel-level parity of libtise_b200.so against the CPU oracle / torch-CPU fp32."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import restate
from tise_toolbox_b200 import synth

pytestmark = pytest.mark.gpu


# ----------------------------------------------------------------------------- A2 resize (bit-exact)
@pytest.mark.parametrize("h,w", [(256, 256), (64, 48), (500, 375), (299, 299), (17, 333), (480, 16), (299, 100),
                                 (1, 1), (1024, 777)])
def test_resize_bit_exact_uniform(engine, h, w):
    imgs = synth.images(7, 0, 0, 3, h, w)
    out = engine.resize(imgs.cuda()).cpu().numpy()
    for i in range(3):
        ref = restate.pillow_resize_bilinear(imgs[i].numpy())
        assert np.array_equal(out[i], ref), "image %d differs in %d bytes" % (i, int((out[i] != ref).sum()))


def test_resize_bit_exact_ragged(engine):
    sizes = synth.crop_sizes(3, 0, 12) + [(299, 299), (16, 480), (480, 16)]
    imgs = [synth.images(9, 1, i, 1, h, w)[0] for i, (h, w) in enumerate(sizes)]
    out = engine.resize([t.cuda() for t in imgs]).cpu().numpy()
    for i, t in enumerate(imgs):
        ref = restate.pillow_resize_bilinear(t.numpy())
        assert np.array_equal(out[i], ref), "crop %d %s differs" % (i, tuple(t.shape))


def test_resize_empty(engine):
    out = engine.resize(torch.empty((0, 32, 32, 3), dtype=torch.uint8, device="cuda"))
    assert tuple(out.shape) == (0, 299, 299, 3)


# ----------------------------------------------------------------------------- conv engines, shape by shape
def _conv_case(engine, impl, n, h, w, cin, cout, kh, kw, stride, ph, pw, passes=2, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(n, h, w, cin, generator=g).abs() * 0.8).half()  # NHWC, post-ReLU-like
    wt = torch.randn(cout, cin, kh, kw, generator=g) * (2.0 / (cin * kh * kw)) ** 0.5
    scale = torch.rand(cout, generator=g) + 0.5
    shift = torch.randn(cout, generator=g) * 0.2
    ho = (h + 2 * ph - kh) // stride + 1
    wo = (w + 2 * pw - kw) // stride + 1
    xd = x.cuda()
    yd = torch.full((n, ho, wo, cout), -7.0, dtype=torch.float16, device="cuda")
    wn, sn, bn = wt.numpy(), scale.numpy(), shift.numpy()
    rc = engine.lib.tise_debug_conv(engine.ctx, impl, passes, xd.data_ptr(), n, h, w, cin, wn.ctypes.data,
                                    sn.ctypes.data, bn.ctypes.data, cout, kh, kw, stride, ph, pw, 1, yd.data_ptr())
    engine._check(rc)
    weff = wt if passes == 2 else wt.half().float()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), weff, stride=stride, padding=(ph, pw))
    ref = torch.relu(ref * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)).permute(0, 2, 3, 1)
    got = yd.float().cpu()
    err = (got - ref).abs().max().item()
    tol = 2e-3 * max(1.0, ref.abs().max().item())  # fp16 output rounding (2^-11 relative) + accumulation order
    if err > tol or not torch.isfinite(got).all():
        bad = ((got - ref).abs() > tol) | ~torch.isfinite(got)
        rows = bad.reshape(-1, cout).any(dim=1).nonzero().flatten()
        cols = bad.reshape(-1, cout).any(dim=0).nonzero().flatten()
        raise AssertionError(
            "max |err| %.3e > %.3e (ref max %.3f); %d/%d elements off; bad rows %d (first %s) bad cols %d (first %s); "
            "got[0,0,0,:4]=%s ref=%s untouched(-7)=%d" % (
                err, tol, ref.abs().max().item(), int(bad.sum()), bad.numel(), rows.numel(), rows[:8].tolist(),
                cols.numel(), cols[:8].tolist(), got[0, 0, 0, :4].tolist(), ref[0, 0, 0, :4].tolist(),
                int((got == -7).sum())))


CONV_SHAPES = [
    # n, h, w, cin, cout, kh, kw, stride, ph, pw   (shapes of torchvision Inception3)
    (2, 17, 17, 768, 192, 1, 1, 1, 0, 0),
    (3, 35, 35, 192, 64, 1, 1, 1, 0, 0),
    (2, 35, 35, 64, 96, 3, 3, 1, 1, 1),
    (2, 35, 35, 48, 64, 5, 5, 1, 2, 2),
    (1, 73, 73, 80, 192, 3, 3, 1, 0, 0),
    (1, 73, 73, 64, 80, 1, 1, 1, 0, 0),
    (1, 75, 75, 32, 32, 3, 3, 1, 0, 0),
    (1, 75, 75, 32, 64, 3, 3, 1, 1, 1),
    (2, 35, 35, 288, 384, 3, 3, 2, 0, 0),
    (3, 17, 17, 128, 128, 1, 7, 1, 0, 3),
    (3, 17, 17, 160, 192, 7, 1, 1, 3, 0),
    (2, 17, 17, 192, 320, 3, 3, 2, 0, 0),
    (5, 8, 8, 1280, 448, 1, 1, 1, 0, 0),
    (5, 8, 8, 384, 384, 1, 3, 1, 0, 1),
    (5, 8, 8, 384, 384, 3, 1, 1, 1, 0),
    (3, 8, 8, 448, 384, 3, 3, 1, 1, 1),
    (40, 8, 8, 2048, 320, 1, 1, 1, 0, 0),   # more M tiles than one wave of n tiles
    (1, 4, 5, 16, 16, 1, 1, 1, 0, 0),       # a single partial tile
]


@pytest.mark.parametrize("shape", CONV_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_conv_simt(engine, shape):
    _conv_case(engine, 1, *shape)


@pytest.mark.parametrize("shape", CONV_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_conv_umma(engine, shape):
    _conv_case(engine, 0, *shape)


def test_conv_umma_single_pass(engine):
    _conv_case(engine, 0, 2, 17, 17, 768, 192, 1, 1, 1, 0, 0, passes=1)


# Batches large enough that the tcgen05 engine pairs two M tiles per work item (conv_umma.cu: msub = 2).
PAIRED_SHAPES = [
    (40, 35, 35, 64, 96, 3, 3, 1, 1, 1),     # nt 96: two MMAs per K step and tile; odd number of M tiles
    (40, 35, 35, 48, 64, 5, 5, 1, 2, 2),     # nt 64: hi | lo planes in one MMA; 3-step chunks
    (150, 17, 17, 128, 128, 1, 7, 1, 0, 3),  # nt 128: accumulators fill the 256-column TMEM stage
    (150, 17, 17, 160, 192, 7, 1, 1, 3, 0),  # two N tiles, short last chunk
    (20, 73, 73, 80, 192, 3, 3, 1, 0, 0),
    (300, 8, 8, 448, 384, 3, 3, 1, 1, 1),    # three N tiles of 128
    (64, 35, 35, 288, 240, 1, 1, 1, 0, 0),   # nt 240 stays unpaired
]


@pytest.mark.parametrize("shape", PAIRED_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_conv_umma_paired_tiles(engine, shape):
    _conv_case(engine, 0, *shape)


def test_conv_umma_paired_tiles_single_pass(engine):
    _conv_case(engine, 0, 40, 35, 35, 64, 96, 3, 3, 1, 1, 1, passes=1)


# ----------------------------------------------------------------------------- statistics kernels
def test_logits(engine, state_dict):
    g = torch.Generator().manual_seed(1)
    feat = torch.rand(37, 2048, generator=g) * 3
    ref = feat @ state_dict["fc.weight"].t() + state_dict["fc.bias"]
    got = engine.logits(feat.cuda(), use_bias=True).cpu()
    assert torch.allclose(got, ref, rtol=1e-5, atol=1e-3)
    got = engine.logits(feat.cuda(), use_bias=False).cpu()
    assert torch.allclose(got, feat @ state_dict["fc.weight"].t(), rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("n", [1, 50, 300, 1000])
def test_moments_match_numpy(engine, n):
    from tise_toolbox_b200.engine import Moments

    g = torch.Generator().manual_seed(n)
    feat = (torch.rand(n, 2048, generator=g) * 0.3 + 1.7).float()
    m = Moments(engine)
    for i in range(0, n, 128):  # additive over batches
        m.update(feat[i:i + 128].cuda())
    if n == 1:
        return
    mu, sigma = m.finalize()
    act = feat.numpy().astype(np.float64)
    rmu, rsig = restate.calculate_activation_statistics(act)
    assert np.allclose(mu, rmu, rtol=1e-13, atol=0)
    assert np.abs(sigma - rsig).max() <= 1e-12 * max(1.0, np.abs(rsig).max())
    assert np.array_equal(sigma, sigma.T)


def _is_reference(logits, variant, splits):
    if variant == "ois":
        return restate.ois_inception_score(restate.ois_preds(logits), splits)
    if variant == "coco":
        x = logits / torch.tensor(restate.T_COCO, dtype=torch.float32)
        return restate.isstar_inception_score(F.softmax(x, dim=1).numpy(), splits)
    return restate.isstar_inception_score(restate.bird_preds(logits), splits)


@pytest.mark.parametrize("variant,n,c", [("ois", 203, 80), ("coco", 157, 1008), ("bird", 128, 51)])
def test_inception_score_partials(engine, variant, n, c):
    from tise_toolbox_b200 import inception_score as isc

    g = torch.Generator().manual_seed(c)
    logits = torch.randn(n, c, generator=g) * 3 + torch.randn(1, c, generator=g) * 2
    mean, std = isc.score_from_logits(engine, logits.cuda(), variant, splits=10)
    rmean, rstd = _is_reference(logits, variant, 10)
    assert abs(mean - rmean) <= 1e-5 * rmean, (mean, rmean)
    assert abs(std - rstd) <= 1e-4 * max(rstd, 1e-3), (std, rstd)


@pytest.mark.parametrize("align", [True, False])
def test_upsample_float_matches_aten(engine, align):
    """tise_upsample_bilinear_f32 vs torch's CPU upsample_bilinear2d (the op behind F.upsample / nn.Upsample in the
    reference, inception.py:117-118 align_corners=True, object_centric_inception_score.py:49 align_corners=False).
    Tolerance, inputs in [0, 1]: align_corners=True agrees to one float32 rounding (1.2e-7 measured); for
    align_corners=False ATen's CPU kernel rounds the source index / weights in a slightly different sequence than its CUDA
    kernel, whose float arithmetic this kernel follows: up to 3.4e-6 on these sizes, far below anything the score sees."""
    import torch.nn.functional as F

    g = torch.Generator().manual_seed(3)
    for (h, w) in [(64, 48), (300, 511), (299, 299), (1, 1), (17, 333), (256, 256)]:
        x = torch.rand(2, 3, h, w, generator=g)
        ref = F.interpolate(x, size=(299, 299), mode="bilinear", align_corners=align)
        got = engine.upsample_float(x, align_corners=align).cpu()
        assert float((got - ref).abs().max()) <= (5e-7 if align else 1e-5), (h, w, align)
    assert tuple(engine.upsample_float(torch.empty(0, 3, 8, 8), align_corners=align).shape) == (0, 3, 299, 299)


def test_channel_affine_is_bit_identical_to_the_eager_ops(engine):
    """InceptionV3.forward's normalize_input pre-amble (image_realism/FID/inception.py:119-124) as ONE libtise kernel:
    same bits as the reference's three eager `x[:, c] = x[:, c] * s + b` statements (multiply and add rounded apart)."""
    g = torch.Generator().manual_seed(5)
    x = torch.rand(5, 3, 37, 41, generator=g)
    ref = x.clone()
    ref[:, 0] = ref[:, 0] * (0.229 / 0.5) + (0.485 - 0.5) / 0.5
    ref[:, 1] = ref[:, 1] * (0.224 / 0.5) + (0.456 - 0.5) / 0.5
    ref[:, 2] = ref[:, 2] * (0.225 / 0.5) + (0.406 - 0.5) / 0.5
    got = engine.channel_affine(x.cuda(), (0.229 / 0.5, 0.224 / 0.5, 0.225 / 0.5),
                                ((0.485 - 0.5) / 0.5, (0.456 - 0.5) / 0.5, (0.406 - 0.5) / 0.5))
    assert torch.equal(got.cpu(), ref)
