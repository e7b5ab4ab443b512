"""-m gpu: the Inception-v3 trunk through libtise_b200 against the CPU oracle (torchvision Inception3 fp32)."""
import numpy as np
import pytest
import torch

from oracle import restate
from tise_toolbox_b200 import synth

pytestmark = pytest.mark.gpu

# pool3 relative L2 bound of the fp16-activation x (hi+lo)-fp16-weight model: measured 1.4e-4 on CPU
# emulation (tools/precision_probe.py); single fp16 pass: 5.9e-4.
POOL3_REL_L2 = {2: 5e-4, 1: 2e-3}


@pytest.fixture(scope="module")
def batch():
    imgs = torch.cat([synth.images(1, 0, 0, 5), synth.images(2, 1, 0, 4)])  # 9 images: odd batch, two sets
    return imgs


@pytest.fixture(scope="module")
def oracle_blocks(batch, state_dict):
    trunk = restate.Trunk(state_dict)
    x = restate.preprocess([b.numpy() for b in batch], "fid")
    return x, trunk.blocks(x)


def _rel_l2(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.mark.parametrize("impl,passes", [("simt", 2), ("umma", 2), ("umma", 1)])
def test_pool3_parity(engine, batch, oracle_blocks, impl, passes):
    engine.set_conv_mode(impl, passes)
    engine.set_variant("fid")
    try:
        got = engine.features(batch.cuda()).cpu().numpy()
    finally:
        engine.set_conv_mode("umma", 2)
    ref = oracle_blocks[1][3].reshape(len(batch), -1).numpy()
    assert got.shape == ref.shape == (9, 2048)
    assert np.isfinite(got).all()
    err = _rel_l2(got, ref)
    assert err <= POOL3_REL_L2[passes], "pool3 rel L2 %.3e" % err


def test_pool3_parity_default_policy(engine, batch, oracle_blocks):
    """The default operand policy: the tcgen05 implicit-GEMM layers run the W_hi pass only, with the mean-field W_lo
    correction folded into their shift at load time (tise_calibrate_single_pass).  Same pool3 bound as the strict
    two-pass policy; the block-0 tap (the flat stem layers, single pass too) stays within its own bound."""
    from tise_toolbox_b200.engine import POOL3_REL_L2_BOUND, SINGLE_PASS_ALL

    engine.set_variant("fid")
    strict = engine.features(batch.cuda()).cpu().numpy()
    tap64 = engine.features(batch.cuda(), dims=64).cpu().numpy()
    engine.set_pass_policy(SINGLE_PASS_ALL)
    try:
        single = engine.features(batch.cuda()).cpu().numpy()
        tap64_single = engine.features(batch.cuda(), dims=64).cpu().numpy()
    finally:
        engine.set_pass_policy(0)
    ref = oracle_blocks[1][3].reshape(len(batch), -1).numpy()
    assert engine.calibrated
    assert not np.array_equal(single, strict)  # the policy is live
    assert _rel_l2(single, ref) <= POOL3_REL_L2_BOUND["default"], "pool3 rel L2 %.3e" % _rel_l2(single, ref)
    assert _rel_l2(strict, ref) <= POOL3_REL_L2_BOUND["strict"]
    ref64 = oracle_blocks[1][0].mean(dim=(2, 3)).numpy()
    assert not np.array_equal(tap64, tap64_single)
    assert _rel_l2(tap64_single, ref64) <= 1e-3 and _rel_l2(tap64, ref64) <= 1e-3


@pytest.mark.parametrize("dims,block", [(64, 0), (192, 1), (768, 2)])
def test_block_taps(engine, batch, oracle_blocks, dims, block):
    got = engine.features(batch.cuda(), dims=dims).cpu().numpy()
    ref = oracle_blocks[1][block].mean(dim=(2, 3)).numpy()
    assert got.shape == ref.shape
    assert _rel_l2(got, ref) <= 1e-3


def test_float_input_path_matches_u8_path(engine, batch, oracle_blocks):
    x = oracle_blocks[0]  # float32 [n,3,299,299], already normalised by the oracle
    a = engine.features_from_float(x.cuda()).cpu().numpy()
    b = engine.features(batch.cuda()).cpu().numpy()
    # Same floats in (the LUT is bit-exact), but the float path runs Conv2d_1a on the CUDA cores in fp32 and the uint8
    # path on the tensor cores (input and weights as fp16 hi + lo pairs): equal up to the fp16 rounding of activations.
    assert _rel_l2(a, b) <= 3e-4


def test_micro_batching_is_invisible(engine, batch):
    full = engine.features(batch.cuda()).cpu().numpy()
    keep = engine.micro_batch
    engine.micro_batch = 4
    try:
        part = engine.features(batch.cuda()).cpu().numpy()
    finally:
        engine.micro_batch = keep
    assert np.array_equal(full, part)


def test_forward_resizes_float_inputs_of_other_sizes(state_dict):
    """Quirk 3 of SURVEY.md §9: a caller that bypasses the Dataset and feeds non-299 float tensors gets the reference's
    F.upsample(align_corners=True) (inception.py:117-118) -- here a libtise kernel -- before the trunk."""
    import torch.nn.functional as F

    from tise_toolbox_b200.inception import InceptionV3

    model = InceptionV3([3], state_dict=state_dict, micro_batch=8, single_pass_max_hw=0)
    x = synth.images(1, 0, 0, 4).permute(0, 3, 1, 2).float().div(255)  # 256 x 256 floats in [0, 1]
    a = model(x)[0].reshape(4, -1).cpu().numpy()
    b = model(F.interpolate(x, size=(299, 299), mode="bilinear", align_corners=True))[0].reshape(4, -1).cpu().numpy()
    assert a.shape == (4, 2048)
    # (inputs equal to ~1e-7; fp16 activation roundings downstream may flip: same bound as the float-vs-uint8 path check)
    assert np.linalg.norm(a - b) / np.linalg.norm(b) <= 5e-4
