"""-m gpu: parity with the CPU oracle AT BASELINE.json's sizes.

cfg1 -- FID, 1 000 vs 1 000 synthetic 256 x 256 images, batch 50, through ``fid_score.calculate_fid_given_paths`` on PNG
        folders (/root/reference/image_realism/FID/fid_score.py:223-238) and through ``RealismJob``: FID within 1e-3
        relative of the oracle, pool3 relative L2 and |d mu| / |mu| printed and bounded.
cfg2 -- IS* with temperature scaling, 10 splits, on a 2 000-image subset of the 30 k set, bird (51-way, class 0 sliced,
        T = 0.598; inception_score_star_bird.py:98-108,187-194) and coco (1008-way, no bias, T = 0.909;
        inception_score_star_coco.py:53-60,104-108) heads, calibrated on BOTH sides: IS within 1e-3 relative.

The oracle (torchvision Inception3 fp32 on the host cores, PIL resize, numpy / scipy statistics) needs about a minute
per 2 000 images on the GPU box's 16 cores.
"""
import os

import numpy as np
import pytest
import torch

from oracle import harness, restate
from tise_toolbox_b200 import fid_score, synth
from tise_toolbox_b200.engine import POOL3_REL_L2_BOUND, SINGLE_PASS_ALL, T_BIRD, T_COCO, Engine
from tise_toolbox_b200.inception import InceptionV3
from tise_toolbox_b200.pipeline import RealismJob, fid_from_stats

pytestmark = pytest.mark.gpu

N_CFG1 = int(os.environ.get("TISE_PARITY_N", "1000"))   # images per set (cfg1: 1000 vs 1000)
N_CFG2 = int(os.environ.get("TISE_PARITY_N_IS", "2000"))  # cfg2 subset
FID_TOL = 1e-3   # BASELINE.json north_star: "FID and IS within 1e-3 relative"
IS_TOL = 1e-3
# stated pool3 bound, relative L2 over the whole set: 5e-4 for both operand policies -- the default
# (W_hi pass + calibrated mean-field W_lo correction in the tcgen05 implicit-GEMM layers) and strict (two passes everywhere)
POLICIES = [("default", SINGLE_PASS_ALL), ("strict", 0)]


@pytest.fixture(scope="module")
def cfg1(state_dict):
    sets = [synth.images(1, 0, 0, N_CFG1), synth.images(2, 1, 0, N_CFG1)]
    ref_feats = [harness.oracle_pool3(state_dict, s, "fid", batch=50) for s in sets]
    ref_stats = [restate.calculate_activation_statistics(f.astype(np.float64)) for f in ref_feats]
    ref_fid = float(restate.calculate_frechet_distance(ref_stats[0][0], ref_stats[0][1], ref_stats[1][0], ref_stats[1][1]))
    return {"sets": sets, "feats": ref_feats, "stats": ref_stats, "fid": ref_fid}


@pytest.mark.parametrize("policy,max_hw", POLICIES)
def test_cfg1_fid_1000_vs_1000_realism_job(state_dict, cfg1, policy, max_hw):
    eng = Engine(state_dict, device="cuda:0", micro_batch=500, single_pass_max_hw=max_hw)
    res, feats = [], []
    for s in cfg1["sets"]:
        job = RealismJob(eng, n_total=len(s))
        f = [job.step(s[i:i + 50].cuda()) for i in range(0, len(s), 50)]  # the reference's batches of 50
        feats.append(torch.cat(f).cpu().numpy())
        res.append(job.finish())
    got = fid_from_stats(eng, res[0], res[1])
    allf, allr = np.concatenate(feats), np.concatenate(cfg1["feats"])
    rel_l2 = float(np.linalg.norm(allf - allr) / np.linalg.norm(allr))
    mu_rel = max(float(np.linalg.norm(r["mu"].cpu().numpy() - st[0]) / np.linalg.norm(st[0]))
                 for r, st in zip(res, cfg1["stats"]))
    fid_rel = abs(got - cfg1["fid"]) / abs(cfg1["fid"])
    print("cfg1 RealismJob [%s]: FID %.6f oracle %.6f rel %.2e | pool3 rel L2 %.2e | |dmu|/|mu| %.2e"
          % (policy, got, cfg1["fid"], fid_rel, rel_l2, mu_rel))
    assert rel_l2 <= POOL3_REL_L2_BOUND[policy] and mu_rel <= POOL3_REL_L2_BOUND[policy]
    assert fid_rel <= FID_TOL, (got, cfg1["fid"])


def test_cfg1_fid_1000_vs_1000_cli_on_png_folders(tmp_path, state_dict, cfg1):
    """The reference's own entry point on two folders of PNG files, batch 50 (fid_score.py:223-238): device PNG decode,
    resize, trunk, moments, Frechet distance."""
    from concurrent.futures import ThreadPoolExecutor

    from PIL import Image

    dirs = []
    for k, s in enumerate(cfg1["sets"]):
        d = tmp_path / ("set%d" % k)
        d.mkdir()
        arr = s.numpy()

        def save(i, d=d, arr=arr):
            Image.fromarray(arr[i]).save(str(d / ("img_%05d.png" % i)), compress_level=1)

        with ThreadPoolExecutor(16) as ex:
            list(ex.map(save, range(len(arr))))
        dirs.append(str(d))
    model = InceptionV3([3], state_dict=state_dict, micro_batch=500)  # default operand policy, as the CLI runs it
    got = float(fid_score.calculate_fid_given_paths(dirs, 50, "0", 2048, model=model))
    fid_rel = abs(got - cfg1["fid"]) / abs(cfg1["fid"])
    print("cfg1 CLI on PNG folders: FID %.6f oracle %.6f rel %.2e" % (got, cfg1["fid"], fid_rel))
    assert fid_rel <= FID_TOL, (got, cfg1["fid"])


@pytest.mark.parametrize("variant,classes,temp,off", [("bird", 51, T_BIRD, 1), ("coco", 1008, T_COCO, 0)])
def test_cfg2_isstar_subset_calibrated_head(variant, classes, temp, off):
    n = N_CFG2
    if variant == "bird":
        n = (n // 64) * 64  # the bird script keeps floor(N / 64) full batches (inception_score_star_bird.py:83)
    sd = synth.inception_state_dict(0, num_classes=classes)
    imgs = synth.images(1, 0, 0, n)
    trunk = restate.Trunk(sd, num_classes=classes)
    pool3 = torch.from_numpy(harness.oracle_pool3(sd, imgs, variant, num_classes=classes, batch=64, trunk=trunk))
    sd = synth.calibrate_head(sd, pool3[:256], temp, seed=classes, class_offset=off)  # same head on both sides
    trunk = restate.Trunk(sd, num_classes=classes)
    with torch.no_grad():
        if variant == "coco":
            ref = restate.isstar_inception_score(restate.coco_preds(pool3, sd["fc.weight"]), 10)
        else:
            ref = restate.isstar_inception_score(restate.bird_preds(trunk.net.fc(pool3)), 10)
    assert ref[0] > 1.5, "vacuous IS target"
    eng = Engine(sd, device="cuda:0", variant=variant, micro_batch=500)
    for policy, max_hw in POLICIES:
        eng.set_pass_policy(max_hw)
        job = RealismJob(eng, n_total=n, is_variant=variant, splits=10, input_variant=variant, want_moments=False)
        feats = []
        for i in range(0, n, 1000):
            feats.append(job.step(imgs[i:i + 1000].cuda()))
        res = job.finish()
        f = torch.cat(feats).cpu().numpy()
        rel_l2 = float(np.linalg.norm(f - pool3.numpy()) / np.linalg.norm(pool3.numpy()))
        is_rel = abs(res["is_mean"] - ref[0]) / ref[0]
        print("cfg2 %s [%s]: IS* %.6f +- %.6f oracle %.6f +- %.6f rel %.2e | pool3 rel L2 %.2e"
              % (variant, policy, res["is_mean"], res["is_std"], ref[0], ref[1], is_rel, rel_l2))
        assert rel_l2 <= POOL3_REL_L2_BOUND[policy]
        assert is_rel <= IS_TOL, (policy, res["is_mean"], ref)
        assert abs(res["is_std"] - ref[1]) <= 2e-2 * ref[1] + 1e-4, (policy, res["is_std"], ref)
