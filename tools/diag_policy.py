"""GPU diagnostic: operand policies (strict / default / uncorrected single pass) against the fp32 CPU oracle on
(a) the O-IS crop set of tests/test_gpu_e2e.py and (b) the FID sets.  Prints pool3 rel L2, |dmu|/|mu| and IS.

  python tools/diag_policy.py [n_crops] [n_fid]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import harness, restate  # noqa: E402
from tise_toolbox_b200 import synth  # noqa: E402
from tise_toolbox_b200.engine import SINGLE_PASS_ALL, Engine  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def main():
    n_crops = int(sys.argv[1]) if len(sys.argv) > 1 else 26
    n_fid = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    harness.use_all_cores()
    sd80 = synth.inception_state_dict(3, num_classes=80)
    head = np.load(os.path.join(GOLD, "ois_head.npy"))
    sd80["fc.weight"] = torch.from_numpy(head[:, :2048].copy())
    sd80["fc.bias"] = torch.from_numpy(head[:, 2048].copy())
    sizes = synth.crop_sizes(8, 0, n_crops, lo=24, hi=320)
    crops = [synth.images(5, i % 2, i, 1, h, w)[0] for i, (h, w) in enumerate(sizes)]
    trunk = restate.Trunk(sd80, num_classes=80)
    x = restate.preprocess([c.numpy() for c in crops], "ois")
    ref_pool = trunk.pool3(x).numpy()
    ref_logits = trunk.logits(x)
    ref_is = restate.ois_inception_score(restate.ois_preds(ref_logits), 2)
    print("O-IS oracle: %.6f +- %.6f" % ref_is)
    for name, hw, cal in (("strict", 0, "1"), ("default", SINGLE_PASS_ALL, "1"), ("uncorrected", SINGLE_PASS_ALL, "0")):
        os.environ["TISE_LO_CALIBRATION"] = cal
        eng = Engine(sd80, device="cuda:0", variant="ois", micro_batch=16, single_pass_max_hw=hw)
        f = eng.features([c.cuda() for c in crops])
        lg = eng.logits(f).cpu()
        got = restate.ois_inception_score(restate.ois_preds(lg), 2)
        fn = f.cpu().numpy()
        print("crops %-12s pool3 rel L2 %.2e  |dmu|/|mu| %.2e  logits rel %.2e  O-IS %.6f rel %.2e"
              % (name, rel(fn, ref_pool), rel(fn.mean(0), ref_pool.mean(0)), rel(lg.numpy(), ref_logits.numpy()),
                 got[0], abs(got[0] - ref_is[0]) / ref_is[0]), flush=True)
        eng.close()
    sd = synth.inception_state_dict(0, num_classes=1000)
    imgs = torch.cat([synth.images(1, 0, 0, n_fid // 2), synth.images(2, 1, 0, n_fid // 2)])
    ref = harness.oracle_pool3(sd, imgs, "fid", batch=50)
    for name, hw, cal in (("strict", 0, "1"), ("default", SINGLE_PASS_ALL, "1"), ("uncorrected", SINGLE_PASS_ALL, "0")):
        os.environ["TISE_LO_CALIBRATION"] = cal
        eng = Engine(sd, device="cuda:0", variant="fid", micro_batch=64, single_pass_max_hw=hw)
        f = eng.features(imgs.cuda()).cpu().numpy()
        print("fid   %-12s pool3 rel L2 %.2e  |dmu|/|mu| %.2e" % (name, rel(f, ref), rel(f.mean(0), ref.mean(0))), flush=True)
        eng.close()


if __name__ == "__main__":
    main()
