#!/usr/bin/env python3
"""GPU experiment (gpurun): FID / pool3 error of the trunk against the CPU oracle at cfg1 size (1000 vs 1000) for the
operand policies of tise_set_conv_mode / tise_set_pass_policy, with the trunk throughput of each.

    python tools/pass_policy_probe.py [n_per_set] [out.json]

The oracle features are also saved (gpurun_out/oracle_pool3_cfg1.npz) so that later analysis needs no CPU re-run.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import harness, restate  # noqa: E402
from tise_toolbox_b200 import synth  # noqa: E402
from tise_toolbox_b200.engine import Engine, Moments  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    out_path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "pass_policy.json")
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    sd = synth.inception_state_dict(0, num_classes=1000)
    sets = [synth.images(1, 0, 0, n), synth.images(2, 1, 0, n)]
    t0 = time.perf_counter()
    ref = [harness.oracle_pool3(sd, s, "fid") for s in sets]
    cpu_s = time.perf_counter() - t0
    np.savez_compressed(os.path.join(os.path.dirname(out_path), "oracle_pool3_cfg1.npz"), a=ref[0], b=ref[1])
    eng = Engine(sd, device="cuda:0", micro_batch=500)  # mean-field W_lo correction calibrated at load
    os.environ["TISE_LO_CALIBRATION"] = "0"
    eng_raw = Engine(sd, device="cuda:0", micro_batch=500)  # single-pass layers without the correction
    del os.environ["TISE_LO_CALIBRATION"]

    def gpu_fid(fa, fb):
        st = []
        for f in (fa, fb):
            m = Moments(eng)
            m.update(torch.as_tensor(f, device="cuda").float().contiguous())
            st.append(m.finalize_device())
        return eng.frechet(st[0][0], st[0][1], st[1][0], st[1][1])[0]

    fid_ref = gpu_fid(ref[0], ref[1])
    t0 = time.perf_counter()
    fid_scipy = harness.fid_from_features(ref[0], ref[1])
    scipy_s = time.perf_counter() - t0
    res = {"n_per_set": n, "cpu_oracle_s": cpu_s, "cpu_cores": os.cpu_count(), "fid_oracle_scipy": fid_scipy,
           "fid_oracle_feats_gpu_frechet": fid_ref, "scipy_fid_s": scipy_s, "policies": []}
    dsets = [s.cuda() for s in sets]
    for name, e, passes, hw in [("strict (2 passes)", eng, 2, 0),
                                ("1pass hw<=17 calibrated", eng, 2, 17), ("1pass hw<=35 calibrated", eng, 2, 35),
                                ("1pass all umma layers calibrated (default)", eng, 2, 1 << 20),
                                ("1pass hw<=17 uncorrected", eng_raw, 2, 17), ("1pass all umma layers uncorrected", eng_raw, 2, 1 << 20)]:
        e.set_conv_mode("umma", passes)
        e.set_pass_policy(hw)
        gpu_feats = e.features
        feats = [gpu_feats(d) for d in dsets]  # also warms up the plans of this policy
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            for d in dsets:
                gpu_feats(d)
        e1.record()
        torch.cuda.synchronize()
        ips = 3 * 2 * n / (e0.elapsed_time(e1) / 1e3)
        f = [x.cpu().numpy() for x in feats]
        fid = gpu_fid(f[0], f[1])
        allf, allr = np.concatenate(f), np.concatenate(ref)
        row = {"policy": name, "fid": fid, "fid_rel_err": abs(fid - fid_ref) / abs(fid_ref),
               "pool3_rel_l2": float(np.linalg.norm(allf - allr) / np.linalg.norm(allr)),
               "mu_rel": float(np.linalg.norm(allf.mean(0) - allr.mean(0)) / np.linalg.norm(allr.mean(0))),
               "images_per_s": ips}
        print(json.dumps(row), flush=True)
        res["policies"].append(row)
    with open(out_path, "w") as fo:
        json.dump(res, fo, indent=1)
    print(json.dumps({k: v for k, v in res.items() if k != "policies"}))


if __name__ == "__main__":
    main()
