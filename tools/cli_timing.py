"""Where the wall time of a whole FID command goes (weights -> engine, files -> statistics, Frechet distance).
    python tools/cli_timing.py [--n 4000]"""
import argparse
import os
import shutil
import sys
import tempfile
import time

import torch
from PIL import Image

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tise_toolbox_b200 import fid_score, synth  # noqa: E402
from tise_toolbox_b200.inception import InceptionV3  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4000)
a = ap.parse_args()
tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
try:
    dirs = []
    t0 = time.perf_counter()
    for k, (seed, sid) in enumerate([(1, 0), (2, 1)]):
        d = os.path.join(tmp, "set%d" % k)
        os.makedirs(d)
        base = synth.images(seed, sid, 0, 200).numpy()
        blobs = []
        for i in range(200):
            p = os.path.join(d, "b_%05d.png" % i)
            Image.fromarray(base[i]).save(p)
            blobs.append(open(p, "rb").read())
        for i in range(200, a.n):  # replicate the encoded files (encoding is the slow part of this setup)
            with open(os.path.join(d, "r_%05d.png" % i), "wb") as f:
                f.write(blobs[i % 200])
        dirs.append(d)
    print("setup (PNG encode on the host) %.2f s" % (time.perf_counter() - t0))
    sd = synth.inception_state_dict(0, num_classes=1000)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    model = InceptionV3([3], state_dict=sd)  # default micro-batch, as the command line runs it
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print("engine: weights folded / split / uploaded in %.3f s" % (t1 - t0))
    for rep in range(2):
        t1 = time.perf_counter()
        v = float(fid_score.calculate_fid_given_paths(dirs, 50, "0", 2048, model=model))  # the command line's own call
        t3 = time.perf_counter()
        print("run %d: FID of 2 x %d PNG files (batch 50), files -> scalar: %.3f s = %.0f img/s; FID %.6f"
              % (rep, a.n, t3 - t1, 2 * a.n / (t3 - t1), v))
finally:
    shutil.rmtree(tmp, ignore_errors=True)
