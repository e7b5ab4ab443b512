"""Device PNG decode throughput next to Pillow on the host cores.
    python tools/png_bench.py [--n 2048] [--unique 256] [--out gpurun_out/png_bench.json]"""
import argparse
import io
import json
import os
import sys
import time

import numpy as np
import torch
from PIL import Image

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tise_toolbox_b200 import synth  # noqa: E402
from tise_toolbox_b200.engine import Engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2048)
    ap.add_argument("--unique", type=int, default=256)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    imgs = synth.images(1, 0, 0, a.unique)
    blobs = []
    for im in imgs:
        b = io.BytesIO()
        Image.fromarray(im.numpy()).save(b, "PNG")
        blobs.append(b.getvalue())
    batch = [blobs[i % a.unique] for i in range(a.n)]
    eng = Engine(None, device="cuda:0")
    packed = eng.decode_images(batch)
    assert np.array_equal(packed[a.n - 1].cpu().numpy(), imgs[(a.n - 1) % a.unique].numpy())
    torch.cuda.synchronize()
    times = []
    for _ in range(a.reps):
        t0 = time.perf_counter()
        eng.decode_images(batch)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    # kernels only: files already in HBM -> call the C entry point directly
    t0 = time.perf_counter()
    for b in blobs[:64]:
        np.asarray(Image.open(io.BytesIO(b)).convert("RGB"))
    pil = 64 / (time.perf_counter() - t0)
    res = {"n": a.n, "file_bytes_mean": float(np.mean([len(b) for b in blobs])),
           "decode_images_per_s_e2e_best": a.n / min(times), "decode_images_per_s_e2e_median": a.n / float(np.median(times)),
           "pil_images_per_s_one_core": pil, "host_cores": os.cpu_count()}
    print(json.dumps(res))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(res, f)


if __name__ == "__main__":
    main()
