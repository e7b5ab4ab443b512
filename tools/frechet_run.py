"""One tise_frechet call at d = 2048 on seeded covariances (for ncu launch lists / timing).
    python tools/frechet_run.py [--d 2048] [--reps 3]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tise_toolbox_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--d", type=int, default=2048)
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
g = torch.Generator(device="cpu").manual_seed(0)
d = a.d
basis = torch.randn(d, d, generator=g, dtype=torch.float64) / d ** 0.5
xa = (torch.randn(a.n, d, generator=g, dtype=torch.float64) * torch.linspace(0.05, 1.0, d, dtype=torch.float64)) @ basis + 1.7
xb = (torch.randn(a.n, d, generator=g, dtype=torch.float64) * torch.linspace(0.08, 1.2, d, dtype=torch.float64)) @ basis + 1.8
ma, mb = xa.mean(0).cuda(), xb.mean(0).cuda()
sa, sb = torch.cov(xa.t()).cuda(), torch.cov(xb.t()).cuda()
eng = Engine(None, device="cuda:0")
for r in range(a.reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    v, flags = eng.frechet(ma, sa, mb, sb)
    print("rep %d: %.3f ms  value %.9f flags %d" % (r, 1e3 * (time.perf_counter() - t0), v, flags))
