"""Where does the end-to-end loop lose against the resident loop?  Times K steps of B images in five forms:
  resident | resident + unrelated H2D traffic on a side stream | from host, no read-back | from host + read-back |
  resident + read-back
  python tools/e2e_probe.py [--batch 4096] [--steps 10]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tise_toolbox_b200 import synth  # noqa: E402
from tise_toolbox_b200.engine import Engine, T_BIRD  # noqa: E402
from tise_toolbox_b200.pipeline import RealismJob  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--micro-batch", type=int, default=1024)
ap.add_argument("--steps", type=int, default=10)
a = ap.parse_args()
B, K = a.batch, a.steps
dev = torch.device("cuda", 0)
sd = synth.inception_state_dict(0, num_classes=51)
eng = Engine(sd, device=dev, variant="bird", micro_batch=a.micro_batch)
pool = torch.cat([synth.images(1, 0, i, 256, device=dev) for i in range(0, B, 256)])
sd = synth.calibrate_head(sd, eng.features(pool[:64]).cpu(), T_BIRD, class_offset=1)
eng.load_state_dict(sd)
host = torch.empty((2 * B, 256, 256, 3), dtype=torch.uint8).pin_memory()
host[:B].copy_(pool.cpu())
host[B:].copy_(pool.cpu())
hb = lambda s: host[(s % 2) * B:(s % 2) * B + B]


def timed(fn, name):
    for s in range(2):
        fn(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(K):
        fn(s)
    done()
    e1.record()
    torch.cuda.synchronize()
    print("%-44s %8.0f img/s" % (name, B * K / (e0.elapsed_time(e1) / 1e3)), flush=True)


pend = []


def done():
    while pend:
        pend.pop(0).wait()


def mk():
    return RealismJob(eng, n_total=B * (K + 2), is_variant="bird", splits=10, input_variant="bird")


job = mk()
timed(lambda s: job.step(pool), "resident")
side = torch.cuda.Stream(device=dev)
junk = torch.empty_like(pool)
job = mk()


def with_traffic(s):
    with torch.cuda.stream(side):
        junk.copy_(hb(s), non_blocking=True)
    job.step(pool)


timed(with_traffic, "resident + unrelated H2D on a side stream")
job = mk()
timed(lambda s: job.step_from_host(hb(s), hb(s + 1)), "from host, no read-back")
job = mk()


def full(s):
    job.step_from_host(hb(s), hb(s + 1))
    pend.append(job.read_partials_async())
    if len(pend) > 1:
        pend.pop(0).wait()


timed(full, "from host + read-back (bench e2e)")
job = mk()


def rb(s):
    job.step(pool)
    pend.append(job.read_partials_async())
    if len(pend) > 1:
        pend.pop(0).wait()


timed(rb, "resident + read-back")
job = mk()
timed(lambda s: job.step(pool), "resident (again)")

# ---- read-back variants
job = mk()
snap_host = [torch.empty(job.parts.f.numel(), dtype=torch.float64).pin_memory() for _ in range(2)]
evs = [None, None]


class _H:
    def __init__(self, ev):
        self.ev = ev

    def wait(self):
        self.ev.synchronize()


def full_same_stream(s):
    job.step_from_host(hb(s), hb(s + 1))
    k = s % 2
    snap_host[k].copy_(job.parts.f, non_blocking=True)  # D2H in order on the compute stream
    ev = torch.cuda.Event()
    ev.record()
    pend.append(_H(ev))
    if len(pend) > 1:
        pend.pop(0).wait()


timed(full_same_stream, "from host + read-back on the compute stream")
job = mk()


def full_no_wait(s):
    job.step_from_host(hb(s), hb(s + 1))
    pend.append(job.read_partials_async())
    if len(pend) > 3:
        pend.pop(0).wait()


timed(full_no_wait, "from host + read-back, awaited 3 steps later")

# ---- paced H2D (chunks released by the trunk's progress event) vs one copy per step
for pace in (False, True, False, True):
    job = mk()
    job.pace_h2d = pace
    timed(full, "from host + read-back, paced H2D = %s" % pace)
