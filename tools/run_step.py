"""One steady-state step of bench.py's hot path (RealismJob.step: resize -> trunk -> moments -> fc -> IS* sums) between
cudaProfilerStart / cudaProfilerStop, for ncu --profile-from-start off.

  python tools/run_step.py [--batch 1024] [--micro-batch 1024] [--steps 1]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tise_toolbox_b200 import synth  # noqa: E402
from tise_toolbox_b200.engine import Engine, T_BIRD  # noqa: E402
from tise_toolbox_b200.pipeline import RealismJob  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--micro-batch", type=int, default=1024)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--warmup", type=int, default=2)
a = ap.parse_args()
dev = torch.device("cuda", 0)
sd = synth.inception_state_dict(0, num_classes=51)
eng = Engine(sd, device=dev, variant="bird", micro_batch=a.micro_batch)
pool = torch.cat([synth.images(1, 0, i, min(256, a.batch - i), device=dev) for i in range(0, a.batch, 256)])
sd = synth.calibrate_head(sd, eng.features(pool[:64]).cpu(), T_BIRD, class_offset=1)
eng.load_state_dict(sd)
job = RealismJob(eng, n_total=a.batch * (a.steps + a.warmup), is_variant="bird", splits=10, input_variant="bird")
for _ in range(a.warmup):
    job.step(pool)
torch.cuda.synchronize()
torch.cuda.profiler.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    job.step(pool)
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("step: %.3f ms for %d images" % (e0.elapsed_time(e1) / a.steps, a.batch))
