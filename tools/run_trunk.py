"""Runs the trunk once or a few times on synthetic images (used under ncu).  python tools/run_trunk.py [n] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tise_toolbox_b200 import synth  # noqa: E402
from tise_toolbox_b200.engine import Engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
sd = synth.inception_state_dict(0, num_classes=51)
eng = Engine(sd, device="cuda:0", variant="fid", micro_batch=n, passes=int(os.environ.get("TISE_PASSES", "2")))
imgs = torch.cat([synth.images(1, 0, i, min(128, n - i), device="cuda") for i in range(0, n, 128)])
u8 = eng.resize(imgs)
for _ in range(reps):
    f = eng.features_from_resized(u8)
torch.cuda.synchronize()
print("ok", float(f.mean()))
