"""One profiled trunk pass (after a warm-up) -> per-launch dump in $TISE_PROFILE_DUMP.  python tools/profile_trunk.py [n]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tise_toolbox_b200 import synth  # noqa: E402
from tise_toolbox_b200.engine import Engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sd = synth.inception_state_dict(0, num_classes=51)
eng = Engine(sd, device="cuda:0", variant="fid", micro_batch=n, passes=int(os.environ.get("TISE_PASSES", "2")))
imgs = torch.cat([synth.images(1, 0, i, min(128, n - i), device="cuda") for i in range(0, n, 128)])
u8 = eng.resize(imgs)
for _ in range(2):
    eng.features_from_resized(u8)
torch.cuda.synchronize()
eng.profile(True)
eng.features_from_resized(u8)
print(eng.profile_read())
