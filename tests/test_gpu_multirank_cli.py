"""-m gpu, needs >= 2 GPUs (skipped on a 1-GPU box; run with `gpurun --gpus 2`): the FID command line launched
under torchrun shards the folder's batches over the ranks, all-reduces the moments once (NCCL) and must print the
same FID as a single process."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch
from PIL import Image

from tise_toolbox_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fid(out):
    return float(re.findall(r"^FID: (\S+)$", out, flags=re.M)[-1])


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_fid_cli_two_ranks_equals_one(tmp_path):
    dirs = []
    for k, (seed, set_id, n) in enumerate([(1, 0, 130), (2, 1, 127)]):
        d = tmp_path / ("set%d" % k)
        d.mkdir()
        for i, im in enumerate(synth.images(seed, set_id, 500, n)):
            Image.fromarray(im.numpy()).save(str(d / ("img_%05d.png" % i)))
        dirs.append(str(d))
    w = str(tmp_path / "w.pth")
    torch.save(synth.inception_state_dict(0, num_classes=1000), w)
    env = dict(os.environ, TISE_INCEPTION_WEIGHTS=w, PYTHONPATH=ROOT)
    args = ["--batch-size", "10", "--path1", dirs[0], "--path2", dirs[1], "--gpu", "0"]
    one = subprocess.run([sys.executable, "-m", "tise_toolbox_b200.fid_score"] + args, env=env, cwd=ROOT, capture_output=True,
                         text=True, timeout=600)
    assert one.returncode == 0, one.stderr[-2000:]
    two = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29541", "-m", "tise_toolbox_b200.fid_score"] + args +
                         ["--saved_file", str(tmp_path / "fid.txt")], env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert two.returncode == 0, two.stderr[-2000:]
    a, b = _fid(one.stdout), _fid(two.stdout)
    # 130 / 127 images in 2048 dimensions: rank-deficient covariances, where tr sqrt(s1 s2) responds to the 1e-16
    # summation-order differences of the fp64 moments with O(sqrt(eps)) -- measured 1.6e-8 relative on 2 x B200
    assert abs(a - b) <= 1e-6 * abs(a), (a, b)
    assert open(str(tmp_path / "fid.txt")).read() == "FID: %s" % re.findall(r"^FID: (\S+)$", two.stdout, flags=re.M)[-1]
    assert two.stdout.count("FID: ") == 1  # rank 0 alone reports


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_engines_on_two_devices_in_one_process():
    """Kernel attributes (opt-in shared memory) are per device and tise_ctx_create must not move the caller's current
    device: a second Engine on cuda:1 in the same process gives the same features as the one on cuda:0."""
    from tise_toolbox_b200.engine import Engine, Moments

    sd = synth.inception_state_dict(0, num_classes=1000)
    imgs = synth.images(1, 0, 0, 24)
    torch.cuda.set_device(0)
    e0 = Engine(sd, device="cuda:0", micro_batch=16)
    f0 = e0.features(imgs.to("cuda:0"))
    e1 = Engine(sd, device="cuda:1", micro_batch=16)
    assert torch.cuda.current_device() == 0  # ctx creation left torch's device alone
    f1 = e1.features(imgs.to("cuda:1"))
    assert torch.equal(f0.cpu(), f1.cpu())
    st = []
    for e, f in ((e0, f0), (e1, f1)):
        m = Moments(e)
        m.update(f)
        st.append(m.finalize())
    assert np.array_equal(st[0][1], st[1][1])
    v1, _ = e1.frechet(st[1][0], st[1][1], st[1][0] + 0.01, st[1][1] * 1.1)  # 64 KB dynamic smem kernels on device 1
    v0, _ = e0.frechet(st[0][0], st[0][1], st[0][0] + 0.01, st[0][1] * 1.1)
    assert v0 == v1
    e1.close()
    assert torch.cuda.current_device() == 0
