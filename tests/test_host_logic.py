"""CPU (-m "not gpu"): host-side logic of the path -- split assignment, sharding, result formats and
the N>1 exchange step over gloo (world_size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import restate
from tise_toolbox_b200 import inception_score as isc
from tise_toolbox_b200 import pipeline


@pytest.mark.parametrize("n,splits", [(23, 10), (100, 10), (7, 3), (30000, 10), (9, 10)])
def test_split_ids_follow_the_reference_slicing(n, splits):
    ois = isc.split_ids(n, splits, "ois")
    star = isc.split_ids(n, splits, "coco")
    per = n // splits
    for k in range(splits):
        assert np.array_equal(np.nonzero(ois == k)[0], np.arange(k * per, (k + 1) * per))  # O-IS :73
        assert np.array_equal(np.nonzero(star == k)[0], np.arange(k * n // splits, (k + 1) * n // splits))  # coco :55
    assert (ois[per * splits:] == -1).all() and (star >= 0).all()


def test_split_ids_are_independent_of_sharding():
    n = 1003
    full = isc.split_ids(n, 10, "coco")
    for world in (2, 3, 8):
        parts = []
        for r in range(world):
            lo, hi = pipeline.shard_range(n, r, world)
            parts.append(isc.split_ids(n, 10, "coco", lo, hi - lo))
        assert np.array_equal(np.concatenate(parts), full)


def test_split_ids_device_form_equals_host_form():
    for variant in ("ois", "coco", "bird"):
        for n, splits in [(100, 10), (101, 10), (7, 10), (64, 3), (30000, 10), (5, 1)]:
            for start, count in [(0, n), (n // 3, n - n // 3), (n - 1, 4)]:
                host = isc.split_ids(n, splits, variant, start, count)
                dev = isc.split_ids_device(n, splits, variant, start, count, "cpu")
                assert dev.dtype == torch.int32 and np.array_equal(host, dev.numpy()), (variant, n, splits, start, count)


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 30000, 200001):
        for world in (1, 2, 4, 8):
            r = [pipeline.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1


def test_kept_images_bird_drops_partial_batch():
    assert isc.kept_images(130, "bird") == 128 and isc.kept_images(130, "coco") == 130


def test_one_pass_is_equals_reference_two_pass():
    rng = np.random.RandomState(0)
    logits = torch.from_numpy(rng.randn(97, 80).astype(np.float32) * 2)
    preds = restate.ois_preds(logits)
    ref = restate.ois_inception_score(preds, 10)
    sid = isc.split_ids(97, 10, "ois")
    scores = []
    for k in range(10):
        part = preds[sid == k]
        m = part.mean(0)
        scores.append(np.exp((part * np.log(part)).sum() / len(part) - (m * np.log(m)).sum()))
    assert abs(np.mean(scores) - ref[0]) <= 1e-8 * ref[0] and abs(np.std(scores) - ref[1]) <= 1e-6 * ref[1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, d, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.RandomState(1)
    x = rng.rand(n, d) + 1.5
    lo, hi = pipeline.shard_range(n, rank, world)
    # the packed exchange buffer of the path: [count | sum x | sum x x^T]  (engine.Moments.buf)
    buf = torch.zeros(1 + d + d * d, dtype=torch.float64)
    xs = torch.from_numpy(x[lo:hi])
    buf[0] = hi - lo
    buf[1:1 + d] = xs.sum(0)
    buf[1 + d:] = (xs.t() @ xs).reshape(-1)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    if rank == 0:
        cnt = buf[0].item()
        mu = buf[1:1 + d] / cnt
        sigma = (buf[1 + d:].view(d, d) - cnt * torch.outer(mu, mu)) / (cnt - 1)
        torch.save({"mu": mu, "sigma": sigma, "cnt": cnt}, out)
    dist.destroy_process_group()


def test_two_rank_moment_exchange_over_gloo(tmp_path):
    n, d = 101, 16
    out = str(tmp_path / "r.pt")
    mp.spawn(_worker, args=(2, _free_port(), n, d, out), nprocs=2, join=True)
    r = torch.load(out)
    rng = np.random.RandomState(1)
    x = rng.rand(n, d) + 1.5
    assert r["cnt"] == n
    assert np.allclose(r["mu"].numpy(), x.mean(0), rtol=1e-13)
    assert np.allclose(r["sigma"].numpy(), np.cov(x, rowvar=False), rtol=1e-9, atol=1e-13)


def _pack_worker(rank, world, port, out):
    """StatsPack.allreduce over gloo with a stub engine (device = cpu): the packing / counter round trip is host logic."""
    from types import SimpleNamespace

    from tise_toolbox_b200.engine import StatsPack

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = SimpleNamespace(device=torch.device("cpu"))
    pack = StatsPack(eng, d=8, is_classes=5, is_splits=3)
    assert pack.buf.numel() == 1 + 8 + 64 + 3 + 15 + 3 + 1
    pack.moments.count = 10 + rank
    pack.moments.sum.add_(float(rank + 1))
    pack.moments.sxx.add_(0.5)
    pack.parts.f.add_(float(rank + 2))
    pack.parts.i.add_(torch.tensor([1, 2, 3, rank], dtype=torch.int64) * (rank + 1))
    pack.allreduce()
    if rank == 0:
        torch.save({"count": pack.moments.count, "sum": pack.moments.sum.clone(), "sxx": pack.moments.sxx.clone(),
                    "f": pack.parts.f.clone(), "i": pack.parts.i.clone()}, out)
    dist.destroy_process_group()


def test_stats_pack_single_allreduce_over_gloo(tmp_path):
    out = str(tmp_path / "p.pt")
    mp.spawn(_pack_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["count"] == 21
    assert torch.equal(r["sum"], torch.full((8,), 3.0, dtype=torch.float64))
    assert torch.equal(r["sxx"], torch.full((64,), 1.0, dtype=torch.float64))
    assert torch.equal(r["f"], torch.full((18,), 5.0, dtype=torch.float64))
    assert r["i"].tolist() == [3, 6, 9, 2]


def test_result_file_formats(tmp_path):
    assert pipeline.format_result("fid", 26.581254719518654) == "FID: 26.581254719518654"
    assert pipeline.format_result("ofid", -2.6147972675971687e-12) == "O-FID: -2.6147972675971687e-12"
    assert pipeline.format_result("coco", 54.63, 1.54) == "[Inception Score] mean: 54.63000 std: 1.54000"
    assert pipeline.format_result("bird", 15.13, 0.2) == "IS = 15.13  +-  0.2"
    assert pipeline.format_result("ois", 8.63, 0.06) == "O-IS: 8.63 +-  0.06"


def test_ranking_table_matches_the_reference_script_golden(tmp_path):
    """tests/golden/calibration.json["ranking"]: the reference's ranking_score.py run on its own methods/*.json."""
    import json
    from collections import OrderedDict

    from tise_toolbox_b200 import ranking_score as rs

    with open(os.path.join(os.path.dirname(__file__), "golden", "calibration.json")) as f:
        g = json.load(f)["ranking"]
    ordered = OrderedDict((m, g["methods"][m]) for m in g["listdir_order"])
    assert rs.ranking_table(method_scores=ordered) == g["table"]
    # emission in the reference's schema, then the directory form
    for m in g["listdir_order"]:
        rs.update_method_scores(str(tmp_path), m, **{k: g["methods"][m][k] for k in rs.METRICS})
    p = rs.update_method_scores(str(tmp_path), g["listdir_order"][0], FID=g["methods"][g["listdir_order"][0]]["FID"])
    assert json.load(open(p))["FID"] == float(g["methods"][g["listdir_order"][0]]["FID"])
    scores = rs.ranking_scores(ordered)
    table_dir = rs.ranking_table(str(tmp_path))
    for m in ordered:
        assert ("%g" % scores[m][-1]) in table_dir
    with pytest.raises(KeyError):
        rs.update_method_scores(str(tmp_path), "x", BLEU=1.0)


def test_shard_batches_partition_is_exactly_the_drop_last_set():
    for n, b, world in [(1000, 50, 8), (1003, 50, 8), (23, 5, 2), (7, 10, 4), (30000, 64, 8), (49, 50, 3)]:
        ranges = [pipeline.shard_batches(n, b, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == (n // b) * b
        for (lo, hi), (lo2, _) in zip(ranges, ranges[1:]):
            assert hi == lo2
        assert all((hi - lo) % b == 0 and hi >= lo for lo, hi in ranges)


def test_shard_format_round_trip_and_validation(tmp_path):
    from tise_toolbox_b200 import shards, synth

    imgs = [synth.images(3, 0, i, 1, h, w)[0].numpy() for i, (h, w) in enumerate([(256, 256), (1, 1), (17, 333), (64, 48)])]
    p = str(tmp_path / "a.shard")
    nbytes = shards.write_shard(p, imgs)
    assert nbytes == os.path.getsize(p)
    offs, hw, size = shards.read_index(p)
    assert hw.tolist() == [[256, 256], [1, 1], [17, 333], [64, 48]] and not (offs % 16).any() and size == nbytes
    back = shards.read_images(p)
    assert all(np.array_equal(a, b) for a, b in zip(imgs, back))
    assert np.array_equal(shards.read_images(p, [2])[0], imgs[2])
    shards.write_shard(str(tmp_path / "empty.shard"), [])
    assert len(shards.read_index(str(tmp_path / "empty.shard"))[0]) == 0
    with pytest.raises(ValueError):
        shards.write_shard(str(tmp_path / "bad.shard"), [np.zeros((4, 4), np.uint8)])
    with open(p, "rb") as f:
        blob = bytearray(f.read())
    with open(str(tmp_path / "trunc.shard"), "wb") as f:
        f.write(blob[:len(blob) // 2])
    with pytest.raises(ValueError):
        shards.read_index(str(tmp_path / "trunc.shard"))
    with open(str(tmp_path / "magic.shard"), "wb") as f:
        f.write(b"NOTASHARD" + blob[9:])
    with pytest.raises(ValueError):
        shards.read_index(str(tmp_path / "magic.shard"))
    # entries out of file order (index rows 0 and 2 swapped): every entry is in bounds, but the loader's span
    # arithmetic would address outside the uploaded bytes
    swapped = bytearray(blob)
    e0, e2 = bytes(swapped[16:32]), bytes(swapped[48:64])
    swapped[16:32], swapped[48:64] = e2, e0
    with open(str(tmp_path / "order.shard"), "wb") as f:
        f.write(swapped)
    with pytest.raises(ValueError):
        shards.read_index(str(tmp_path / "order.shard"))


def test_device_image_loader_grouping_with_a_stub_engine(tmp_path):
    """Host logic of img_data.DeviceImageLoader without a GPU: order, drop_last, decode groups that are whole numbers of
    consumer batches, batch slices of a group."""
    from tise_toolbox_b200 import img_data
    from tise_toolbox_b200.engine import PackedImages

    class StubEngine:
        def __init__(self):
            self.staged_groups = []

        def stage_files(self, paths, pool=None):
            self.staged_groups.append(list(paths))
            return {"paths": list(paths)}

        def decode_staged(self, staged):
            n = len(staged["paths"])
            ids = np.array([int(os.path.basename(p)[1:-4]) for p in staged["paths"]], dtype=np.int64)
            return PackedImages(torch.zeros(16, dtype=torch.uint8), ids * 48, np.tile(np.array([[4, 4]], np.int32), (n, 1)))

    names = [str(tmp_path / ("f%04d.png" % i)) for i in range(103)]
    eng = StubEngine()
    loader = img_data.DeviceImageLoader(names, eng, batch_size=10, drop_last=True, decode_batch=35)
    assert len(loader) == 10 and loader.group == 30
    seen = []
    for batch in loader:
        assert len(batch) == 10
        seen += (batch.offsets // 48).tolist()
    assert seen == list(range(100))  # the last 3 files are dropped, order kept
    assert [len(g) for g in eng.staged_groups] == [30, 30, 30, 10]
    eng = StubEngine()
    loader = img_data.DeviceImageLoader(names, eng, batch_size=10, drop_last=False, decode_batch=1000)
    sizes = [len(b) for b in loader]
    assert sizes == [10] * 10 + [3] and len(loader) == 11
    assert sum(len(g) for g in img_data.DeviceImageLoader(names[:7], StubEngine(), 10, drop_last=True).groups()) == 0
    # ramp-up and chaining: a short first group (the GPU starts after ~1000 files), and the NEXT loader's first group is
    # staged before this loader's last group is decoded (fid_score.calculate_fid_given_paths walks two folders like this)
    events = []

    class Recorder(StubEngine):
        def stage_files(self, paths, pool=None):
            events.append(("stage", os.path.basename(paths[0])[0], len(paths)))
            return super().stage_files(paths, pool)

        def decode_staged(self, staged):
            import time

            time.sleep(0.2)  # (the GPU work of a group; the next group's file reads must have started by its end)
            events.append(("decoded", os.path.basename(staged["paths"][0])[0], len(staged["paths"])))
            return super().decode_staged(staged)

    eng = Recorder()
    a = img_data.DeviceImageLoader([str(tmp_path / ("a%04d.png" % i)) for i in range(5020)], eng, 50, drop_last=True)
    b = img_data.DeviceImageLoader([str(tmp_path / ("b%04d.png" % i)) for i in range(1200)], eng, 50, drop_last=True)
    assert [len(g) for g in a._group_lists()] == [1000, 4000] and [len(g) for g in b._group_lists()] == [1200]
    n_a = sum(len(p) for p in a.groups(next_loader=b))
    n_b = sum(len(p) for p in b.groups())
    assert (n_a, n_b) == (5000, 1200)
    assert events.index(("stage", "b", 1200)) < events.index(("decoded", "a", 4000))
    assert events.index(("stage", "a", 4000)) < events.index(("decoded", "a", 1000))


def test_build_is_serialised_across_processes(tmp_path):
    """torchrun starts one process per GPU and every one of them calls build_library(): with a fresh library none of
    them rebuilds, and the lock file / atomic rename keep concurrent callers from corrupting the .so."""
    import subprocess
    import sys

    from tise_toolbox_b200 import build

    before = os.path.getmtime(build.build_library())
    code = "from tise_toolbox_b200 import build; print(build.build_library())"
    procs = [subprocess.Popen([sys.executable, "-c", code], cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                              stdout=subprocess.PIPE, text=True) for _ in range(4)]
    outs = [p.communicate(timeout=600)[0].strip() for p in procs]
    assert all(p.returncode == 0 for p in procs) and all(o.endswith("libtise_b200.so") for o in outs), outs
    assert os.path.getmtime(build.LIB) == before  # nobody rebuilt a fresh library


def test_bench_helpers_ramp_and_isolated_legs():
    """bench.py host logic: the fixed-size run's batch ramp covers the shard exactly, and a CPU leg that fails in its
    fresh interpreter is reported (after the retry) instead of raising or hanging the GPU arm."""
    import importlib.util
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    saved_fd1 = os.dup(1)  # bench.py points fd 1 at fd 2 at import time (its stdout carries exactly one JSON line)
    try:
        spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(root, "bench.py"))
        bench = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(bench)
    finally:
        os.dup2(saved_fd1, 1)
        os.close(saved_fd1)
    for n in (1, 255, 600, 3750, 30000):
        sizes = bench.ramp_sizes(n, 2048)
        assert sum(sizes) == n and all(0 < s <= 2048 for s in sizes)
    assert bench.ramp_sizes(3750, 2048) == [256, 768, 2048, 678] and bench.ramp_sizes(600, 2048) == [256, 344]
    out, err = bench.run_isolated(["--child", "cpu_leg:/nonexistent/spec.pt"], timeout_s=120, tries=2)
    assert out is None and "exit code" in err
