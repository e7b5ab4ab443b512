#!/usr/bin/env python3
"""Benchmark of the image-realism hot path on B200 (contract: see the task prompt / DESIGN.md §Measurement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference] [--workload cfg2|cfg4]

One "step" = one batch of B synthetic 256x256 uint8 images per rank through the whole path:
Pillow-exact resize -> Inception-v3 trunk (tcgen05) -> pool3 -> FID moments + fc head + IS* partial sums.
N = 1 is BASELINE.json configs[1] (IS* on CUB-shaped images, bird head, 10 splits; the FID moments ride
along), N > 1 is configs[2] (FID + IS* sharded, one all-reduce at the end).  Besides the steady-state `value`
and `e2e` the line carries `fixed` -- north_star's target itself: 30 000 images from pinned host memory to the
FID and IS* scalars on rank 0, all-reduce and Frechet distance included, wall seconds -- and `parity`, the deltas
against the CPU reference path on a sample of the same images from the same run.
`--workload cfg4` measures BASELINE.json configs[3] instead (O-IS / O-FID on ragged object crops).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONV_FLOP_PER_IMAGE = 2 * 5711168096  # 94 convolutions of torchvision Inception3 at 299x299
IMG_BYTES = 256 * 256 * 3


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--batch", type=int, default=6144, help="images per rank per step (20 steps cover >= 2 s)")
    p.add_argument("--micro-batch", type=int, default=1024)
    p.add_argument("--pool", type=int, default=6144, help="distinct synthetic images resident per rank")
    p.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg4"])
    p.add_argument("--fixed", type=int, default=30000, help="images of the fixed-size end-to-end run (0: skip)")
    p.add_argument("--crops", type=int, default=100000, help="cfg4: number of ragged crops")
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--child", default="", help=argparse.SUPPRESS)  # internal: isolated CPU legs (see run_isolated)
    p.add_argument("--passes", type=int, default=2)
    p.add_argument("--strict", action="store_true", help="two MMA passes in every layer (default: mixed operand policy)")
    p.add_argument("--cpu-sample", type=int, default=500, help="images of the CPU reference leg (two sets of half)")
    p.add_argument("--ref-step", type=int, default=200, help="--impl reference: images per step")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-frechet", action="store_true")
    p.add_argument("--no-ingest", action="store_true", help="skip the PNG-ingest measurement (SURVEY.md §8f row 1)")
    return p.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"tflops": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1398.3))),
                "tflops_burst": float(d.get("bf16_tflops", 1668.8)),
                "hbm_gbs": float(d.get("hbm_gbs", 6551.0)), "source": "MEASURED_PEAKS.json (sustained cuBLAS bf16)"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback of B200_PROFILING.md"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md 'clocks line'): one streaming
    nvidia-smi process (-lms 100); only lines that arrive between mark_start() and mark_stop() are summarised."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.t0, self.t1 = index, [], None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def run(self):
        if not self.proc:
            return
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.strip().split(",")]))

    def mark_start(self):
        self.t0 = time.perf_counter()

    def mark_stop(self):
        self.t1 = time.perf_counter()

    def median_mhz(self, t0, t1):
        """Median SM clock of the samples between two perf_counter() stamps (other phases of the run)."""
        v = [float(r[0]) for t, r in self.rows if t0 <= t <= t1 + 0.15 and len(r) >= 7 and r[0].replace(".", "", 1).isdigit()]
        return float(np.median(v)) if v else None

    def close(self):
        if self.proc:
            self.proc.terminate()

    def summary(self):
        rows = [r for t, r in self.rows if len(r) >= 7 and self.t0 is not None and self.t0 <= t <= (self.t1 or 1e30) + 0.15]
        if not rows:  # very short timed region: fall back to everything seen while the job ran
            rows = [r for t, r in self.rows if len(r) >= 7]
        num = lambda v: float(v) if v.replace(".", "", 1).isdigit() else None
        sm = [num(r[0]) for r in rows if num(r[0]) is not None]
        mx = [num(r[1]) for r in rows if num(r[1]) is not None]
        pw = [num(r[2]) for r in rows if num(r[2]) is not None]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


# --------------------------------------------------------------------------------------------- CPU arms
def _write_pngs(imgs, folder):
    """uint8 [n, H, W, 3] -> folder/img_00000.png ... (lossless; untimed set-up of the CPU legs)."""
    from concurrent.futures import ThreadPoolExecutor

    from PIL import Image

    os.makedirs(folder, exist_ok=True)
    arr = imgs.cpu().numpy() if torch.is_tensor(imgs) else imgs

    def save(i):
        Image.fromarray(arr[i]).save(os.path.join(folder, "img_%05d.png" % i), compress_level=1)

    with ThreadPoolExecutor(8) as ex:
        list(ex.map(save, range(len(arr))))


class ReferenceCpu:
    """The reference's OWN CPU implementation of the path (image_realism/FID/fid_score.py run from git-ignored
    oracle/_ref, see oracle/build_ref.py): ``_compute_statistics_of_path`` = img_data.Dataset + DataLoader(batch 50,
    drop_last, num_workers=8) + PIL Resize + ToTensor + InceptionV3.forward + np.mean / np.cov, unmodified.  The IS*
    scripts of the reference are TensorFlow-1 graphs that cannot run; their split / KL lines are restated in
    oracle/restate.py and applied to the pool3 activations the reference model produced (captured by a forward hook,
    so the trunk runs once, as in the GPU arm).  Falls back to the oracle port (``kind = "port"``) if oracle/_ref is
    absent."""

    def __init__(self, state_dict, num_classes):
        from oracle import harness, reference_shim, restate

        self.cores = harness.use_all_cores()
        self.restate = restate
        self.sd = state_dict
        self.num_classes = num_classes
        self.captured = []
        self.kind = "port"
        self.mod = None
        if reference_shim.available():
            self.mod = reference_shim.load_fid_module(state_dict, num_classes=num_classes)
            self.model = self.mod.InceptionV3([3])
            self.model.register_forward_hook(lambda m, i, o: self.captured.append(o[0].detach().reshape(o[0].shape[0], -1)))
            self.kind = "reference"
        else:
            self.trunk = restate.Trunk(state_dict, num_classes=num_classes)
        self.fc_w = state_dict["fc.weight"].float()
        self.fc_b = state_dict["fc.bias"].float()

    def file_order(self, folder):
        """Global indices of the folder's files in the order the reference's Dataset walks them."""
        from tise_toolbox_b200 import img_data

        names = (self.mod.img_data.Dataset(folder).file_names if self.mod else img_data.Dataset(folder).file_names)
        return [int(os.path.basename(n)[4:9]) for n in names]

    def statistics_of_path(self, folder, batch=50):
        """(mu, sigma, pool3 [n_used, 2048] float32) of one folder."""
        self.captured = []
        if self.mod is not None:
            m, s = self.mod._compute_statistics_of_path(folder, self.model, batch, 2048, "")
            feats = torch.cat(self.captured).float()
        else:  # port: same steps restated (oracle/restate.py)
            from PIL import Image

            from oracle import harness
            from tise_toolbox_b200 import img_data

            names = img_data.Dataset(folder).file_names
            names = names[:len(names) // batch * batch]
            rows = []
            for i0 in range(0, len(names), batch):
                ims = [np.asarray(Image.open(n).convert("RGB")) for n in names[i0:i0 + batch]]
                rows.append(self.trunk.pool3(harness.pil_preprocess(ims, "fid")))
            feats = torch.cat(rows)
            m, s = self.restate.calculate_activation_statistics(feats.numpy().astype(np.float64))
        return m, s, feats

    def isstar_bird(self, feats, splits=10):
        with torch.no_grad():
            logits = feats @ self.fc_w.t() + self.fc_b
        return self.restate.isstar_inception_score(self.restate.bird_preds(logits), splits)

    def frechet(self, m1, s1, m2, s2):
        if self.mod is not None:
            return float(self.mod.calculate_frechet_distance(m1, s1, m2, s2))
        return float(self.restate.calculate_frechet_distance(m1, s1, m2, s2))


def cpu_leg(state_dict, n_images, tmp_root):
    """The CPU reference path on `n_images` of the same synthetic images (half from each FID set), timed on the host
    cores; returns the baseline record and what the parity block needs."""
    from tise_toolbox_b200 import synth

    half = max(50, n_images // 2 // 50 * 50)
    sets = [synth.images(1, 0, 0, half), synth.images(2, 1, 0, half)]
    dirs = [os.path.join(tmp_root, "cpu_set%d" % k) for k in range(2)]
    for s, d in zip(sets, dirs):
        _write_pngs(s, d)
    ref = ReferenceCpu(state_dict, int(state_dict["fc.weight"].shape[0]))
    order = [ref.file_order(d) for d in dirs]
    t0 = time.perf_counter()
    st = [ref.statistics_of_path(d, 50) for d in dirs]
    feats = torch.cat([st[0][2], st[1][2]])
    is_mean, is_std = ref.isstar_bird(feats, 10)
    secs = time.perf_counter() - t0
    t0 = time.perf_counter()
    fid = ref.frechet(st[0][0], st[0][1], st[1][0], st[1][1])
    fid_s = time.perf_counter() - t0
    n_used = int(feats.shape[0])
    rec = {"value": n_used / secs, "unit": "images/s", "cores": ref.cores, "kind": ref.kind,
           "sample": "%d of the same synthetic images (2 sets x %d PNG files, batch 50, DataLoader num_workers=8), %.1f s "
                     "for features + statistics + IS*; the Frechet distance (scipy sqrtm) took %.1f s more; "
                     "os.cpu_count()=%s" % (n_used, half, secs, fid_s, os.cpu_count())}
    return rec, {"sets": sets, "order": order, "feats": feats.numpy(), "fid": fid, "is_mean": is_mean, "is_std": is_std,
                 "sqrtm_s": fid_s}


def gpu_parity(eng, cpu):
    """The same images, in the reference's file order, through the GPU path with the reference's FID input map:
    FID(A, B), IS* (bird head, 10 splits) and pool3 against the CPU leg of this run."""
    from tise_toolbox_b200.pipeline import RealismJob, fid_from_stats

    dev = eng.device
    ordered = [s[torch.tensor(o)].to(dev) for s, o in zip(cpu["sets"], cpu["order"])]
    n_used = [len(o) // 50 * 50 for o in cpu["order"]]
    res, feats = [], []
    for imgs, n in zip(ordered, n_used):
        job = RealismJob(eng, n_total=n, input_variant="fid")
        feats.append(job.step(imgs[:n]))
        res.append(job.finish())
    fid = fid_from_stats(eng, res[0], res[1])
    f = torch.cat(feats)
    job = RealismJob(eng, n_total=int(f.shape[0]), is_variant="bird", splits=10, input_variant="fid", want_moments=False)
    job.step(torch.cat([ordered[0][:n_used[0]], ordered[1][:n_used[1]]]))
    isr = job.finish()
    f = f.cpu().numpy()
    return {"pool3_rel_l2": float(np.linalg.norm(f - cpu["feats"]) / np.linalg.norm(cpu["feats"])),
            "fid_rel": abs(fid - cpu["fid"]) / abs(cpu["fid"]), "is_rel": abs(isr["is_mean"] - cpu["is_mean"]) / cpu["is_mean"],
            "fid": fid, "fid_cpu": cpu["fid"], "is_mean": isr["is_mean"], "is_mean_cpu": cpu["is_mean"],
            "images": int(f.shape[0]),
            "note": "same PNG-lossless images and calibrated 51-way head on both sides; trunk input map = the FID affine "
                    "(the CPU leg is the reference's FID module); IS* trunk is torchvision's on both sides because the "
                    "reference's TF graphs cannot run (split / KL lines restated)"}


def run_isolated(child_args, timeout_s, tries=2):
    """The CPU legs fork DataLoader workers (the reference's own code: DataLoader(num_workers=8)) out of a process that
    has already run multi-threaded torch CPU code -- and, in the GPU arm, holds a CUDA context and helper threads.  A
    fork in that state can leave a worker waiting on a lock some other thread held at the moment of the fork (seen once
    in ~10 runs: the whole bench line was lost to a 15-minute hang).  So every CPU leg runs in a FRESH interpreter with a
    timeout and one retry; a leg that still fails is reported as unavailable instead of taking the GPU numbers with it."""
    cmd = [sys.executable, os.path.abspath(__file__)] + list(child_args)
    last = "not run"
    for attempt in range(tries):
        try:
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout_s,
                               env=dict(os.environ, TISE_BENCH_CHILD="1"))
            if r.returncode == 0:
                sys.stderr.write(r.stderr[-2000:])
                return r.stdout, None
            last = "exit code %d: %s" % (r.returncode, r.stderr[-300:].replace("\n", " | "))
        except subprocess.TimeoutExpired:
            last = "timed out after %d s (attempt %d)" % (timeout_s, attempt + 1)
        print("bench: isolated CPU leg failed (%s)%s" % (last, ", retrying" if attempt + 1 < tries else ""), file=sys.stderr)
    return None, last


def child_cpu_leg(spec_path):
    """--child cpu_leg:<spec.pt>: cpu_leg() in a fresh process; the spec holds the calibrated head, the sample size and the
    output path."""
    from tise_toolbox_b200 import synth

    spec = torch.load(spec_path)
    sd = synth.inception_state_dict(0, num_classes=int(spec["fc.weight"].shape[0]))
    sd["fc.weight"], sd["fc.bias"] = spec["fc.weight"], spec["fc.bias"]
    rec, cpu = cpu_leg(sd, int(spec["n_images"]), spec["tmp_root"])
    np.savez(spec["out"], rec=json.dumps(rec), order0=np.array(cpu["order"][0]), order1=np.array(cpu["order"][1]),
             feats=cpu["feats"], fid=cpu["fid"], is_mean=cpu["is_mean"], is_std=cpu["is_std"], sqrtm_s=cpu["sqrtm_s"])


def cpu_leg_isolated(state_dict, n_images, tmp_root):
    """cpu_leg() in a fresh interpreter (run_isolated); returns (record, parity inputs) or (unavailable record, None)."""
    from tise_toolbox_b200 import synth

    spec = os.path.join(tmp_root, "cpu_leg_spec.pt")
    out = os.path.join(tmp_root, "cpu_leg_out.npz")
    torch.save({"fc.weight": state_dict["fc.weight"].cpu(), "fc.bias": state_dict["fc.bias"].cpu(), "n_images": n_images,
                "tmp_root": tmp_root, "out": out}, spec)
    _, err = run_isolated(["--child", "cpu_leg:" + spec], timeout_s=240)
    if err is not None or not os.path.exists(out):
        return {"unavailable": "CPU leg failed twice: %s" % err}, None
    z = np.load(out, allow_pickle=False)
    half = max(50, n_images // 2 // 50 * 50)
    cpu = {"sets": [synth.images(1, 0, 0, half), synth.images(2, 1, 0, half)],
           "order": [z["order0"].tolist(), z["order1"].tolist()], "feats": z["feats"], "fid": float(z["fid"]),
           "is_mean": float(z["is_mean"]), "is_std": float(z["is_std"]), "sqrtm_s": float(z["sqrtm_s"])}
    return json.loads(str(z["rec"])), cpu


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path on the host cores (rank 0 only); each step is
    a bounded sample: `--ref-step` PNG files through _compute_statistics_of_path + the IS* lines."""
    if rank != 0:
        return
    if not os.environ.get("TISE_BENCH_CHILD"):  # fresh interpreter + timeout + one retry (run_isolated)
        per_step = max(50, args.ref_step // 50 * 50)
        budget = int(max(240, (args.steps + max(args.warmup, 0)) * per_step / 30.0 * 1.5 + 90))  # ~30 img/s worst case seen
        out, err = run_isolated([a for a in sys.argv[1:]], timeout_s=budget)
        if out is None:
            _emit(json.dumps({"impl": "reference", "unavailable": "reference arm failed twice: %s" % err}))
        else:
            _emit(out.strip().splitlines()[-1])
        return
    import shutil
    import tempfile

    from tise_toolbox_b200 import synth

    per_step = max(50, args.ref_step // 50 * 50)
    sd = synth.inception_state_dict(0, num_classes=51)
    tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        folder = os.path.join(tmp, "step")
        _write_pngs(synth.images(1, 0, 0, per_step), folder)
        ref = ReferenceCpu(sd, 51)

        def step():
            m, s, feats = ref.statistics_of_path(folder, 50)
            ref.isstar_bird(feats, 10)

        for _ in range(max(args.warmup, 0)):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = time.perf_counter() - t0
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    v = per_step * args.steps / dt
    line = {"impl": "reference", "metric": "FID/IS* images/sec", "value": v, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg2: IS* (bird head, T=0.598, 10 splits) + FID moments",
                       "image": "256x256x3 uint8 (PNG files on tmpfs, as the reference reads them)",
                       "images_per_step": per_step,
                       "how": ("the reference's own fid_score._compute_statistics_of_path (Dataset + DataLoader(batch 50, "
                               "num_workers=8) + PIL resize + torchvision Inception3 fp32 + np.mean / np.cov) from oracle/_ref, "
                               "plus the IS* split / KL lines on its pool3" if ref.kind == "reference" else
                               "oracle/restate.py port of the same steps (oracle/_ref absent)")},
            "cpu_baseline": {"value": v, "unit": "images/s", "cores": ref.cores, "kind": ref.kind,
                             "sample": "%d steps x %d images, torch CPU fp32 on %d threads, os.cpu_count()=%s"
                                       % (args.steps, per_step, torch.get_num_threads(), os.cpu_count())},
            "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(json.dumps(line))


# --------------------------------------------------------------------------------------------- GPU arm
# stdout carries exactly ONE line, the JSON result: everything else a library may print there (e.g. the "NCCL version"
# banner on some boxes) is sent to stderr by pointing fd 1 at fd 2 for the lifetime of the process.
_RESULT_FD = os.dup(1)
os.dup2(2, 1)


def _emit(text: str) -> None:
    os.write(_RESULT_FD, (text + "\n").encode())


def ramp_sizes(n, batch, ramp=(256, 768)):
    """Batch sizes of a run over n images: short first batches (while more than twice their size is left), then `batch`."""
    sizes, left = [], int(n)
    for want in ramp:
        if left > 2 * want:
            sizes.append(want)
            left -= want
    while left > 0:
        sizes.append(min(batch, left))
        left -= sizes[-1]
    return sizes


def fixed_size_run(eng, host, n_fixed, rank, world, real_stats, barrier, batch=2048):
    """north_star's target: `n_fixed` images (this rank's contiguous shard of them) from pinned host memory through
    H2D -> resize -> trunk -> moments + IS* sums -> ONE packed all-reduce -> (rank 0) Frechet distance against the
    real-set statistics -> FID and IS* scalars on the host.  Returns wall seconds (max over ranks is taken by the
    caller) and the pieces."""
    from tise_toolbox_b200.pipeline import RealismJob, fid_from_stats, shard_range

    lo, hi = shard_range(n_fixed, rank, world)
    # ramp-up: the first H2D copy cannot hide behind compute, so the first batches are short (256, 768 images) -- with 8
    # ranks pulling from one host at once a 2048-image first copy alone cost ~20 ms of a 95 ms run
    sizes = ramp_sizes(hi - lo, batch)
    offs = [0]
    for n in sizes:
        offs.append(offs[-1] + n)

    def hb(j, n):
        o = offs[j] % max(1, host.shape[0] - n + 1)
        return host[o:o + n]

    job = RealismJob(eng, n_total=n_fixed, start=lo, is_variant="bird", splits=10, input_variant="bird")
    barrier()
    t0 = time.perf_counter()
    for j, n in enumerate(sizes):
        nxt = hb(j + 1, sizes[j + 1]) if j + 1 < len(sizes) else None
        job.step_from_host(hb(j, n), nxt)
    torch.cuda.synchronize()
    t_feat = time.perf_counter()
    res = job.finish()
    torch.cuda.synchronize()
    t_fin = time.perf_counter()
    fid = None
    if rank == 0:
        fid = fid_from_stats(eng, res, real_stats)  # synchronous: the scalar is on the host when it returns
    t_fid = time.perf_counter()
    barrier()
    t1 = time.perf_counter()
    return {"seconds": t1 - t0, "features_s": t_feat - t0, "finish_ms": 1e3 * (t_fin - t_feat),
            "frechet_ms": 1e3 * (t_fid - t_fin), "fid": fid, "is_mean": res.get("is_mean"), "count": res.get("count")}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.child.startswith("cpu_leg:"):
        child_cpu_leg(args.child[len("cpu_leg:"):])
        return
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist

    from tise_toolbox_b200 import fid_score, synth
    from tise_toolbox_b200.engine import Engine, Moments, T_BIRD, warm_collective
    from tise_toolbox_b200.inception import InceptionV3
    from tise_toolbox_b200.pipeline import RealismJob, shard_range

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.workload == "cfg4":
        return run_cfg4(args, rank, world, dev)
    B, K, W = args.batch, args.steps, args.warmup
    n_total = world * B * (K + W)
    lo, _ = shard_range(n_total, rank, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- setup (untimed): weights, resident image pool, calibrated IS head, real-set statistics, warm communicator
    sd = synth.inception_state_dict(0, num_classes=51)
    eng = Engine(sd, device=dev, variant="bird", micro_batch=args.micro_batch, passes=args.passes,
                 single_pass_max_hw=0 if args.strict else None)
    pool = torch.cat([synth.images(1, 0, rank * args.pool + i, min(256, args.pool - i), device=dev)
                      for i in range(0, args.pool, 256)])
    calib = eng.features(pool[:64])
    sd = synth.calibrate_head(sd, calib.cpu(), T_BIRD, class_offset=1)
    eng.load_state_dict(sd)
    real = Moments(eng)  # the "real set" side of the FID (an .npz in the reference, fid_score.py:200-203)
    for i in range(0, 2048, 512):
        real.update(eng.features(synth.images(2, 1, i, 512, device=dev)))
    real_stats = dict(zip(("mu", "sigma"), real.finalize_device()))
    warm_collective(dev)
    eng.frechet(real_stats["mu"], real_stats["sigma"], real_stats["mu"], real_stats["sigma"])  # first-use set-up
    job = RealismJob(eng, n_total=n_total, start=lo, is_variant="bird", splits=10, input_variant="bird")

    def batch_of(step):
        o = (step * B) % args.pool
        if o + B <= args.pool:
            return pool[o:o + B]
        return torch.cat([pool[o:], pool[:o + B - args.pool]])

    for s in range(W):
        job.step(batch_of(s))
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    if sampler:
        sampler.mark_start()
    l0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(K):
        job.step(batch_of(W + s))
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    launches = eng.launches - l0
    if sampler:
        sampler.mark_stop()
    barrier()
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    value = world * B * K / (ms / 1e3)

    # ---- the path's one exchange step + finalisation (not part of the steady-state `value`; part of `fixed` below)
    t0 = time.perf_counter()
    res = job.finish()
    torch.cuda.synchronize()
    finish_ms = 1e3 * (time.perf_counter() - t0)

    # ---- end-to-end: the same step through the public API (RealismJob.step_from_host) from pinned host memory: every
    # step's uint8 batch crosses PCIe (copy stream, double-buffered staging) and every step's result -- the running IS
    # partial sums as they stand after that step -- is read back to the host (pipelined by one step, see e2e_step)
    host = torch.empty((min(args.pool, 2 * B), 256, 256, 3), dtype=torch.uint8).pin_memory()
    host.copy_(pool[:host.shape[0]].cpu())
    job2 = RealismJob(eng, n_total=world * B * (K + W), start=lo, is_variant="bird", splits=10, input_variant="bird")
    d2h = torch.empty(job2.parts.f.numel(), dtype=torch.float64).pin_memory()

    def host_batch(s):
        o = (s * B) % max(1, host.shape[0] - B + 1)
        return host[o:o + B]

    pending = []

    def e2e_step(s):
        # H2D of this step's batch (copy stream, double-buffered), the step, and a D2H read of the step's result (the
        # running IS partial sums).  The read of step s is awaited after step s + 1 has been issued
        # (RealismJob.read_partials_async), so the host never leaves the GPU idle; every step's result is on the host
        # before the timed region ends.
        job2.step_from_host(host_batch(s), host_batch(s + 1))
        pending.append(job2.read_partials_async())
        if len(pending) > 1:
            d2h.copy_(pending.pop(0).wait())

    for s in range(min(W, 2)):
        e2e_step(s)
    while pending:
        d2h.copy_(pending.pop(0).wait())
    barrier()
    e0.record()
    for s in range(K):
        e2e_step(s)
    while pending:
        d2h.copy_(pending.pop(0).wait())
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    barrier()
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * B * K / (float(e2e_ms.item()) / 1e3)
    t_e2e = (time.perf_counter() - float(e2e_ms.item()) / 1e3, time.perf_counter())

    # ---- the resident-input loop once more, AFTER the end-to-end loop: the chip runs at its power cap, so the clock it
    # sustains -- not the H2D traffic -- decides both numbers; this separates "end to end costs x %" from "later in
    # the run the clock is lower"
    barrier()
    t_again0 = time.perf_counter()
    e0.record()
    for s in range(K):
        job2.step(batch_of(s))
    e1.record()
    torch.cuda.synchronize()
    t_again = (t_again0, time.perf_counter())
    again_ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(again_ms, op=dist.ReduceOp.MAX)
    value_again = world * B * K / (float(again_ms.item()) / 1e3)

    # ---- what the PCIe link of this box delivers for the same pinned buffer, GPU otherwise idle (e2e cannot exceed
    # h2d_gbs / bytes-per-image whatever the kernels do)
    torch.cuda.synchronize()
    stg = torch.empty_like(host[:B], device=dev)
    stg.copy_(host[:B], non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        stg.copy_(host[:B], non_blocking=True)
    torch.cuda.synchronize()
    h2d_gbs = 3 * B * IMG_BYTES / (time.perf_counter() - t0) / 1e9
    del stg

    # ---- end-to-end through the reference-facing drop-in function itself: fid_score.calculate_activation_statistics on
    # an iterable of uint8 HOST batches (what a DataLoader(collate_u8) yields), (mu, sigma) returned as numpy arrays
    class HostBatches(list):
        pass

    model = InceptionV3([3], engine=eng)
    kd = max(2, K // 2)
    hb = HostBatches([host_batch(s) for s in range(kd)])
    fid_score.calculate_activation_statistics(HostBatches(hb[:2]), model, batch_size=B, verbose=False)  # warm-up
    barrier()
    t0 = time.perf_counter()
    mu_d, sigma_d = fid_score.calculate_activation_statistics(hb, model, batch_size=B, dims=2048, cuda=True, verbose=False)
    torch.cuda.synchronize()
    dropin_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dropin_s, op=dist.ReduceOp.MAX)
    dropin_value = world * B * kd / float(dropin_s.item())
    eng.set_variant("bird")
    if sampler:
        sampler.close()

    # ---- north_star's target itself: a fixed-size set, first H2D to the scalars on rank 0 (run twice; the second is warm)
    fixed = None
    if args.fixed > 0:
        runs = [fixed_size_run(eng, host, args.fixed, rank, world, real_stats, barrier) for _ in range(2)]
        secs = torch.tensor([r["seconds"] for r in runs], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(secs, op=dist.ReduceOp.MAX)
        fixed = {"images": args.fixed, "seconds": float(secs[1].item()), "first_run_seconds": float(secs[0].item()),
                 "images_per_s": args.fixed / float(secs[1].item()), "features_s": runs[1]["features_s"],
                 "finish_ms": runs[1]["finish_ms"], "frechet_ms": runs[1]["frechet_ms"], "fid": runs[1]["fid"],
                 "is_mean": runs[1]["is_mean"], "count": runs[1]["count"],
                 "what": "wall clock (max over ranks) from the first H2D copy to the FID and IS* scalars on rank 0: "
                         "per-rank shard from pinned host memory in batches of 256, 768, then 2048, one packed all-reduce "
                         "(finish_ms, every rank finalises mu / sigma / IS*), Frechet distance on rank 0 against "
                         "precomputed real-set statistics (frechet_ms)"}

    # ---- roofline of the dominant kernel (conv_umma_kernel): CUDA events around every launch, 2 extra steps
    pk = peaks()
    eng.profile(True)
    for s in range(2):
        job2.step(batch_of(s))
    torch.cuda.synchronize()
    prof = eng.profile_read()
    eng.profile(False)
    conv_ms, conv_launches, conv_flops = prof["conv"]
    trunk_ms = sum(v[0] for v in prof.values())
    achieved = conv_flops / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
    traffic, traffic_src = None, None
    for name in ("ncu_conv_r02.json", "ncu_conv_r01.json"):
        ncu_path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(ncu_path):
            with open(ncu_path) as f:
                nd = json.load(f)
            traffic = nd.get("dram_bytes_per_launch")
            traffic_src = ("profiles/%s: dram__bytes_read.sum + dram__bytes_write.sum averaged over the %s tcgen05 conv "
                           "launches of one %s-image micro-batch" % (name, nd.get("launches"), nd.get("micro_batch", 512)))
            break
    n_mb = 2 * ((B + args.micro_batch - 1) // args.micro_batch)  # micro-batches in the 2 profiled steps
    roofline = {"bound": "tensor",
                "kernel": "conv_umma_kernel + conv_flat_kernel (the tcgen05 convolutions: %d launches per micro-batch)"
                          % (conv_launches // max(1, n_mb)),
                "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"],
                "frac_of_burst_peak": achieved / pk["tflops_burst"], "burst_peak": pk["tflops_burst"],
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": pk["source"],
                "flops_per_launch": conv_flops / max(1, conv_launches),
                "launches_per_step": conv_launches // 2, "avg_launch_us": 1e3 * conv_ms / max(1, conv_launches),
                "algorithmic_gflop_per_image": conv_flops / (2 * B) / 1e9,
                "share_of_trunk_time": conv_ms / trunk_ms if trunk_ms else None,
                "other_kernels_ms_per_step": {k: v[0] / 2 for k, v in prof.items() if k != "conv"},
                "note": "algorithmic FLOPs; default operand policy: one MMA pass per FLOP in conv_umma_kernel and "
                        "conv_flat_kernel (strict policy: two everywhere)",
                "how": "cudaEvent pairs on the launch stream around each launch, summed over 2 steps after the timed region"}

    line = {"metric": "FID/IS* images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": (("f16 x (f16 hi + f16 lo) operands, f32 accumulate" if eng.pass1_max_hw == 0 else
                       "f16 x f16 (W_hi) operands + calibrated mean-field W_lo term, f32 accumulate (Conv2d_1a: "
                       "f16 (hi + lo) input x f16 (hi + lo) weights)") if args.passes == 2 else "f16, f32 accumulate"),
            "data": "synthetic",
            "config": {"workload": ("cfg2: IS* (bird head, T=0.598, 10 splits) + FID moments" if world == 1 else
                                    "cfg3: FID + IS* sharded over %d GPUs, one all-reduce at the end" % world),
                       "image": "256x256x3 uint8", "images_per_step_per_gpu": B, "micro_batch": args.micro_batch,
                       "resident_pool_images": args.pool, "l2": "inputs (%.0f MB/step) larger than L2" % (B * IMG_BYTES / 1e6),
                       "random_init": "seeded He-normal Inception3, calibrated 51-way head",
                       "operand_policy": ("strict: two MMA passes (W_hi, W_lo) in every layer" if eng.pass1_max_hw == 0 else
                                          "default: the tcgen05 implicit-GEMM layers issue the W_hi pass only; W_lo . mean("
                                          "activation) is folded into their shift by a one-off calibration at load time "
                                          "(DESIGN.md section 2; parity at BASELINE sizes: tests/test_gpu_parity_baseline.py)")},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": B * IMG_BYTES,
                    "d2h_bytes_per_step": int(d2h.numel() * 8),
                    "resident_loop_rerun_after_e2e": {"value": value_again, "unit": "images/s",
                                                      "sm_mhz": sampler.median_mhz(*t_again) if sampler else None,
                                                      "sm_mhz_e2e": sampler.median_mhz(*t_e2e) if sampler else None},
                    "h2d_link_gbs": h2d_gbs, "h2d_link_images_per_s": h2d_gbs * 1e9 / IMG_BYTES,
                    "h2d_link_note": "pinned-memory H2D copy of one step's batch with the GPU idle: the ceiling of any "
                                     "end-to-end number that ships raw uint8 images over this link",
                    "through_dropin": {"value": dropin_value, "unit": "images/s", "steps": kd,
                                       "call": "fid_score.calculate_activation_statistics(uint8 host batches, model, "
                                               "batch_size=%d) -> (mu, sigma) numpy: H2D of every batch (prefetched one "
                                               "batch ahead), 33.6 MB D2H of the result" % B}},
            "gpu_launches": int(launches), "roofline": roofline, "fixed": fixed,
            "finish_ms": finish_ms, "is_mean": res.get("is_mean"), "count": res.get("count")}
    if rank == 0:
        line["clocks"] = sampler.summary()
        if world == 1 and not args.no_cpu_baseline:
            import shutil
            import tempfile

            # the CPU leg is the reference's FID module, whose input map is the FID affine: give both sides a head
            # calibrated for THAT feature distribution (the bird-calibrated head makes IS* ~ 1.3 there: a vacuous target)
            eng.set_variant("fid")
            sd_par = synth.calibrate_head(sd, eng.features(pool[:64]).cpu(), T_BIRD, class_offset=1)
            eng.load_state_dict(sd_par)
            tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
            try:
                rec, cpu = cpu_leg_isolated(sd_par, args.cpu_sample, tmp)
            finally:
                shutil.rmtree(tmp, ignore_errors=True)
            line["cpu_baseline"] = rec
            if cpu is not None:
                line["parity"] = gpu_parity(eng, cpu)
                line["parity"]["operand_policy"] = "strict" if eng.pass1_max_hw == 0 else "default (calibrated single pass)"
            else:
                line["parity"] = None
            eng.load_state_dict(sd)
            eng.set_variant("bird")
        if not args.no_frechet:
            line["frechet"] = frechet_latency(eng, res, world == 1 and not args.no_cpu_baseline)
        if world == 1 and not args.no_ingest:
            line["ingest"] = ingest_throughput(eng, pool, pk)
        _emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_cfg4(args, rank, world, dev):
    """BASELINE.json configs[3]: O-IS + O-FID on ragged object crops (h, w log-uniform in [16, 480]; resize-dominated
    small-image path; object_fidelity/O-IS/object_centric_inception_score.py:17-81, O-FID/fid_score.py).  One step = B
    crops through BOTH metrics, as the reference runs them: O-IS (Normalize(.5,.5) input map, 80-way fc with bias,
    T = 2.17, 10 splits) and O-FID (FID input map, pool3 moments) -- two trunk passes per crop.  `value` counts crops."""
    import torch.distributed as dist

    from tise_toolbox_b200 import synth
    from tise_toolbox_b200.engine import Engine, PackedImages, T_OIS
    from tise_toolbox_b200.pipeline import RealismJob, shard_range

    B, K, W = min(args.batch, 4096), args.steps, args.warmup
    U = 2 * B  # distinct crops resident per rank

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sd = synth.inception_state_dict(3, num_classes=80)
    eng = Engine(sd, device=dev, variant="ois", micro_batch=args.micro_batch, passes=args.passes,
                 single_pass_max_hw=0 if args.strict else None)
    sizes = synth.crop_sizes(8, rank * U, U)
    src = synth.images(5, 1, rank * 8, 8, 480, 480, device=dev)
    hw = np.array(sizes, dtype=np.int32)
    nbytes = hw[:, 0].astype(np.int64) * hw[:, 1] * 3
    offs = np.zeros(U, dtype=np.int64)
    offs[1:] = np.cumsum((nbytes + 15) // 16 * 16)[:-1]
    total = int(offs[-1] + nbytes[-1])
    buf = torch.zeros(total + 16, dtype=torch.uint8, device=dev)
    for i, (h, w) in enumerate(sizes):  # crop i = a window of a 480 x 480 synthetic image (crop_object.py cuts boxes like this)
        y0, x0 = (i * 37) % (481 - h), (i * 101) % (481 - w)
        buf[offs[i]:offs[i] + h * w * 3] = src[i % 8, y0:y0 + h, x0:x0 + w].reshape(-1)
    ends = offs + nbytes

    def pack_of(step):
        i0 = (step * B) % U
        return PackedImages(buf, offs[i0:i0 + B], hw[i0:i0 + B]), i0

    calib = eng.features(pack_of(0)[0])[:256]
    sd = synth.calibrate_head(sd, calib.cpu(), T_OIS, seed=80)
    eng.load_state_dict(sd)
    n_total = world * B * (K + W)
    lo, _ = shard_range(n_total, rank, world)

    def jobs():
        return (RealismJob(eng, n_total=n_total, start=lo, is_variant="ois", splits=10, input_variant="ois", want_moments=False),
                RealismJob(eng, n_total=n_total, start=lo, input_variant="fid"))

    j_ois, j_fid = jobs()

    def step(s):
        p, _ = pack_of(s)
        j_ois.step(p)
        j_fid.step(p)

    for s in range(W):
        step(s)
    sampler = ClockSampler(dev.index) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    if sampler:
        sampler.mark_start()
    l0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(K):
        step(W + s)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    launches = eng.launches - l0
    if sampler:
        sampler.mark_stop()
    barrier()
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    value = world * B * K / (ms / 1e3)
    r_ois, r_fid = j_ois.finish(), j_fid.finish()

    # ---- the resize kernel alone on the same ragged batches (HBM-bound: reads the crop, writes 299 x 299 x 3)
    out = torch.empty((B, 299, 299, 3), dtype=torch.uint8, device=dev)
    eng.resize(pack_of(0)[0], out=out)
    torch.cuda.synchronize()
    e0.record()
    rs_bytes = 0
    for s in range(K):
        p, i0 = pack_of(s)
        eng.resize(p, out=out)
        rs_bytes += int(nbytes[i0:i0 + B].sum()) + B * 299 * 299 * 3
    e1.record()
    torch.cuda.synchronize()
    rs_ms = e0.elapsed_time(e1)
    pk = peaks()
    rs_gbs = rs_bytes / (rs_ms / 1e3) / 1e9

    # ---- end to end: the step's crops cross PCIe from pinned host memory (contiguous byte range of the packed pool),
    # the O-IS partial sums are read back every step
    host = torch.empty(buf.numel(), dtype=torch.uint8).pin_memory()
    host.copy_(buf.cpu())
    stage = [torch.empty(int((ends[i0 + B - 1] - offs[i0] + 15) // 16 * 16) + 16, dtype=torch.uint8, device=dev) for i0 in (0, B)]
    copy_stream = torch.cuda.Stream(device=dev)
    j_ois, j_fid = jobs()
    d2h = torch.empty(j_ois.parts.f.numel(), dtype=torch.float64).pin_memory()
    main = torch.cuda.current_stream(dev)
    done = [None, None]
    h2d_bytes = 0

    def upload(s):
        k = s % 2
        i0 = (s * B) % U
        a, b = int(offs[i0]), int(ends[i0 + B - 1])
        if done[k] is not None:
            copy_stream.wait_event(done[k])
        with torch.cuda.stream(copy_stream):
            stage[k][:b - a].copy_(host[a:b], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev, b - a

    def e2e_loop(n_steps):
        nonlocal h2d_bytes
        pend = []
        ev, nb = upload(0)
        for s in range(n_steps):
            nxt = upload(s + 1) if s + 1 < n_steps else None
            main.wait_event(ev)
            k, i0 = s % 2, (s * B) % U
            p = PackedImages(stage[k], offs[i0:i0 + B] - offs[i0], hw[i0:i0 + B])
            j_ois.step(p)
            j_fid.step(p)
            done[k] = torch.cuda.Event()
            done[k].record(main)
            h2d_bytes += nb
            pend.append(j_ois.read_partials_async())
            if len(pend) > 1:
                d2h.copy_(pend.pop(0).wait())
            if nxt:
                ev, nb = nxt
        while pend:
            d2h.copy_(pend.pop(0).wait())

    e2e_loop(2)
    barrier()
    h2d_bytes = 0
    e0.record()
    e2e_loop(K)
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * B * K / (float(e2e_ms.item()) / 1e3)
    if sampler:
        sampler.close()

    line = {"metric": "O-IS + O-FID crops/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8 resize (integer, bit-exact) + f16 x (f16 hi + f16 lo) trunk, f32 accumulate", "data": "synthetic",
            "config": {"workload": "cfg4: O-IS (80-way head, T=2.17, 10 splits) + O-FID moments on ragged object crops",
                       "crops": "h, w log-uniform in [16, 480], %d distinct crops resident per GPU (%.0f MB > L2), "
                                "mean %.0f x %.0f" % (U, total / 1e6, hw[:, 0].mean(), hw[:, 1].mean()),
                       "crops_per_step_per_gpu": B, "crops_total": world * B * K, "micro_batch": args.micro_batch,
                       "trunk_passes_per_crop": 2},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d_bytes // max(K, 1),
                    "d2h_bytes_per_step": int(d2h.numel() * 8)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "resize_fused_kernel (ragged crops -> 299 x 299 x 3, Pillow-exact)",
                         "achieved": rs_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": rs_gbs / pk["hbm_gbs"],
                         "traffic": None, "bytes_per_launch": rs_bytes / K, "avg_launch_us": 1e3 * rs_ms / K,
                         "share_of_step": rs_ms / ms, "fused_share": 1.0,
                         "note": "every crop <= 480 x 480 takes the fused single-kernel path (the uint8 temporary between "
                                 "Pillow's two passes stays in shared memory); the resize_h/v fallback is only used "
                                 "when a strip's input rows exceed 96 KB of shared memory",
                         "peak_source": pk["source"]},
            "ois_mean": r_ois.get("is_mean"), "count": r_fid.get("count")}
    if rank == 0:
        line["clocks"] = sampler.summary()
        if world == 1 and not args.no_cpu_baseline:
            line.update(cfg4_cpu_leg(eng, sd, buf, offs, hw, min(args.cpu_sample, 400)))
        _emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cfg4_cpu_leg(eng, sd, buf, offs, hw, n):
    """The reference's O-IS function (exec'd from its own source) and its O-FID statistics path on `n` of the same
    crops written as PNG files; O-IS / pool3 deltas of the GPU path against it."""
    import shutil
    import tempfile

    from oracle import harness, reference_shim, restate
    from tise_toolbox_b200 import object_centric_inception_score as ois
    from tise_toolbox_b200.engine import PackedImages
    from tise_toolbox_b200.pipeline import RealismJob

    n = max(64, n // 32 * 32)
    crops = [buf[int(offs[i]):int(offs[i]) + int(hw[i, 0]) * int(hw[i, 1]) * 3].view(int(hw[i, 0]), int(hw[i, 1]), 3).cpu().numpy()
             for i in range(n)]
    tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        from PIL import Image

        for i, c in enumerate(crops):
            Image.fromarray(c).save(os.path.join(tmp, "img_%05d.png" % i), compress_level=1)
        cores = harness.use_all_cores()
        kind = "reference" if reference_shim.available() else "port"
        t0 = time.perf_counter()
        if kind == "reference":
            run, ds_cls = reference_shim.load_ois_function(sd, 80)
            ds = ds_cls(tmp)
            names = list(ds.namelist) if hasattr(ds, "namelist") else sorted(os.listdir(tmp))
            ois_ref = run(ds, batch_size=32, splits=10)
        else:
            names = os.listdir(tmp)
            ims = [np.asarray(Image.open(os.path.join(tmp, f)).convert("RGB")) for f in names]
            trunk = restate.Trunk(sd, num_classes=80)
            ois_ref = restate.ois_inception_score(restate.ois_preds(trunk.logits(harness.pil_preprocess(ims, "ois"))), 10)
        t_ois = time.perf_counter() - t0
        ref = ReferenceCpu(sd, 80)
        t0 = time.perf_counter()
        m, s, feats = ref.statistics_of_path(tmp, 32)
        t_fid = time.perf_counter() - t0
        order_fid = ref.file_order(tmp)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    order_ois = [int(f[4:9]) for f in names]
    # GPU side, same crops in the same orders
    def pack(order):
        o = np.array(order)
        return PackedImages(buf, offs[o], hw[o])

    j = RealismJob(eng, n_total=n, is_variant="ois", splits=10, input_variant="ois", want_moments=False)
    j.step(pack(order_ois))
    got_ois = j.finish()["is_mean"]
    j = RealismJob(eng, n_total=len(order_fid) // 32 * 32, input_variant="fid")
    f = j.step(pack(order_fid[:len(order_fid) // 32 * 32])).cpu().numpy()
    rf = feats.numpy()
    return {"cpu_baseline": {"value": n / (t_ois + t_fid), "unit": "images/s", "cores": cores, "kind": kind,
                             "sample": "%d of the same crops as PNG files: the reference's O-IS inception_score (batch 32, "
                                       "%.1f s) + O-FID _compute_statistics_of_path (batch 32, %.1f s)" % (n, t_ois, t_fid)},
            "parity": {"ois_rel": abs(got_ois - ois_ref[0]) / ois_ref[0], "ois": got_ois, "ois_cpu": ois_ref[0],
                       "pool3_rel_l2": float(np.linalg.norm(f - rf) / np.linalg.norm(rf)), "images": n}}


def ingest_throughput(eng, pool, pk, n=4096, unique=128):
    """PNG ingest: `unique` images of the resident pool are PNG-encoded by Pillow on the host
    (default compression level), replicated to n files, and decoded on the device.  `device_images_per_s`: files
    already in HBM, CUDA events around tise_png_decode (3 kernels); `e2e_images_per_s`: Engine.decode_images from host
    bytes (packing into pinned memory, H2D, decode, status read-back).  Pillow decode of a sample on one host core is
    timed beside it.  The inflate kernel is latency / instruction bound (one warp walks one Huffman stream), so its
    HBM fraction is reported for the record, not as a target."""
    import io

    from PIL import Image

    imgs = pool[:unique].cpu().numpy()
    blobs = []
    for im in imgs:
        b = io.BytesIO()
        Image.fromarray(im).save(b, "PNG")
        blobs.append(b.getvalue())
    batch = [blobs[i % unique] for i in range(n)]
    packed = eng.decode_images(batch)  # warm-up + parity
    ok = bool((packed[n - 1] == pool[(n - 1) % unique]).all().item())
    staged = eng.stage_blobs(batch)
    d_files = staged["pinned"][:staged["total"]].to(eng.device)
    index = np.arange(n)
    out = torch.empty_like(packed.buf)
    off = packed.offsets
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        status = eng.decode_staged_png(staged, d_files, index, out, off)
    e1.record()
    torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1) / reps
    ok = ok and not bool(status.any().item())
    t0 = time.perf_counter()
    for _ in range(3):
        eng.decode_images(batch)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / 3
    # the CLI's path: files on disk (tmpfs) -> DeviceImageLoader (readinto pinned memory on 8 threads, H2D, decode),
    # consumed in the reference's batches of 50
    import shutil
    import tempfile

    from tise_toolbox_b200.img_data import DeviceImageLoader

    shm_ok = os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 2 * n * 200000
    tmp = tempfile.mkdtemp(dir="/dev/shm" if shm_ok else None)
    seen, files_s = 0, float("nan")
    try:
        names = []
        for i, b in enumerate(batch):
            names.append(os.path.join(tmp, "img_%05d.png" % i))
            with open(names[-1], "wb") as f:
                f.write(b)
        seen = 0
        t0 = time.perf_counter()
        for pack in DeviceImageLoader(names, eng, batch_size=50, drop_last=True):
            seen += len(pack)
        torch.cuda.synchronize()
        files_s = time.perf_counter() - t0
    except OSError as e:  # (no space for the 620 MB of files: the rest of the line is still valid)
        print("ingest: from-files measurement skipped: %s" % e, file=sys.stderr)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    t0 = time.perf_counter()
    for b in blobs[:64]:
        np.asarray(Image.open(io.BytesIO(b)).convert("RGB"))
    pil = 64 / (time.perf_counter() - t0)
    # the same measurement for baseline JPEG files (quality 90, 4:2:0 -- Pillow's defaults; MS-COCO-like real sets)
    jblobs = []
    for im in imgs:
        b = io.BytesIO()
        Image.fromarray(im).save(b, "JPEG", quality=90)
        jblobs.append(b.getvalue())
    jbatch = [jblobs[i % unique] for i in range(n)]
    jpacked = eng.decode_images(jbatch, host_fallback=False)
    jref = np.asarray(Image.open(io.BytesIO(jblobs[(n - 1) % unique])).convert("RGB"))
    jok = bool(np.array_equal(jpacked[n - 1].cpu().numpy(), jref))
    jstaged = eng.stage_blobs(jbatch)
    jd_files = jstaged["pinned"][:jstaged["total"]].to(eng.device)
    jout = torch.empty_like(jpacked.buf)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        jstatus = eng.decode_staged_png(jstaged, jd_files, index, jout, jpacked.offsets, jpeg=True)
    e1.record()
    torch.cuda.synchronize()
    jdev_ms = e0.elapsed_time(e1) / reps
    jok = jok and not bool(jstatus.any().item())
    t0 = time.perf_counter()
    for b in jblobs[:64]:
        np.asarray(Image.open(io.BytesIO(b)).convert("RGB"))
    jpil = 64 / (time.perf_counter() - t0)
    jpeg = {"files": n, "image": "256x256x3 baseline JPEG, quality 90, 4:2:0", "file_bytes_mean": float(np.mean([len(b) for b in jblobs])),
            "device_images_per_s": n / (jdev_ms / 1e3), "device_ms": jdev_ms, "pil_images_per_s_one_core": jpil,
            "bit_exact_vs_pillow": jok, "kernels_per_call": 3}
    file_bytes = float(np.mean([len(b) for b in blobs]))
    algo_bytes = n * (2 * file_bytes + 2 * (IMG_BYTES + 256) + 2 * IMG_BYTES)  # gather r+w, inflate w + unfilter r, unfilter w (+ in)
    return {"files": n, "unique": unique, "image": "256x256x3 PNG, Pillow default compression", "file_bytes_mean": file_bytes,
            "device_images_per_s": n / (dev_ms / 1e3), "device_ms": dev_ms, "e2e_images_per_s": n / e2e_s,
            "from_files_images_per_s": (seen / files_s) if seen else None,
            "pil_images_per_s_one_core": pil, "host_cores": os.cpu_count(), "bit_exact": ok,
            "hbm_frac": algo_bytes / (dev_ms / 1e3) / 1e9 / pk["hbm_gbs"], "kernels_per_call": 3, "jpeg": jpeg}


def frechet_latency(eng, res, with_cpu):
    """Frechet 'sqrtm' latency at d = 2048 on this run's own statistics against a perturbed copy."""
    mu, sigma = res["mu"], res["sigma"]
    g = torch.Generator(device="cpu").manual_seed(3)
    d = mu.shape[0]
    q = torch.randn(d, 64, generator=g, dtype=torch.float64).to(mu.device) * 0.02
    sigma2 = sigma * 1.1 + q @ q.t()
    mu2 = mu + 0.05
    eng.frechet(mu, sigma, mu2, sigma2)  # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    v, _ = eng.frechet(mu, sigma, mu2, sigma2)
    gpu_ms = 1e3 * (time.perf_counter() - t0)
    out = {"d": int(d), "gpu_ms": gpu_ms, "value": v}
    if with_cpu:
        from oracle import restate
        a = [t.cpu().numpy() for t in (mu, sigma, mu2, sigma2)]
        t0 = time.perf_counter()
        ref = float(restate.calculate_frechet_distance(*a))
        out["cpu_scipy_sqrtm_ms"] = 1e3 * (time.perf_counter() - t0)
        out["rel_diff_vs_scipy"] = abs(v - ref) / abs(ref)
    return out


if __name__ == "__main__":
    main()
