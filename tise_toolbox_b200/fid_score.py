#!/usr/bin/env python3
"""Drop-in for image_realism/FID/fid_score.py and object_fidelity/O-FID/fid_score.py on B200.

Same functions, arguments, error behaviour, CLI flags and result-file format as the reference;
underneath, feature extraction, the activation statistics and the Frechet distance run as CUDA
kernels of libtise_b200.so.  ``--ofid`` switches the result prefix to ``O-FID:``
(object_fidelity/O-FID/fid_score.py:219-222); the network weights decide which metric it is.
"""
import os
import warnings
from argparse import ArgumentDefaultsHelpFormatter, ArgumentParser

import numpy as np
import torch
import torch.utils.data

from . import img_data
from .engine import Moments, PackedImages, default_engine
from .inception import InceptionV3

warnings.filterwarnings("ignore")

parser = ArgumentParser(formatter_class=ArgumentDefaultsHelpFormatter)
parser.add_argument("--batch-size", type=int, default=64, help="Batch size to use")
parser.add_argument("--dims", type=int, default=2048, choices=list(InceptionV3.BLOCK_INDEX_BY_DIM),
                    help=("Dimensionality of Inception features to use. " "By default, uses pool3 features"))
parser.add_argument("-c", "--gpu", default="", type=str, help="GPU to use (leave blank for CPU only)")
parser.add_argument("--path1", type=str, default=64)
parser.add_argument("--path2", type=str, default=64)
parser.add_argument("--saved_file", type=str, default="")
parser.add_argument("--weights", type=str, default=None, help="torchvision Inception3 state_dict (.pth)")
parser.add_argument("--ofid", action="store_true", help="write 'O-FID:' instead of 'FID:'")


def _batch_features(batch, model, dims):
    """One batch -> device float32 [B, dims].  uint8 HWC batches (tensor or ragged list) take the fused
    GPU path; float Bx3xHxW batches in [0,1] take the reference's ``model(batch)[0]`` contract."""
    if isinstance(batch, (list, tuple, PackedImages)) or batch.dtype == torch.uint8:
        return model.features_u8(batch, dims=dims)
    pred = model(batch)[0]
    return pred.reshape(pred.shape[0], -1)


def get_activations(images, model, batch_size=64, dims=2048, cuda=False, verbose=True):
    """fid_score.py:67-118.  ``images`` is an iterable of batches with ``__len__`` = number of batches
    (a DataLoader with drop_last=True); returns (len(images) * batch_size, dims) float64."""
    model.eval()
    d0 = images.__len__() * batch_size
    if batch_size > d0:
        print(("Warning: batch size is bigger than the data size. " "Setting batch size to data size"))
        batch_size = d0
    n_batches = d0 // batch_size if batch_size else 0
    n_used_imgs = n_batches * batch_size
    pred_arr = np.empty((n_used_imgs, dims))
    for i, batch in enumerate(images):
        start = i * batch_size
        end = start + batch_size
        pred_arr[start:end] = _batch_features(batch, model, dims).cpu().numpy().reshape(batch_size, -1)
    if verbose:
        print(" done")
    return pred_arr


def calculate_frechet_distance(mu1, sigma1, mu2, sigma2, eps=1e-6):
    """fid_score.py:121-171: d^2 = |mu1 - mu2|^2 + Tr(C1 + C2 - 2 sqrt(C1 C2)), evaluated in fp64 on the GPU
    (libtise_b200 ``tise_frechet``).  Same asserts; the eps retry and the ValueError on a large imaginary
    component are reported by the kernel's flags."""
    mu1 = np.atleast_1d(mu1)
    mu2 = np.atleast_1d(mu2)
    sigma1 = np.atleast_2d(sigma1)
    sigma2 = np.atleast_2d(sigma2)
    assert mu1.shape == mu2.shape, "Training and test mean vectors have different lengths"
    assert sigma1.shape == sigma2.shape, "Training and test covariances have different dimensions"
    value, flags = default_engine().frechet(mu1, sigma1, mu2, sigma2, eps)
    if flags & 1:
        print(("fid calculation produces singular product; " "adding %s to diagonal of cov estimates") % eps)
    if flags & 2:
        raise ValueError("Imaginary component: sqrt(sigma1 sigma2) has eigenvalues with |Im| > 1e-3")
    return np.float64(value)


def _prefetched(images, device, engine=None):
    """Iterates ``images`` one batch ahead: host tensors (uint8 HWC batches or the reference's float batches) are
    uploaded on a private copy stream while the previous batch is still in the trunk, so the PCIe transfer hides
    behind the compute instead of serialising with it.  Ragged lists / device-resident packs pass through untouched.

    With ``engine`` (pinned host batches): the next batch is uploaded in one chunk per micro-batch of the batch being
    consumed, each chunk released by the trunk's progress event (``Engine.set_progress_event``), so the DMA traffic runs
    beside the compute-bound layers instead of beside the HBM-bound front of the network -- an unpaced transfer costs the
    trunk 3-5 % (tools/e2e_probe.py)."""
    copy_stream = torch.cuda.Stream(device=device)
    main = torch.cuda.current_stream(device)
    progress = None
    bufs, released, n_paced, cur_slot = [None, None], [None, None], 0, None

    def is_host(batch):
        return torch.is_tensor(batch) and not batch.is_cuda

    def upload(batch):  # whole batch at once (the first batch; unpinned or engine-less callers)
        if is_host(batch):
            with torch.cuda.stream(copy_stream):
                dev = batch.to(device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return dev, ev
        return batch, None

    it = iter(images)
    try:
        cur = upload(next(it))
    except StopIteration:
        return
    while cur is not None:
        try:
            nxt_host = next(it)
        except StopIteration:
            nxt_host = None
        batch, ev = cur
        if ev is not None:
            main.wait_event(ev)
            if cur_slot is None:
                batch.record_stream(main)  # allocated on the copy stream, consumed on the compute stream
        paced = engine is not None and nxt_host is not None and is_host(nxt_host) and nxt_host.is_pinned() and len(nxt_host) > 0
        if paced:
            progress = engine.progress_event()
            # two persistent device buffers used alternately (a fresh 1 GB allocation per batch cost more than the pacing
            # gains); a buffer is overwritten only after the consumer's work on the batch it held two steps ago
            slot = n_paced & 1
            n_paced += 1
            nbytes = nxt_host.numel() * nxt_host.element_size()
            if bufs[slot] is None or bufs[slot].numel() < nbytes:
                with torch.cuda.stream(copy_stream):
                    bufs[slot] = torch.empty(nbytes, dtype=torch.uint8, device=device)  # (copy stream's pool)
            if released[slot] is not None:
                copy_stream.wait_event(released[slot])
            nxt_dev = bufs[slot][:nbytes].view(nxt_host.dtype).view(nxt_host.shape)
            total, state = int(nxt_host.shape[0]), {"done": 0}

            def hook(k, count, nxt_host=nxt_host, nxt_dev=nxt_dev, total=total, state=state):
                hi = total if k == count - 1 else (total * (k + 1)) // count
                lo = state["done"]
                if hi > lo:
                    copy_stream.wait_event(progress)  # the record of micro-batch k has just been issued
                    with torch.cuda.stream(copy_stream):
                        nxt_dev[lo:hi].copy_(nxt_host[lo:hi], non_blocking=True)
                    state["done"] = hi

            engine.on_micro_batch = hook
            try:
                yield batch  # the consumer runs the trunk on `batch`: the hook fires once per micro-batch
            finally:
                engine.on_micro_batch = None
            if cur_slot is not None:  # the consumer's work on the buffer `batch` lives in has been enqueued
                released[cur_slot] = torch.cuda.Event()
                released[cur_slot].record(main)
            cur_slot = slot
            with torch.cuda.stream(copy_stream):
                if state["done"] < total:  # (the consumer did not run the trunk, or ran fewer micro-batches)
                    nxt_dev[state["done"]:].copy_(nxt_host[state["done"]:], non_blocking=True)
                ev2 = torch.cuda.Event()
                ev2.record(copy_stream)
            cur = (nxt_dev, ev2)
        else:
            nxt = upload(nxt_host) if nxt_host is not None else None
            yield batch
            if cur_slot is not None:
                released[cur_slot] = torch.cuda.Event()
                released[cur_slot].record(main)
                cur_slot = None
            cur = nxt


def _accumulate_moments(images, model, batch_size, dims, next_loader=None):
    model.eval()
    mom = Moments(model.engine, dims)
    if hasattr(images, "groups") and next_loader is not None:
        for pack in images.groups(next_loader=next_loader):
            mom.update(_batch_features(pack, model, dims))
        return mom
    if hasattr(images, "groups"):
        # DeviceImageLoader: the sums do not depend on how the set is cut into batches, so whole decode groups (a whole
        # number of drop_last batches each) go through the trunk at once instead of batch_size images per call
        for pack in images.groups():
            mom.update(_batch_features(pack, model, dims))
        return mom
    for batch in _prefetched(images, model.engine.device, model.engine):
        if len(batch) != batch_size:
            raise ValueError("calculate_activation_statistics: batch of %d, expected %d (drop_last=True, "
                             "fid_score.py:216)" % (len(batch), batch_size))
        mom.update(_batch_features(batch, model, dims))
    return mom


def calculate_activation_statistics(images, model, batch_size=64, dims=2048, cuda=False, verbose=True):
    """fid_score.py:174-196, one pass: the batches' sum(x) and sum(x x^T) accumulate in fp64 on the device
    and only (mu, sigma) come back -- the N x dims activation matrix is never materialised."""
    mom = _accumulate_moments(images, model, batch_size, dims)
    if verbose:
        print(" done")
    return mom.finalize()


def _world():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def _folder_loader(path, model, batch_size):
    """DeviceImageLoader over this rank's share of a folder (None for .npz / .shard paths or TISE_GPU_DECODE=0)."""
    if path.endswith(".npz") or (path.endswith(".shard") and os.path.isfile(path)) or \
            os.environ.get("TISE_GPU_DECODE", "1") == "0":
        return None
    dataset = img_data.Dataset(path, raw_u8=True)
    rank, world = _world()
    if world > 1:
        # launched under torchrun: every rank takes a contiguous run of whole batches of the folder and the packed
        # moments are summed with ONE all-reduce; each rank ends up with the same (mu, sigma)
        from .pipeline import shard_batches

        lo, hi = shard_batches(len(dataset.file_names), batch_size, rank, world)
        dataset.file_names = dataset.file_names[lo:hi]
    return img_data.DeviceImageLoader(dataset.file_names, model.engine, batch_size, drop_last=True)


def _compute_statistics_of_path(path, model, batch_size, dims, cuda, loader=None, next_loader=None):
    """fid_score.py:199-220.  ``loader`` / ``next_loader`` (optional, folders only): this path's DeviceImageLoader built
    by the caller, and the loader of the path that follows -- its first group is read while this path's last group is
    still on the GPU (calculate_fid_given_paths)."""
    if path.endswith(".npz"):
        f = np.load(path, allow_pickle=True)  # fid_score.py:200-203
        m, s = f["mu"][:], f["sigma"][:]
        f.close()
    elif path.endswith(".shard") and os.path.isfile(path):
        # raw uint8 shard (tise_toolbox_b200/shards.py): file -> pinned memory -> HBM, no codec anywhere
        from .shards import ShardLoader

        rank, world = _world()
        mom = _accumulate_moments(ShardLoader(path, model.engine, batch_size, drop_last=True, rank=rank, world=world),
                                  model, batch_size, dims)
        if world > 1:
            mom.allreduce()
        print(" done")
        m, s = mom.finalize()
    else:
        # fid_score.py:206-217 with decode, Resize and ToTensor moved onto the GPU: host threads only read bytes.
        # TISE_GPU_DECODE=0 keeps PIL decode on DataLoader workers (resize / ToTensor still on the GPU).
        rank, world = _world()
        dataloader = loader if loader is not None else _folder_loader(path, model, batch_size)
        if dataloader is None:  # TISE_GPU_DECODE=0
            dataset = img_data.Dataset(path, raw_u8=True)
            if world > 1:
                from .pipeline import shard_batches

                lo, hi = shard_batches(len(dataset.file_names), batch_size, rank, world)
                dataset.file_names = dataset.file_names[lo:hi]
            dataloader = torch.utils.data.DataLoader(dataset=dataset, batch_size=batch_size, shuffle=False,
                                                     drop_last=True, num_workers=8, collate_fn=img_data.collate_u8)
            next_loader = None
        mom = _accumulate_moments(dataloader, model, batch_size, dims, next_loader=next_loader)
        if world > 1:
            mom.allreduce()
        print(" done")
        m, s = mom.finalize()
    return m, s


def calculate_fid_given_paths(paths, batch_size, cuda, dims, model=None):
    """Calculates the FID of two paths (fid_score.py:223-238)."""
    for p in paths:
        if not os.path.exists(p):
            raise RuntimeError("Invalid path: %s" % p)
    block_idx = InceptionV3.BLOCK_INDEX_BY_DIM[dims]
    if model is None:
        model = InceptionV3([block_idx])
    # both folders are one stream of decode groups: the second set's first files are read while the first set's last
    # group is on the GPU
    loaders = [_folder_loader(p, model, batch_size) for p in paths]
    m1, s1 = _compute_statistics_of_path(paths[0], model, batch_size, dims, cuda, loader=loaders[0], next_loader=loaders[1])
    m2, s2 = _compute_statistics_of_path(paths[1], model, batch_size, dims, cuda, loader=loaders[1])
    return calculate_frechet_distance(m1, s1, m2, s2)


def save_statistics(path, mu, sigma):
    """Writes the .npz cache _compute_statistics_of_path reads (keys ``mu``, ``sigma``)."""
    np.savez(path, mu=mu, sigma=sigma)


def main(argv=None):
    args = parser.parse_args(argv)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:  # torchrun --nproc-per-node N -m tise_toolbox_b200.fid_score ...: one process per GPU
        import torch.distributed as dist

        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        from .engine import warm_collective

        warm_collective(torch.device("cuda", local))  # the end-of-run exchange then costs what the wire costs
    elif args.gpu != "":
        os.environ["CUDA_VISIBLE_DEVICES"] = args.gpu
    if args.weights:
        os.environ["TISE_INCEPTION_WEIGHTS"] = args.weights
    paths = [args.path1, args.path2]
    rank = int(os.environ.get("RANK", "0"))
    if rank == 0:
        print(paths)
    fid_value = calculate_fid_given_paths(paths, args.batch_size, args.gpu, args.dims).item()
    tag = "O-FID" if args.ofid else "FID"
    if rank == 0:
        if args.saved_file:
            with open(args.saved_file, "w") as f:
                f.write(f"{tag}: {fid_value}")
        print(f"{tag}: {fid_value}")
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()
    return fid_value


if __name__ == "__main__":
    main()
