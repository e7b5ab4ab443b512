"""-m gpu: the path at BASELINE.json's FULL sizes, checked through size-independent properties (the CPU oracle would
need hours at these sizes): additivity of the statistics over shards (what the multi-GPU all-reduce relies on),
FID(X, X) = 0 (the reference's only transferable known answer, object_fidelity/results/O-FID/real_images.txt),
split bookkeeping, and agreement of repeated inputs across the kernels' internal chunk boundaries."""
import numpy as np
import pytest
import torch

from oracle import restate
from tise_toolbox_b200 import synth
from tise_toolbox_b200.engine import T_BIRD, Engine, ISPartials, Moments, PackedImages
from tise_toolbox_b200.pipeline import RealismJob, fid_from_stats, shard_range

pytestmark = pytest.mark.gpu


def _run(eng, n_total, lo, hi, pool, batch=1000):
    job = RealismJob(eng, n_total=n_total, start=lo, is_variant="bird", splits=10, input_variant="bird")
    for s in range(lo, hi, batch):
        e = min(hi, s + batch)
        idx = torch.arange(s, e, device=pool.device) % pool.shape[0]
        job.step(pool[idx])
    return job


def test_cfg2_cfg3_30k_images_sharded_equals_single(state_dict):
    """cfg2 / cfg3: IS* (10 splits) + FID moments over 30 000 256 x 256 images; 8 shards summed == one pass."""
    n = 30000
    sd = synth.inception_state_dict(0, num_classes=51)
    eng = Engine(sd, device="cuda:0", variant="bird", micro_batch=500)
    pool = torch.cat([synth.images(1, 0, i, 250, device="cuda") for i in range(0, 1500, 250)])  # image i = pool[i % 1500]
    calib = eng.features(pool[:64])
    sd = synth.calibrate_head(sd, calib.cpu(), T_BIRD, class_offset=1)
    eng.load_state_dict(sd)
    single = _run(eng, n, 0, n, pool)
    res = single.finish()
    assert res["count"] == n
    assert np.all(single.parts.i.cpu().numpy()[:10] == n // 10)  # every split holds exactly N / 10 images
    assert np.isfinite(res["is_mean"]) and 1.5 < res["is_mean"] < 51
    # 8 "ranks" run one after the other on this GPU; summing their packed buffers is what all_reduce(SUM) does
    world = 8
    mom = Moments(eng, 2048)
    parts = ISPartials(eng, 50, 10)
    for r in range(world):
        lo, hi = shard_range(n, r, world)
        job = _run(eng, n, lo, hi, pool)
        job.moments.buf[0] = float(job.moments.count)
        mom.buf += job.moments.buf
        parts.f += job.parts.f
        parts.i += job.parts.i
    mom.count = int(round(float(mom.buf[0].item())))
    assert mom.count == n
    mu, sigma = mom.finalize_device()
    assert float((mu - res["mu"]).abs().max()) <= 1e-12 * float(res["mu"].abs().max())
    assert float((sigma - res["sigma"]).abs().max()) <= 1e-10 * float(res["sigma"].abs().max())
    m8, s8, scores8 = parts.finalize(nan_on_zero=True)
    assert abs(m8 - res["is_mean"]) <= 1e-12 * res["is_mean"] and abs(s8 - res["is_std"]) <= 1e-9
    # FID(X, X) = 0 at full size, and FID against a shifted set is positive and symmetric
    assert abs(fid_from_stats(eng, res, res)) <= 1e-6 * float(torch.trace(res["sigma"]))
    other = {"mu": res["mu"] + 0.01, "sigma": res["sigma"] * 1.05}
    ab, ba = fid_from_stats(eng, res, other), fid_from_stats(eng, other, res)
    # (the pool has 1500 distinct images: rank-deficient covariances, so the two orders differ by the O(sqrt(delta))
    # effect of the Cholesky jitter on the null space -- bounded like FID(X, X) above)
    assert ab > 0 and abs(ab - ba) <= 1e-6 * float(torch.trace(res["sigma"]))


def test_cfg5_200k_feature_statistics_additive_and_exact(engine):
    """cfg5: the statistics path on 200 000 x 2048 features: additivity over batches of any size and agreement with an
    independent fp64 evaluation (np.cov on a 20 000-row prefix, torch fp64 GEMM for the whole set)."""
    n, d = 200000, 2048
    g = torch.Generator(device="cuda").manual_seed(5)
    basis = torch.randn(64, d, generator=g, device="cuda") * 0.3
    mom_a, mom_b = Moments(engine, d), Moments(engine, d)
    sum64 = torch.zeros(d, dtype=torch.float64, device="cuda")
    sxx64 = torch.zeros(d, d, dtype=torch.float64, device="cuda")
    prefix = []
    for i0 in range(0, n, 20000):
        x = (torch.randn(20000, 64, generator=g, device="cuda") @ basis + torch.randn(20000, d, generator=g, device="cuda") * 0.1
             + 1.7).float().contiguous()
        mom_a.update(x)  # one call per 20 000 rows
        for j0 in range(0, 20000, 3333):  # ragged batches
            mom_b.update(x[j0:j0 + 3333])
        xd = x.double()
        sum64 += xd.sum(0)
        sxx64 += xd.t() @ xd
        if i0 == 0:
            prefix = x.cpu().numpy().astype(np.float64)
            m0, s0 = Moments(engine, d), None
            m0.update(x)
            mu0, sig0 = m0.finalize()
            rmu, rsig = restate.calculate_activation_statistics(prefix)
            assert np.allclose(mu0, rmu, rtol=1e-12) and np.abs(sig0 - rsig).max() <= 1e-10 * np.abs(rsig).max()
    mu_a, sig_a = mom_a.finalize_device()
    mu_b, sig_b = mom_b.finalize_device()
    assert mom_a.count == mom_b.count == n
    assert float((mu_a - mu_b).abs().max()) <= 1e-13 and float((sig_a - sig_b).abs().max()) <= 1e-11
    mu_ref = sum64 / n
    sig_ref = (sxx64 - n * torch.outer(mu_ref, mu_ref)) / (n - 1)
    assert float((mu_a - mu_ref).abs().max()) <= 1e-12 * float(mu_ref.abs().max())
    assert float((sig_a - sig_ref).abs().max()) <= 1e-9 * float(sig_ref.abs().max())
    assert float((sig_a - sig_a.t()).abs().max()) == 0.0  # exactly symmetric


def test_cfg4_40k_ragged_crops_across_chunk_boundaries(engine):
    """cfg4: 40 000 variable-size crops in one resize call (the kernels chunk the grid at 32 768 images): crop i is a
    repeat of base crop i % 192, so equal inputs must give equal outputs wherever they fall; a sample is checked
    against the Pillow restatement."""
    sizes = synth.crop_sizes(8, 0, 192, lo=16, hi=300)
    base = [synth.images(5, 1, i, 1, h, w)[0] for i, (h, w) in enumerate(sizes)]
    flat = torch.cat([b.reshape(-1) for b in base]).cuda()
    nbytes = np.array([b.numel() for b in base], dtype=np.int64)
    boff = np.zeros(192, dtype=np.int64)
    boff[1:] = np.cumsum(nbytes)[:-1]
    n = 40000
    pick = np.arange(n) % 192
    packed = PackedImages(flat, boff[pick], np.array([list(base[k].shape[:2]) for k in pick], dtype=np.int32))
    out = engine.resize(packed)
    assert tuple(out.shape) == (n, 299, 299, 3)
    first = out[:192]
    for rep in (1, 100, 170, 171, 207):  # 170 * 192 = 32 640, 171 * 192 = 32 832: both sides of the 32 768 boundary
        blk = out[rep * 192:(rep + 1) * 192]
        assert bool((blk == first[:blk.shape[0]]).all().item()), rep
    for k in (0, 17, 95, 191):
        assert np.array_equal(first[k].cpu().numpy(), restate.pillow_resize_bilinear(base[k].numpy())), k
