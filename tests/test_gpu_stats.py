"""GPU: K2/K3 statistics kernels vs the oracle (min/max, exact scalers, exact percentiles)."""

import numpy as np
import pytest
import torch

from adell_mri_b200 import _lib, stats
from oracle import monai_restated as M

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _vols(R, n, shape, kind):
    out = []
    for i in range(n):
        if kind == "uniform":
            v = R.rand(*shape).astype(np.float32) * (i + 1) * 700
        elif kind == "lognormal":
            v = R.lognormal(0, 1, size=shape).astype(np.float32)
        elif kind == "signed":
            v = R.normal(0, 100, size=shape).astype(np.float32)
        else:  # mri-like: 40 % exact-zero background + skewed foreground
            v = R.gamma(2.0, 300.0, size=shape).astype(np.float32)
            v[R.rand(*shape) < 0.4] = 0.0
        out.append(torch.from_numpy(v))
    return out


@pytest.mark.parametrize("kind", ["uniform", "lognormal", "signed", "mri"])
def test_minmax_and_exact_scalers(kind):
    R = np.random.RandomState(0)
    vols = _vols(R, 3, (40, 36, 20), kind)
    dv = [v.to(DEV) for v in vols]
    mm = stats.minmax(dv)
    for i, v in enumerate(vols):
        assert float(mm[i, 0]) == float(v.min()) and float(mm[i, 1]) == float(v.max())
    # ScaleIntensityd(minv=0, maxv=1)
    outs = stats.intensity_map(dv, stats.scaler_coefs(mm, _lib.SCALER_MINMAX, 0.0, 1.0))
    for o, v in zip(outs, vols):
        assert torch.equal(o.cpu(), M.scale_intensity(v[None], 0.0, 1.0)[0])
    # GenerationTransforms-style minv=-1, maxv=1
    outs = stats.intensity_map(dv, stats.scaler_coefs(mm, _lib.SCALER_MINMAX, -1.0, 1.0))
    for o, v in zip(outs, vols):
        assert torch.equal(o.cpu(), M.scale_intensity(v[None], -1.0, 1.0)[0])
    # ADC (segmentation order): ConditionalRescalingd(500, .001) -> ScaleIntensityd(factor=-2/3)
    outs = stats.intensity_map(dv, stats.scaler_coefs(mm, _lib.SCALER_ADC_SEG, 500, 0.001))
    for o, v in zip(outs, vols):
        ref = M.scale_intensity(M.conditional_rescaling(v[None], 500, 0.001), None, None, -2 / 3)[0]
        assert torch.equal(o.cpu(), ref)
    # ADC (classification order): + Offsetd(None) in between
    outs = stats.intensity_map(dv, stats.scaler_coefs(mm, _lib.SCALER_ADC_CLASS, 500, 0.001))
    for o, v in zip(outs, vols):
        ref = M.scale_intensity(M.offset(M.conditional_rescaling(v[None], 500, 0.001)), None, None, -2 / 3)[0]
        assert torch.equal(o.cpu(), ref)


def test_constant_volume_minmax_scaler():
    v = torch.full((8, 8, 8), 3.5)
    dv = [v.to(DEV)]
    out = stats.intensity_map(dv, stats.scaler_coefs(stats.minmax(dv), _lib.SCALER_MINMAX, 0.0, 1.0))[0]
    assert torch.equal(out.cpu(), M.scale_intensity(v[None], 0.0, 1.0)[0])


@pytest.mark.parametrize("kind", ["uniform", "lognormal", "signed", "mri"])
@pytest.mark.parametrize("shape", [(30, 20, 10), (128, 100, 90)])
def test_percentiles_match_numpy(kind, shape):
    R = np.random.RandomState(1)
    vols = _vols(R, 2, shape, kind)
    qs = [0.5, 99.5]
    got = stats.percentiles([v.to(DEV) for v in vols], qs).cpu().numpy()
    for i, v in enumerate(vols):
        ref = np.percentile(v.numpy().reshape(-1), np.asarray(qs, np.float64)).astype(np.float32)
        assert np.array_equal(got[i], ref), (got[i], ref)


def test_percentiles_edge_quantiles_and_int16():
    R = np.random.RandomState(2)
    v = torch.from_numpy(R.randint(-300, 4000, size=(50, 40, 30)).astype(np.int16))
    qs = [0.0, 1.0, 50.0, 100.0]
    got = stats.percentiles([v.to(DEV)], qs).cpu().numpy()[0]
    ref = np.percentile(v.numpy().astype(np.float32).reshape(-1), np.asarray(qs, np.float64)).astype(np.float32)
    assert np.array_equal(got, ref)
    u = torch.from_numpy(R.randint(0, 256, size=(33, 31, 7)).astype(np.uint8))
    got = stats.percentiles([u.to(DEV)], [2.0, 98.0]).cpu().numpy()[0]
    ref = np.percentile(u.numpy().astype(np.float32).reshape(-1), np.asarray([2.0, 98.0])).astype(np.float32)
    assert np.array_equal(got, ref)


def test_dataset_wide_percentile_pools_volumes():
    R = np.random.RandomState(3)
    vols = _vols(R, 4, (32, 32, 16), "lognormal")
    got = stats.percentiles([v.to(DEV) for v in vols], [1.0, 99.0], dataset_wide=True).cpu().numpy()[0]
    pooled = np.concatenate([v.numpy().reshape(-1) for v in vols])
    ref = np.percentile(pooled, np.asarray([1.0, 99.0])).astype(np.float32)
    assert np.array_equal(got, ref)


def test_percentile_scaler_fused_and_exact():
    """ScaleIntensityRangePercentilesd(0.5, 99.5, 0, 1, clip=True): exact program vs oracle and
    fused {scale, offset} + per-tap clip inside K1 within 1e-4."""
    from adell_mri_b200.plan import BatchPlan
    from tests.helpers import run_plan_cuda

    R = np.random.RandomState(4)
    v = _vols(R, 1, (128, 100, 90), "mri")[0]
    dv = v.to(DEV)
    pct = stats.percentiles([dv], [0.5, 99.5])
    coefs = stats.scaler_coefs(pct, _lib.SCALER_RANGE, 0.0, 1.0)
    exact = stats.intensity_map([dv], coefs, clip=(0.0, 1.0))[0]
    ref = M.scale_intensity_range_percentiles(v[None], 0.5, 99.5, 0.0, 1.0, clip=True)[0]
    assert torch.equal(exact.cpu(), ref)
    plan = BatchPlan([dv]).intensity_from_device(stats.coefs_to_affine(coefs)).clip(0.0, 1.0)
    fused = run_plan_cuda(plan)[0]
    assert torch.allclose(fused.cpu(), ref, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("nonzero,channel_wise", [(False, False), (True, False), (False, True)])
def test_normalize_intensity_transform(nonzero, channel_wise):
    """NormalizeIntensityd (north-star name; z-score) vs the oracle: fp64-accumulated statistics,
    tolerance 1e-5 relative to the dynamic range (torch reduces in fp32)."""
    from adell_mri_b200 import transforms as T

    R = np.random.RandomState(3)
    x = torch.from_numpy(R.gamma(2.0, 300.0, size=(2, 40, 36, 20)).astype(np.float32))
    x[R.rand(*x.shape) < 0.3] = 0.0
    got = T.NormalizeIntensityd(["image"], nonzero=nonzero, channel_wise=channel_wise)({"image": x.to(DEV)})["image"].cpu()
    want = M.normalize_intensity(x, nonzero=nonzero, channel_wise=channel_wise)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-5 * float(want.abs().max()))
    ms = stats.meanstd([x[0].contiguous().to(DEV)])
    assert abs(float(ms[0, 0]) - float(x[0].double().mean())) <= 1e-6 * abs(float(x[0].double().mean()))
    assert abs(float(ms[0, 1]) - float(x[0].double().std(unbiased=False))) <= 1e-6 * float(x[0].double().std(unbiased=False))
    # constant volume: std 0 -> 1
    c = torch.full((1, 8, 8, 8), 2.5)
    assert torch.equal(T.NormalizeIntensityd(["i"])({"i": c.to(DEV)})["i"].cpu(), torch.zeros_like(c))


@pytest.mark.parametrize("clip", [False, True])
def test_percentile_scaler_transform_then_classification_chain(clip):
    """Config D: ScaleIntensityRangePercentilesd(0.5, 99.5 -> [0,1]) on the cached volume (exact),
    then the classification chain flip -> affine(zeros) -> centre crop through the lazy surface."""
    from adell_mri_b200 import collate, transform_factory as F, transforms as T
    from oracle import pipelines_ref as P

    R = np.random.RandomState(8)
    shape, crop = (56, 56, 40), [40, 40, 24]
    raw = [torch.from_numpy(R.gamma(2.0, 300.0, size=(1, *shape)).astype(np.float32)) for _ in range(3)]
    T.set_mode(strict=True)
    try:
        scaler = T.ScaleIntensityRangePercentilesd(["t2"], 0.5, 99.5, 0.0, 1.0, clip=clip)
        tf = F.ClassificationTransforms(["t2"], adc_keys=[], crop_size=crop)
        aug = F.get_augmentations_class(["flip", "affine"], ["t2"], None, [], flip_axis=[0, 1, 2], prob=0.7)
        lazy = T.Compose([scaler, *tf.pre_transforms()[-2:], aug, *tf.post_transforms()]).set_random_state(4)
        ref = P.Chain([P.CenterCropD(["t2"], [c + 16 for c in crop]), P.classification(["flip", "affine"], ["t2"], None, (0, 1, 2), 0.7),
                       P.CenterCropD(["t2"], crop), P.ConcatD(["t2"], "image")]).seed(4)
        got = collate.safe_collate([lazy({"t2": r.to(DEV)}) for r in raw])["image"].cpu()
        for b, r in enumerate(raw):
            scaled = M.scale_intensity_range_percentiles(r, 0.5, 99.5, 0.0, 1.0, clip=clip)
            # volumes under 1e6 voxels: MONAI interpolates the percentile with torch.quantile in fp32, the
            # device kernels with numpy's float64 'linear' rule (exact for the BASELINE sizes, which
            # are above that threshold: test_percentile_scaler_fused_and_exact) -> 1-ulp level differences
            assert torch.allclose(got[b], ref({"t2": scaled})["image"], rtol=1e-5, atol=1e-6)
    finally:
        T.set_mode(strict=False)


@pytest.mark.gpu
@pytest.mark.parametrize("n,offset", [(0, 0), (1, 0), (7, 1), (4096, 0), (40 * 36 * 20 + 3, 0), (40 * 36 * 20 + 3, 3), (1 << 22, 0)])
def test_rician_map_is_bit_exact(n, offset):
    """adell_rician_map = sqrt((x + n1) ** 2 + n2 ** 2) with every op IEEE-rounded to fp32 (vectorised body,
    scalar tail, unaligned views): bit-exact against numpy op by op; torch's CPU sqrt goes through MKL VML,
    which is within 1 ulp but not correctly rounded (0.7 % of the elements differ by one ulp)."""
    R = np.random.RandomState(n % 1000 + offset)
    x = torch.from_numpy(R.gamma(2.0, 0.3, size=n + offset).astype(np.float32))[offset:]
    want, n1, n2 = M.rand_rician_noise(x, np.random.RandomState(4), 0.1)
    xd = torch.empty(n + offset, dtype=torch.float32, device=DEV)[offset:]   # offset: a view that is not 16-byte aligned
    xd.copy_(x)
    got = stats.rician_map(xd, n1.to(DEV), n2.to(DEV))
    a = x.numpy() + n1.numpy()
    ieee = np.sqrt(a * a + n2.numpy() * n2.numpy())
    assert ieee.dtype == np.float32
    assert got.shape == x.shape and np.array_equal(got.cpu().numpy(), ieee)
    assert torch.allclose(got.cpu(), want, rtol=2e-7, atol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("gamma", [0.5, 0.77, 1.0, 2.01, 4.5])
@pytest.mark.parametrize("kind", ["f32", "f32_unaligned", "i16"])
def test_gamma_map_error_is_bounded_by_1e6_of_the_range(gamma, kind):
    """adell_gamma_map vs monai AdjustContrast restated in float64 and in torch fp32: the SFU power keeps
    the absolute error below 2e-6 of the intensity range (the path is held to 1e-4); min and max map to
    themselves (pow(0) = 0, pow(1 - eps) ~ 1)."""
    R = np.random.RandomState(int(gamma * 100))
    n = 50_003
    if kind == "i16":
        x = torch.from_numpy(R.randint(-200, 4000, size=n).astype(np.int16))
    else:
        x = torch.from_numpy(np.concatenate([R.gamma(2.0, 300.0, size=n - 3), [0.0, 1e-3, 5e3]]).astype(np.float32))
    xd = x.to(DEV)
    if kind == "f32_unaligned":
        buf = torch.empty(n + 1, dtype=torch.float32, device=DEV)
        buf[1:].copy_(x)
        xd = buf[1:]
    mm = stats.minmax([xd])
    got = stats.gamma_map([xd], mm, gamma)[0].cpu()
    xf = x.to(torch.float64)
    lo, rng = xf.min(), xf.max() - xf.min()
    want64 = ((xf - lo) / (rng + 1e-7)) ** gamma * rng + lo
    assert float((got.to(torch.float64) - want64).abs().max()) <= 2e-6 * float(rng)
    want32 = M.adjust_contrast(x.to(torch.float32), gamma)
    assert torch.allclose(got, want32, rtol=2e-5, atol=2e-6 * float(rng))
    assert got[x.argmin()] == float(lo)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["area", "nearest"])
def test_resize_is_bit_identical_to_torch_interpolate(mode):
    """adell_resize vs F.interpolate on the CPU (area = ATen adaptive_avg_pool3d): shrinking, growing, mixed,
    exact halves / doubles, identity, volumes of different input shapes in one call."""
    R = np.random.RandomState(3)
    cases = [((80, 72, 20), (64, 64, 16)), ((41, 37, 11), (64, 48, 16)), ((33, 80, 20), (48, 40, 24)), ((32, 32, 8), (64, 64, 16)),
             ((64, 64, 16), (32, 32, 8)), ((24, 20, 12), (24, 20, 12)), ((97, 101, 37), (80, 80, 20)), ((5, 3, 2), (7, 9, 4)),
             ((41, 37, 11), (30, 30, 10)), ((20, 20, 9), (33, 17, 7)), ((90, 64, 50), (16, 16, 6))]   # rows that are not a multiple of four; 6-wide windows
    for out in sorted({o for _, o in cases}):
        ins = [i for i, o in cases if o == out]
        for _ in range(2):   # two more inputs of random extents between half and twice the output
            ins.append(tuple(int(R.randint(out[ax] // 2 + 1, 2 * out[ax] + 2)) for ax in range(3)))
        vols = [torch.from_numpy((R.rand(*i) * 1000 - 200).astype(np.float32)) for i in ins]
        got = stats.resize([v.to(DEV) for v in vols], out, mode)
        for v, g in zip(vols, got):
            want = torch.nn.functional.interpolate(v[None, None], size=out, mode=mode)[0, 0]
            assert g.shape == want.shape and torch.equal(g.cpu(), want), (tuple(v.shape), out, mode)


@pytest.mark.gpu
def test_resize_full_size_volume_matches_torch():
    """Config-E sized input (512x512x128 -> 256x256x64, exact 2x2x2 means; and -> 300x300x80, ragged windows)."""
    g = torch.Generator().manual_seed(0)
    v = torch.rand((512, 512, 128), generator=g)
    for out in [(256, 256, 64), (300, 300, 80)]:
        got = stats.resize([v.to(DEV)], out, "area")[0].cpu()
        want = torch.nn.functional.interpolate(v[None, None], size=out, mode="area")[0, 0]
        assert torch.equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [0, 7])
def test_randomised_statistics_sweep_against_numpy(seed):
    """tools/fuzz_stats.py as a test: exact percentiles (per volume and pooled) and min / max of adversarial volumes
    — constants, two values, heavy ties, signed zeros, denormals, 60 decades of range, int16 / uint8, one element,
    unaligned views, mixed sizes in one call.  Seed 0 holds the pooled near-tie that was one ulp off while the
    percentile positions used the alpha = beta = 1 formula instead of numpy's (n - 1) * q."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import fuzz_stats

    checked, bad = fuzz_stats.sweep(60, seed)
    assert checked > 100 and bad == 0


# ----------------------------------------------------------------------------- one-read order statistics
def _one_read(vols, qs):
    """(percentiles through adell_quantile_keys, through the three radix passes, whether the one-read call fell back)."""
    dv = [v.to(DEV) for v in vols]
    kern = stats._CudaKernels(dv)
    a = stats.percentiles(dv, qs, kernels=kern).cpu().numpy()
    fb = kern.fell_back(len(qs))
    b = stats.percentiles(dv, qs, one_read=False).cpu().numpy()
    return a, b, fb


@pytest.mark.parametrize("kind", ["uniform", "lognormal", "signed", "mri"])
@pytest.mark.parametrize("qs", [[0.5, 99.5], [1.0, 99.0], [50.0], [0.0, 25.0, 75.0, 100.0]])
def test_one_read_percentiles_equal_radix_passes_and_numpy(kind, qs):
    """Volumes larger than the sample (so the brackets ARE estimates): the sampled brackets hold (no fallback) and the
    keys selected among the candidates equal the three radix passes and np.percentile bit for bit."""
    R = np.random.RandomState(5)
    vols = _vols(R, 3, (160, 128, 48), kind)     # 983 040 elements each, 30 x the sample
    a, b, fb = _one_read(vols, qs)
    assert np.array_equal(a, b)
    assert not fb
    for i, v in enumerate(vols):
        ref = np.percentile(v.numpy().reshape(-1), np.asarray(qs, np.float64)).astype(np.float32)
        assert np.array_equal(a[i], ref), (a[i], ref)


def test_one_read_percentiles_adversarial_layouts_fall_back_and_stay_exact():
    """A volume whose values follow the period of the sample stride shows the sample a single phase: its brackets miss,
    the flag un-gates the radix passes inside the same call, and the result is still exact.  Next to it in the same
    call: a constant volume, heavy ties, a sorted ramp, a tiny volume."""
    R = np.random.RandomState(6)
    n = 160 * 128 * 48
    stride = 4 * ((n // 4) // (32768 // 4))          # elements between two sampled groups (st_quantile_sample)
    phase = (np.arange(n) % stride).astype(np.float32)    # sampled positions all hold 0..3
    periodic = torch.from_numpy(phase + R.rand(n).astype(np.float32) * 0.5).reshape(160, 128, 48)
    const = torch.full((160, 128, 48), 3.25)
    ties = torch.from_numpy(R.randint(0, 5, size=n).astype(np.float32)).reshape(160, 128, 48)
    ramp = torch.arange(n, dtype=torch.float32).reshape(160, 128, 48)
    qs = [0.5, 99.5]
    for vols, expect_fb in (([periodic, const, ties, ramp], True), ([const, ties, ramp], False)):
        a, b, fb = _one_read(vols, qs)
        assert fb == expect_fb
        assert np.array_equal(a, b)
        for i, v in enumerate(vols):
            ref = np.percentile(v.numpy().reshape(-1), np.asarray(qs, np.float64)).astype(np.float32)
            assert np.array_equal(a[i], ref), (i, a[i], ref)
    tiny = [torch.tensor([5.0]).reshape(1, 1, 1), torch.tensor([2.0, -1.0]).reshape(2, 1, 1),
            torch.from_numpy(R.rand(7, 5, 3).astype(np.float32))]
    for v in tiny:    # (volumes of different sizes go in separate calls here: ranks depend on n)
        a, b, fb = _one_read([v], [0.0, 50.0, 100.0])
        assert not fb and np.array_equal(a, b)
        assert np.array_equal(a[0], np.percentile(v.numpy().reshape(-1), [0.0, 50.0, 100.0]).astype(np.float32))


def test_one_read_percentiles_integer_sources_and_mixed_sizes():
    R = np.random.RandomState(7)
    v16 = [torch.from_numpy(R.randint(-300, 4000, size=s).astype(np.int16)) for s in ((96, 96, 40), (50, 40, 30))]
    a, b, fb = _one_read(v16, [0.5, 50.0, 99.5])
    assert np.array_equal(a, b) and not fb
    for i, v in enumerate(v16):
        assert np.array_equal(a[i], np.percentile(v.numpy().astype(np.float32).reshape(-1), [0.5, 50.0, 99.5]).astype(np.float32))
    u8 = [torch.from_numpy((R.rand(128, 128, 32) > 0.7).astype(np.uint8) * 200)]
    a, b, fb = _one_read(u8, [1.0, 99.0])
    assert np.array_equal(a, b) and not fb
    assert np.array_equal(a[0], np.percentile(u8[0].numpy().astype(np.float32).reshape(-1), [1.0, 99.0]).astype(np.float32))


def test_one_read_pooled_percentiles_equal_radix_passes_and_numpy():
    """Dataset-wide (pooled) statistics of one rank through the one-read path: volumes of different sizes pooled into one
    sample, one bracket set, one candidate list."""
    R = np.random.RandomState(9)
    vols = [torch.from_numpy(R.lognormal(0, 1, size=s).astype(np.float32)) for s in ((96, 96, 40), (128, 100, 30), (64, 64, 64))]
    dv = [v.to(DEV) for v in vols]
    pooled = np.concatenate([v.numpy().reshape(-1) for v in vols])
    for qs in ([1.0, 99.0], [50.0], [0.0, 0.5, 99.5, 100.0]):
        kern = stats._CudaKernels(dv)
        a = stats.percentiles(dv, qs, dataset_wide=True, kernels=kern).cpu().numpy()
        assert kern.last_pooled == pooled.size and not kern.fell_back(len(qs))
        b = stats.percentiles(dv, qs, dataset_wide=True, one_read=False).cpu().numpy()
        ref = np.percentile(pooled, np.asarray(qs, np.float64)).astype(np.float32)
        assert np.array_equal(a, b) and np.array_equal(a[0], ref), (a, b, ref)
    # a pool too small for the sampled path takes the radix passes (same result)
    tiny = [torch.from_numpy(R.rand(10, 10, 10).astype(np.float32)).to(DEV) for _ in range(3)]
    got = stats.percentiles(tiny, [5.0, 95.0], dataset_wide=True).cpu().numpy()[0]
    ref = np.percentile(np.concatenate([t.cpu().numpy().reshape(-1) for t in tiny]), [5.0, 95.0]).astype(np.float32)
    assert np.array_equal(got, ref)


def test_one_read_brackets_are_reused_and_a_stale_bracket_falls_back():
    """The second call on the same kernel object (= the same cached volumes) skips the sampling kernel and reuses the
    brackets left in its workspace; when the data changed in place those brackets miss, the gated radix passes run,
    and the percentiles are still exact."""
    R = np.random.RandomState(10)
    vols = _vols(R, 2, (160, 128, 48), "lognormal")
    dv = [v.to(DEV) for v in vols]
    qs = [0.5, 99.5]
    kern = stats._CudaKernels(dv)
    a = stats.percentiles(dv, qs, kernels=kern).cpu().numpy()
    assert kern.brackets_of is not None and not kern.fell_back(2)
    b = stats.percentiles(dv, qs, kernels=kern).cpu().numpy()          # reused brackets
    assert np.array_equal(a, b) and not kern.fell_back(2)
    for i, v in enumerate(vols):
        assert np.array_equal(a[i], np.percentile(v.numpy().reshape(-1), np.asarray(qs)).astype(np.float32))
    dv[0].mul_(1000.0).add_(5.0)                                         # same storage, new content
    c = stats.percentiles(dv, qs, kernels=kern).cpu().numpy()
    assert kern.fell_back(2)
    assert np.array_equal(c[0], np.percentile(dv[0].cpu().numpy().reshape(-1), np.asarray(qs)).astype(np.float32))
    assert np.array_equal(c[1], a[1])


def test_one_read_percentiles_long_candidate_lists():
    """Volumes / pools of >= 2^23 elements select among their candidate lists with the radix kernels on the whole grid
    (st_quantile_place / merge) instead of one block per list."""
    R = np.random.RandomState(11)
    big = torch.from_numpy(R.lognormal(0, 1, size=(256, 256, 128)).astype(np.float32))
    big[R.rand(256, 256, 128) < 0.3] = 0.0                                 # a background level: heavy ties at the low end
    for qs in ([1.0, 99.0], [0.5, 50.0, 99.5]):
        a, b, fb = _one_read([big], qs)
        assert not fb and np.array_equal(a, b)
        assert np.array_equal(a[0], np.percentile(big.numpy().reshape(-1), np.asarray(qs, np.float64)).astype(np.float32))
    parts = [p.contiguous() for p in big.reshape(4, 64, 256, 128)]
    dv = [p.to(DEV) for p in parts]
    kern = stats._CudaKernels(dv)
    got = stats.percentiles(dv, [1.0, 99.0], dataset_wide=True, kernels=kern).cpu().numpy()[0]
    assert kern.last_pooled == big.numel() and not kern.fell_back(2)
    assert np.array_equal(got, np.percentile(big.numpy().reshape(-1), [1.0, 99.0]).astype(np.float32))
