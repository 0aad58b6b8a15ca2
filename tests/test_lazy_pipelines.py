"""The lazy dictionary transforms + fused collation (adell_mri_b200.transforms / transform_factory
/ collate) against the eager oracle restatement of the reference's pipelines
(oracle/pipelines_ref.py) on identical seeds.  CPU: plans execute through the C restatement;
GPU (marked): the same pipelines through the CUDA path."""

import numpy as np
import pytest
import torch

from adell_mri_b200 import collate, engine, transform_factory as F, transforms as T
from oracle import pipelines_ref as P
from tests.helpers import cref_execute, patch_engine_for_cpu


@pytest.fixture(params=["cpu", pytest.param("cuda:0", marks=pytest.mark.gpu)])
def dev(request, monkeypatch):
    if request.param == "cpu":
        patch_engine_for_cpu(monkeypatch)
    T.set_mode(strict=True, fast=False, noise="injected")
    yield request.param
    T.set_mode(strict=False)


def _samples(R, n, keys, shape, mask=True):
    out = []
    for _ in range(n):
        s = {k: torch.from_numpy(R.rand(1, *shape).astype(np.float32)) for k in keys}
        if mask:
            s["mask"] = torch.from_numpy((R.rand(1, *shape) > 0.7).astype(np.float32))
        out.append(s)
    return out


def _to(s, dev):
    return {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in s.items()}


@pytest.mark.parametrize("augment", [["affine", "flip"], ["affine", "shear", "flip"], ["trivial", "affine", "shear", "flip"]])
def test_unet_pipeline_matches_eager_reference(dev, augment):
    R = np.random.RandomState(0)
    keys, shape = ["t2", "adc", "dwi"], (28, 24, 12)
    samples = _samples(R, 6, keys, shape)
    lazy = T.Compose([F.get_augmentations_unet(augment, keys + ["mask"], keys, [], flip_axis=[0, 1, 2]),
                      T.ConcatItemsd(keys, "image"), T.SelectItemsd(["image", "mask"])]).set_random_state(11)
    ref = P.Chain([P.unet(augment, keys + ["mask"], keys, flip_axis=(0, 1, 2)), P.ConcatD(keys, "image")]).seed(11)
    batch = collate.safe_collate([lazy(_to(s, dev)) for s in samples])
    want = [ref(s) for s in samples]
    assert batch["image"].shape == (6, 3, *shape) and batch["mask"].shape == (6, 1, *shape)
    for b, w in enumerate(want):
        assert torch.equal(batch["image"][b].cpu(), w["image"])
        assert torch.equal(batch["mask"][b].cpu(), w["mask"].to(torch.float32))


def test_unet_all_fire_and_crop_sandwich(dev):
    """prob forced to 1 so every sample resamples twice (parity mode = two passes) inside the
    RandSpatialCropd -> augment -> CenterSpatialCropd sandwich (has_label=False)."""
    R = np.random.RandomState(1)
    keys, shape, rc = ["t2", "dwi"], (30, 28, 14), [20, 18, 10]
    samples = _samples(R, 4, keys, shape, mask=False)
    lazy_aug = F.get_augmentations_unet(["affine", "shear", "flip"], keys, keys, [], random_crop_size=rc, has_label=False, flip_axis=[0, 1])
    ref_aug = P.unet(["affine", "shear", "flip"], keys, keys, random_crop_size=rc, has_label=False, flip_axis=(0, 1))
    for t in lazy_aug.flatten().transforms:
        if isinstance(t, T.RandAffined):
            t.sampler.prob = 1.0
    for t in ref_aug.ts[1].ts:
        if isinstance(t, P.AffineD):
            t.draws.prob = 1.0
    lazy = T.Compose([lazy_aug, T.ConcatItemsd(keys, "image")]).set_random_state(5)
    ref = P.Chain([ref_aug, P.ConcatD(keys, "image")]).seed(5)
    batch = collate.safe_collate([lazy(_to(s, dev)) for s in samples])
    for b, s in enumerate(samples):
        assert torch.equal(batch["image"][b].cpu(), ref(s)["image"])


def test_unet_pos_neg_crops(dev):
    R = np.random.RandomState(2)
    keys, shape, rc = ["t2"], (32, 32, 12), [16, 16, 8]
    samples = _samples(R, 3, keys, shape)
    for s in samples:
        flat = (s["mask"] > 0).reshape(-1).numpy()
        s["mask_fg_indices"], s["mask_bg_indices"] = np.nonzero(flat)[0], np.nonzero(~flat)[0]
    lazy = T.Compose([F.get_augmentations_unet(["affine", "flip"], keys + ["mask"], keys, [], random_crop_size=rc, n_crops=2),
                      T.ConcatItemsd(keys, "image"), T.SelectItemsd(["image", "mask"])]).set_random_state(3)
    ref = P.Chain([P.unet(["affine", "flip"], keys + ["mask"], keys, random_crop_size=rc, n_crops=2), P.ConcatD(keys, "image")]).seed(3)
    got = collate.safe_collate_crops([lazy(_to(s, dev)) for s in samples])
    want = [c for s in samples for c in ref(s)]
    assert got["image"].shape == (6, 1, *rc)
    for b, w in enumerate(want):
        assert torch.equal(got["image"][b].cpu(), w["image"])
        assert torch.equal(got["mask"][b].cpu(), w["mask"])


@pytest.mark.parametrize("augment", [["flip", "affine"], ["flip", "affine", "shear"], ["trivial", "flip", "affine", "shear"]])
def test_classification_pipeline_matches_eager_reference(dev, augment):
    R = np.random.RandomState(4)
    keys, shape, crop = ["t2", "adc"], (40, 40, 24), [24, 24, 8]
    samples = _samples(R, 5, keys, shape)
    tf = F.ClassificationTransforms(keys, adc_keys=[], crop_size=crop, mask_key="mask")
    lazy_aug = F.get_augmentations_class(augment, keys, "mask", [], flip_axis=[0, 1, 2], prob=0.6)
    lazy = T.Compose([*tf.pre_transforms()[-2:], lazy_aug, *tf.post_transforms()]).set_random_state(9)
    m = [c + 16 for c in crop]
    ref = P.Chain([P.CenterCropD(keys + ["mask"], m), P.classification(augment, keys, "mask", flip_axis=(0, 1, 2), prob=0.6),
                   P.CenterCropD(keys + ["mask"], crop), P.ConcatD(keys + ["mask"], "image")]).seed(9)
    batch = collate.safe_collate([lazy(_to(s, dev)) for s in samples])
    assert batch["image"].shape == (5, 3, *crop)
    for b, s in enumerate(samples):
        assert torch.equal(batch["image"][b].cpu(), ref(s)["image"])


@pytest.mark.parametrize("different_crop,vicregl", [(False, False), (True, False), (False, True)])
def test_ssl_two_view_pipeline_matches_eager_reference(dev, different_crop, vicregl):
    R = np.random.RandomState(6)
    keys, copied, shape, roi = ["image"], ["image_copy"], (36, 32, 16), [24, 24, 12]
    samples = _samples(R, 5, keys, shape, mask=False)
    # contrast / rician_noise run on the statistics kernels (CUDA only): the CPU leg covers the single-launch members
    names = [m for m in F.FUSED_AUGMENTS if dev != "cpu" or m not in ("contrast", "rician_noise")]
    tf = F.SSLTransforms(keys, copied, adc_keys=[], non_adc_keys=[])
    lazy = tf.transforms(F.get_augmentations_ssl(keys, copied, None, roi, vicregl, different_crop, n_transforms=3,
                                                 aug_list=list(names))).set_random_state(21)
    ref = P.Chain(P.ssl(keys, copied, roi, vicregl, different_crop, names, 3)).seed(21)
    np.random.seed(123)
    got = collate.safe_collate([lazy(_to(s, dev)) for s in samples])
    np.random.seed(123)
    for b, s in enumerate(samples):
        d = dict(s)
        d["image_copy"] = s["image"].clone()
        w = ref(d)
        for gk, wk in (("augmented_image_1", "image"), ("augmented_image_2", "image_copy")):
            g, r = got[gk][b].cpu(), w[wk].to(torch.float32)
            # consecutive intensity members collapse into one {scale, offset} pair: rounding-level differences
            # (fp32 pow of the contrast member on the GPU leg: 2e-5 relative)
            tol = 2e-6 if dev == "cpu" else 2e-5
            assert torch.allclose(g, r, rtol=tol, atol=2e-6), (gk, float((g - r).abs().max()))
        if vicregl:
            for bk, wk in (("box_1", "image"), ("box_2", "image_copy")):
                assert np.array_equal(np.asarray(got[bk][b]), P.M.flatten_box(w["_cropped"][wk], roi))


def test_ssl_pre_transforms_intensity_and_copy(dev):
    """SSLTransforms.pre_transforms: min-max scaling (exact), centre crop then pad, CopyEntryd."""
    if dev == "cpu":
        pytest.skip("the statistics kernels need a CUDA device")
    R = np.random.RandomState(7)
    x = torch.from_numpy((R.rand(1, 20, 18, 10) * 900 + 30).astype(np.float32))
    tf = F.SSLTransforms(["image"], ["image_copy"], adc_keys=[], non_adc_keys=["image"], crop_size=[16, 16, 16], pad_size=[24, 24, 12])
    d = T.Compose(tf.pre_transforms())({"image": x.to(dev)})
    want = P.M.spatial_pad(P.M.center_spatial_crop(P.M.scale_intensity(x, 0.0, 1.0), [16, 16, 16]), [24, 24, 12])
    assert torch.equal(d["image"].tensor().cpu(), want)
    assert torch.equal(d["image_copy"].tensor().cpu(), want)


@pytest.mark.gpu
@pytest.mark.parametrize("augment", [["intensity"], ["intensity", "affine", "flip"], ["trivial", "intensity", "affine"]])
def test_unet_intensity_token_matches_eager_reference(augment):
    """--augment intensity (RandAdjustContrastd + RandStdShiftIntensityd before the spatial members):
    same draws as the eager oracle chain; values within fp32 pow / reduction-order tolerance; the mask
    (untouched by the intensity members) stays bit-exact."""
    dev = "cuda:0"
    T.set_mode(strict=True, fast=False, noise="injected")
    try:
        R = np.random.RandomState(8)
        keys, shape = ["t2", "adc"], (28, 24, 12)
        samples = _samples(R, 8, keys, shape)
        lazy_aug = F.get_augmentations_unet(augment, keys + ["mask"], keys, [], flip_axis=[0, 1, 2])
        ref_aug = P.unet(augment, keys + ["mask"], keys, flip_axis=(0, 1, 2))
        if "trivial" not in augment:   # make the members fire often enough to be exercised
            for t in lazy_aug.flatten().transforms:
                if isinstance(t, (T.RandAdjustContrastd, T.RandStdShiftIntensityd)):
                    t.prob = 0.7
            for t in ref_aug.ts:
                if isinstance(t, (P.ContrastD, P.StdShiftD)):
                    t.prob = 0.7
        lazy = T.Compose([lazy_aug, T.ConcatItemsd(keys, "image"), T.SelectItemsd(["image", "mask"])]).set_random_state(13)
        ref = P.Chain([ref_aug, P.ConcatD(keys, "image")]).seed(13)
        batch = collate.safe_collate([lazy(_to(s, dev)) for s in samples])
        fired = 0
        for b, s in enumerate(samples):
            w = ref(s)
            g, r = batch["image"][b].cpu(), w["image"]
            assert torch.allclose(g, r, rtol=2e-5, atol=2e-6), (b, float((g - r).abs().max()))
            assert torch.equal(batch["mask"][b].cpu(), w["mask"].to(torch.float32))
            fired += int(not torch.equal(r, torch.cat([s[k] for k in keys], 0)))
        assert fired >= 3
    finally:
        T.set_mode(strict=False)


@pytest.mark.gpu
def test_classification_intensity_token_matches_eager_reference():
    dev = "cuda:0"
    T.set_mode(strict=True, fast=False, noise="injected")
    try:
        R = np.random.RandomState(9)
        keys, shape = ["t2", "adc"], (30, 28, 14)
        samples = _samples(R, 8, keys, shape)
        augment = ["intensity", "flip", "affine"]
        lazy = T.Compose([F.get_augmentations_class(augment, keys, "mask", [], flip_axis=[0, 1], prob=0.6),
                          T.ConcatItemsd(keys + ["mask"], "image")]).set_random_state(31)
        ref = P.Chain([P.classification(augment, keys, "mask", flip_axis=(0, 1), prob=0.6), P.ConcatD(keys + ["mask"], "image")]).seed(31)
        batch = collate.safe_collate([lazy(_to(s, dev)) for s in samples])
        for b, s in enumerate(samples):
            g, r = batch["image"][b].cpu(), ref(s)["image"]
            assert torch.allclose(g[:2], r[:2], rtol=2e-5, atol=2e-6), (b, float((g - r).abs().max()))
            assert torch.equal(g[2], r[2])   # the mask channel
    finally:
        T.set_mode(strict=False)


def test_unknown_and_out_of_scope_tokens_raise():
    with pytest.raises(NotImplementedError):
        F.get_augmentations_unet(["bogus"], ["a"], ["a"], [])
    with pytest.raises(NotImplementedError):
        F.get_augmentations_unet(["blur"], ["a"], ["a"], [])
    with pytest.raises(NotImplementedError):
        F.get_augmentations_class(["noise"], ["a"], None, [])
    with pytest.raises(NotImplementedError):   # a workhorse member outside the fused path
        F.get_augmentations_ssl(["a"], ["b"], None, [8, 8, 8], False, False, aug_list=["gaussian_noise", "gibbs_noise", "rotate_x"])
    with pytest.raises(NotImplementedError):
        T.Resized(["a"], [8, 8, 8], mode="trilinear")


@pytest.mark.gpu
@pytest.mark.parametrize("members", [["rician_noise", "contrast", "rotate_z"], ["contrast", "rician_noise", "shift_intensity", "gaussian_noise"]])
def test_workhorse_pointwise_members_match_eager_reference(members):
    """§8(f) row 1: the workhorse's `contrast` (gamma ~ U(0.5, 3*0.5 + 0.51)) and `rician_noise`
    (std 0.2*0.5, two normal volumes per key) members next to the single-launch ones; every member
    fires (N = len(members)), the order comes from the global stream."""
    dev = "cuda:0"
    T.set_mode(strict=True, fast=False, noise="injected")
    try:
        R = np.random.RandomState(12)
        keys, shape = ["image"], (26, 22, 12)
        samples = _samples(R, 6, keys, shape, mask=False)
        lazy = F.AugmentationWorkhorsed(members, keys, [], max_mult=0.5, N=len(members)).set_random_state(5)
        ref = P.Workhorse(members, keys, len(members)).seed(5)
        np.random.seed(3)
        got = [lazy(_to(s, dev)) for s in samples]
        got = [(g["image"].tensor() if isinstance(g["image"], T.Pending) else g["image"]).cpu() for g in got]
        np.random.seed(3)
        for b, s in enumerate(samples):
            r = ref(dict(s))["image"].to(torch.float32)
            assert torch.allclose(got[b], r, rtol=2e-5, atol=2e-6), (b, float((got[b] - r).abs().max()))
            assert not torch.equal(r, s["image"])
    finally:
        T.set_mode(strict=False)


@pytest.mark.gpu
def test_ssl_scaled_crop_matches_eager_reference():
    """get_augmentations_ssl(scaled_crop_size=s): SpatialPadd(s/2) -> RandSpatialCropd(s/2, random_size=True) ->
    Resized(s, "area") -> shared RandSpatialCropd(roi) (augmentations.py:427-444,487-492; skip_augmentations keeps the
    chain deterministic after the crops); the resized volumes are bit-identical to ATen's."""
    dev = "cuda:0"
    R = np.random.RandomState(14)
    keys, copied, shape, scaled, roi = ["image"], ["image_copy"], (44, 30, 18), [40, 40, 16], [40, 40, 16]
    samples = _samples(R, 6, keys, shape, mask=False)
    small = [x // 2 for x in scaled]
    crop_only = F.get_augmentations_ssl(keys, copied, scaled, roi, False, False, skip_augmentations=True)
    assert [type(t).__name__ for t in crop_only] == ["SpatialPadd", "RandSpatialCropd", "Resized"]
    lazy = T.Compose(crop_only).set_random_state(33)
    seeds = P.M.compose_set_random_state(33, 1)
    Rc = np.random.RandomState(seeds[0])
    sizes = set()
    for s in samples:
        d = {"image": s["image"].to(dev), "image_copy": s["image"].clone().to(dev)}
        out = lazy(d)
        x = P.M.spatial_pad(s["image"], small)
        starts, size = P.M.rand_spatial_crop_draw(Rc, tuple(x.shape[1:]), small, random_size=True)
        sizes.add(tuple(size))
        want = P.M.resized(P.M.crop(x, starts, size), scaled, "area")
        for k in ("image", "image_copy"):
            g = out[k].tensor() if isinstance(out[k], T.Pending) else out[k]
            assert tuple(g.shape) == (1, *scaled)
            assert torch.equal(g.cpu(), want), k
    assert len(sizes) > 1    # random window sizes were drawn


@pytest.mark.parametrize("seed", [0, 3])
def test_randomised_builder_configurations_match_the_eager_reference(seed):
    """tools/fuzz_pipelines.py as a test (CPU, through the C restatement): random key sets, shapes, flip axes,
    probabilities, crop sandwiches and augment lists of the unet / classification / SSL builders, lazy surface +
    fused collation against the eager oracle chains on the same seeds; and the segmentation / classification / SSL batch
    fast paths against the dictionary surface (plain and nested seeding; the crop sandwich's nested seed fan-out was
    flat in the batch path until this sweep found it)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import fuzz_pipelines

    n, bad = fuzz_pipelines.sweep(56, seed)
    assert n == 56 and bad == 0


@pytest.mark.gpu
def test_randomised_builder_configurations_through_the_cuda_path():
    """The same sweep with the plans executed by the CUDA kernels instead of the C restatement."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import fuzz_pipelines

    n, bad = fuzz_pipelines.sweep(70, 4, device="cuda:0")
    assert n == 70 and bad == 0
