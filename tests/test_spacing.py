"""Spacingd of the cached stage (reference transform_factory/transforms.py:133-140,
444-454) as one K1 resample item, against the oracle restatement of MONAI's Spacing -> SpatialResample († from MONAI
1.3-1.6: same origin, index-to-index mapping with align_corners=False, border padding, float64 grid).  CPU: the
recorded plan runs through the C restatement; GPU (marked): through CUDA."""

import numpy as np
import pytest
import torch

from adell_mri_b200 import engine, transform_factory as F, transforms as T
from oracle import monai_restated as M
from tests.helpers import cref_execute, patch_engine_for_cpu


@pytest.fixture(params=["cpu", pytest.param("cuda:0", marks=pytest.mark.gpu)])
def dev(request, monkeypatch):
    if request.param == "cpu":
        patch_engine_for_cpu(monkeypatch)
    yield request.param


def test_output_size_follows_compute_shape_offset():
    # round((S - 1) * spacing / pixdim + 1), numpy's round half to even
    assert list(T.Spacingd.output_size((256, 256, 32), (0.5, 0.5, 3.0), (0.5, 0.5, 3.0))) == [256, 256, 32]
    assert list(T.Spacingd.output_size((256, 256, 32), (0.5, 0.5, 3.0), (1.0, 1.0, 3.0))) == [128, 128, 32]   # 127.5 + 1 -> 128
    assert list(T.Spacingd.output_size((101, 64, 20), (0.8, 0.8, 3.3), (0.5, 1.6, 3.0))) == [161, 32, 22]


@pytest.mark.parametrize("spacing,pixdim", [((0.5, 0.5, 3.0), (1.0, 1.0, 3.0)), ((0.8, 0.7, 3.3), (0.5, 0.5, 3.0)),
                                             ((1.0, 1.0, 1.0), (2.0, 0.5, 1.0))])
def test_spacingd_matches_the_restated_spatial_resample(dev, spacing, pixdim):
    R = np.random.RandomState(0)
    shape = (40, 36, 12)
    img = torch.from_numpy(R.rand(1, *shape).astype(np.float32))
    lab = torch.from_numpy((R.rand(1, *shape) > 0.6).astype(np.float32))
    t = T.Spacingd(["image", "mask"], pixdim=pixdim, mode=["bilinear", "nearest"])
    out = t({"image": img.to(dev), "mask": lab.to(dev), "image_spacing": spacing, "mask_spacing": spacing})
    got_i, got_m = out["image"].tensor().cpu(), out["mask"].tensor().cpu()
    want_i = M.spacing(img, spacing, pixdim, "bilinear")
    want_m = M.spacing(lab, spacing, pixdim, "nearest")
    assert got_i.shape == want_i.shape and got_m.shape == want_m.shape
    assert torch.allclose(got_i, want_i, rtol=1e-4, atol=1e-4)
    # nearest: bit-exact wherever the source coordinate o * pixdim / spacing is not a rounding tie.  AT a tie (e.g.
    # a step of 5/8 puts every 8th voxel on x.5) MONAI's own answer is decided by the float64 rounding noise of its
    # normalise / unnormalise round trip, K1's by its fp32 chain: neither is "the" reference there.
    tie = torch.zeros(got_m.shape[1:], dtype=torch.bool)
    for a in range(3):
        u = np.arange(got_m.shape[1 + a], dtype=np.float64) * (pixdim[a] / spacing[a])
        near = torch.from_numpy(np.abs(u - np.floor(u) - 0.5) < 1e-4)
        tie |= near.reshape([-1 if b == a else 1 for b in range(3)])
    assert int(((got_m != want_m)[0] & ~tie).sum()) == 0
    assert np.allclose(out["image_spacing"], pixdim)


def test_unchanged_spacing_hands_the_entry_on_and_oblique_affines_are_refused(dev):
    img = torch.rand(1, 16, 16, 8)
    aff = np.diag([0.5, 0.5, 3.0, 1.0])
    d = T.Spacingd(["image"], pixdim=(0.5, 0.5, 3.0))({"image": img.to(dev), "image_meta_dict": {"affine": aff}})
    assert torch.equal(T.as_pending(d["image"]).tensor().cpu(), img)
    aff[0, 1] = 0.1
    with pytest.raises(NotImplementedError):
        T.Spacingd(["image"], pixdim=(1.0, 1.0, 3.0))({"image": img.to(dev), "image_meta_dict": {"affine": aff}})


def test_factories_place_spacingd_like_the_reference():
    seg = F.SegmentationTransforms(["t2", "mask"], ["t2"], ["mask"], ["t2"], [], target_spacing=[0.5, 0.5, 3.0],
                                   intp_resampling_augmentations=["bilinear", "nearest"]).pre_transforms()
    assert isinstance(seg[0], T.Spacingd) and seg[0].mode == ["bilinear", "nearest"]       # before the intensity scalers
    cls = F.ClassificationTransforms(["t2"], adc_keys=[], target_spacing=[0.5, 0.5, 3.0], mask_key="mask").pre_transforms()
    kinds = [type(x).__name__ for x in cls]
    assert kinds.index("Spacingd") > kinds.index("ScaleIntensityd")                          # after them here
    assert cls[kinds.index("Spacingd")].mode == ["bilinear", "nearest"]
    ssl = F.SSLTransforms(["image"], ["image_copy"], adc_keys=[], non_adc_keys=["image"], target_spacing=[1, 1, 3]).pre_transforms()
    assert isinstance(ssl[0], T.Spacingd)
