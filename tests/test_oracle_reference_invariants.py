"""The invariants the reference's own tests pin for this path, restated on the oracle
(/root/reference/testing/test_crop_from_maskd.py:166-374 — classification pipeline shapes,
[0,1] range, binary mask channel, crop-with-margin then final crop;
/root/reference/adell_mri/transform_factory/augmentations.py:402-406 — flatten_box)."""

import numpy as np
import torch

from oracle import monai_restated as M


def _classification_chain(img, mask, crop):
    """pre_transforms()[1:] + post_transforms() of ClassificationTransforms without file IO:
    ScaleIntensityd(0,1) -> SpatialPadd(crop+16) -> CenterSpatialCropd(crop+16) -> CenterSpatialCropd(crop)
    -> ConcatItemsd(image, mask) (/root/reference/adell_mri/transform_factory/transforms.py:399-509)."""
    margin = [c + 16 for c in crop]
    outs = []
    for x, scale in ((img, True), (mask, False)):
        if scale:
            x = M.scale_intensity(x, 0.0, 1.0)
        x = M.spatial_pad(x, margin)
        x = M.center_spatial_crop(x, margin)
        x = M.center_spatial_crop(x, crop)
        outs.append(x)
    return torch.cat(outs, 0)


def test_classification_pipeline_shapes_ranges_and_binary_mask():
    R = np.random.RandomState(0)
    for shape in [(32, 32, 32), (20, 40, 24), (64, 48, 16)]:
        img = torch.from_numpy((R.rand(1, *shape) * 1000).astype(np.float32))
        mask = torch.zeros(1, *shape)
        mask[:, 4:12, 4:12, 4:12] = 1
        out = _classification_chain(img, mask, (16, 16, 16))
        assert tuple(out.shape) == (2, 16, 16, 16)
        assert float(out[0].min()) >= 0.0 and float(out[0].max()) <= 1.0
        assert set(np.unique(out[1].numpy()).tolist()) <= {0.0, 1.0}


def test_scaling_ratio_is_constant():
    """inverse-ratio invariant of test_crop_from_maskd.py:307-374: min-max scaling is one affine map."""
    R = np.random.RandomState(1)
    x = torch.from_numpy((R.rand(1, 12, 12, 12) * 37 + 5).astype(np.float32))
    y = M.scale_intensity(x, 0.0, 1.0)
    ratio = (x - x.min()) / torch.clamp(y, min=1e-12)
    sel = y > 1e-3
    assert float(ratio[sel].max() - ratio[sel].min()) < 1e-3 * float(ratio[sel].mean())


def test_center_crop_and_pad_index_conventions():
    x = torch.arange(7 * 6 * 5, dtype=torch.float32).reshape(1, 7, 6, 5)
    assert torch.equal(M.center_spatial_crop(x, (4, 3, 2)), x[:, 1:5, 2:5, 1:3])       # start = size//2 - roi//2
    p = M.spatial_pad(x, (10, 7, 5))
    assert tuple(p.shape) == (1, 10, 7, 5) and torch.equal(p[:, 1:8, 0:6, :], x)      # before = w//2
    assert float(p[:, 0].abs().sum()) == 0 and float(p[:, 8:].abs().sum()) == 0


def test_flatten_box_formula():
    starts, sizes, roi = [3, 5, 1], [16, 16, 8], (16, 16, 8)
    cropped = M.cropped_extra_info((32, 32, 12), starts, sizes)
    assert cropped == [3, 13, 5, 11, 1, 3]
    box = M.flatten_box(cropped, roi)
    assert box.dtype == np.float32 and box.tolist() == [3, 5, 1, 3, 5, 5]


def test_intensity_token_known_answers():
    """--augment intensity members († MONAI AdjustContrast / StdShiftIntensity): closed forms on a ramp."""
    import numpy as np
    import torch

    from oracle import monai_restated as M

    x = torch.arange(5, dtype=torch.float32).reshape(1, 5, 1, 1) + 2.0          # 2..6: min 2, range 4
    y = M.adjust_contrast(x, 2.0).reshape(-1).numpy()
    want = ((np.arange(5, dtype=np.float32) / np.float32(np.float32(4) + np.float32(1e-7))) ** 2 * np.float32(4) + np.float32(2))
    assert np.allclose(y, want, rtol=1e-6) and y[0] == 2.0 and abs(y[-1] - 6.0) < 1e-5
    z = M.std_shift_intensity(x, 0.5).reshape(-1).numpy()
    assert np.allclose(z, (np.arange(5) + 2.0) + 0.5 * np.sqrt(2.0), rtol=1e-6)   # population std of 5 consecutive ints = sqrt(2)
    c = M.std_shift_intensity(torch.full((1, 3, 3, 3), 7.0), 0.5)
    assert torch.equal(c, torch.full((1, 3, 3, 3), 7.0))                            # zero std: no shift
