"""GetAllCrops / GetAllCropsd (validation tiling): the reference's own tests
(/root/reference/testing/test_get_all_crops.py:12-33 — crop counts and the collated batch size) on the
lazy surface, plus the crop CONTENT against a literal numpy restatement of the reference loops
(/root/reference/adell_mri/utils/monai_transforms/image_ops.py:257-331).  CPU: plans run through the C
restatement; GPU (marked): through CUDA box copies."""

import numpy as np
import pytest
import torch

from adell_mri_b200 import collate, engine, transform_factory as F, transforms as T
from tests.helpers import cref_execute, patch_engine_for_cpu

input_tensor_size = np.array([1, 128, 128, 16])
crop_size = np.array([32, 32, 8])
n_crops = np.prod(input_tensor_size[1:] / crop_size)


@pytest.fixture(params=["cpu", pytest.param("cuda:0", marks=pytest.mark.gpu)])
def dev(request, monkeypatch):
    if request.param == "cpu":
        patch_engine_for_cpu(monkeypatch)
    return request.param


def reference_crops(X: np.ndarray, size):
    """Literal restatement of GetAllCrops.get_all_crops_3d (pad, loops over the un-padded extent)."""
    sh = list(X.shape[1:])
    rem = [(y - (x % y)) if x > y else 0 for x, y in zip(X.shape, size)]   # full shape: the reference's own quirk
    rem = [x if x < (y // 2) else 0 for x, y in zip(rem, size)]
    Xp = np.pad(X, [(0, 0), *[(x // 2, x - x // 2) for x in rem]], "constant", constant_values=0)
    out = []
    for i in range(0, sh[0], size[0]):
        for j in range(0, sh[1], size[1]):
            for k in range(0, sh[2], size[2]):
                if i + size[0] < sh[0] + 1 and j + size[1] < sh[1] + 1 and k + size[2] < sh[2] + 1:
                    out.append(Xp[:, i:i + size[0], j:j + size[1], k:k + size[2]])
    return out


def test_gac():
    gac = F.GetAllCrops(size=crop_size)
    assert len(gac(torch.zeros(input_tensor_size.tolist()))) == n_crops


def test_gacd():
    gac = F.GetAllCropsd(keys=["image"], size=crop_size)
    assert len(gac({"image": torch.zeros(input_tensor_size.tolist())})) == n_crops


def test_gacd_and_collate(dev):
    gac = F.GetAllCropsd(keys=["image"], size=crop_size)
    out = collate.safe_collate_crops([gac({"image": torch.zeros(input_tensor_size.tolist(), device=dev)})])
    assert out["image"].shape[0] == n_crops
    assert float(out["image"].abs().max()) == 0.0


@pytest.mark.parametrize("shape,size", [((2, 70, 61, 20), (32, 32, 8)), ((1, 64, 45, 17), (16, 16, 8)), ((1, 30, 30, 6), (32, 32, 8))])
def test_crop_content_matches_reference_loops(dev, shape, size):
    R = np.random.RandomState(0)
    img = R.rand(*shape).astype(np.float32)
    mask = (R.rand(1, *shape[1:]) > 0.6).astype(np.float32)
    want_i, want_m = reference_crops(img, size), reference_crops(mask, size)
    crops = F.GetAllCropsd(["image", "mask"], size)({"image": torch.from_numpy(img).to(dev), "mask": torch.from_numpy(mask).to(dev), "id": 7})
    assert len(crops) == len(want_i)
    if not crops:
        return
    assert all(c["id"] == 7 for c in crops)
    got = collate.safe_collate_crops([crops])
    assert got["image"].shape == (len(want_i), shape[0], *size)
    for b in range(len(want_i)):
        assert np.array_equal(got["image"][b].cpu().numpy(), want_i[b])
        assert np.array_equal(got["mask"][b].cpu().numpy(), want_m[b])
