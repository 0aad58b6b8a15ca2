"""The reference's own custom transforms (ConditionalRescalingd, Offsetd, CopyEntryd, AdjustSizesd,
GetAllCropsd) — golden vectors produced by RUNNING THE REFERENCE CODE in the build container
(tests/golden/make_golden_custom.py), so these rows are pinned against the reference itself.
CPU: integer transforms through the C restatement + the oracle's intensity restatements.
GPU (marked): everything through the CUDA path."""

import importlib.util
import os

import numpy as np
import pytest
import torch

from adell_mri_b200 import collate, engine, transform_factory as F, transforms as T
from oracle import monai_restated as M
from tests.helpers import cref_execute, patch_engine_for_cpu

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden_custom", os.path.join(HERE, "golden", "make_golden_custom.py"))
G = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(G)
GOLD = np.load(os.path.join(HERE, "golden", "custom_transforms.npz"))


@pytest.fixture(params=["cpu", pytest.param("cuda:0", marks=pytest.mark.gpu)])
def dev(request, monkeypatch):
    if request.param == "cpu":
        patch_engine_for_cpu(monkeypatch)
    return request.param


def _mat(x):
    return (x.tensor() if isinstance(x, T.Pending) else x).cpu().numpy()


@pytest.mark.parametrize("name,shapes", G.ADJUST_CASES, ids=[c[0] for c in G.ADJUST_CASES])
@pytest.mark.parametrize("mode", ["crop", "pad"])
def test_adjust_sizes(dev, name, shapes, mode):
    data = {k: torch.from_numpy(G.volume(10 + i, s)).to(dev) for i, (k, s) in enumerate(shapes.items())}
    d = F.AdjustSizesd(list(shapes), ndim=3, mode=mode)(data)
    for k in shapes:
        assert np.array_equal(_mat(d[k]), GOLD[f"adjust/{name}/{mode}/{k}"]), (k, mode)


@pytest.mark.parametrize("ci", range(len(G.CROP_CASES)))
def test_get_all_crops(dev, ci):
    shape, size = G.CROP_CASES[ci]
    img, mask = G.ramp(shape), (G.ramp((1, *shape[1:]), 7) > 3).astype(np.float32)
    crops = F.GetAllCropsd(["image", "mask"], size)({"image": torch.from_numpy(img).to(dev), "mask": torch.from_numpy(mask).to(dev), "id": 7})
    assert len(crops) == int(GOLD[f"crops/{ci}/n"])
    if crops:
        got = collate.safe_collate_crops([crops])
        assert np.array_equal(got["image"].cpu().numpy(), GOLD[f"crops/{ci}/image"])
        assert np.array_equal(got["mask"].cpu().numpy(), GOLD[f"crops/{ci}/mask"])


def test_copy_entry(dev):
    x = torch.from_numpy(G.volume(3, (1, 9, 8, 7), 3.0, 0.25)).to(dev)
    d = F.CopyEntryd(["x"], {"x": "x_copy"})({"x": x, "other": 3})
    assert sorted(d.keys()) == list(GOLD["copy/keys"])
    assert np.array_equal(_mat(d["x_copy"]), GOLD["copy/x_copy"])


def test_expose_transform_key_meta():
    p = T.Pending(torch.zeros(1, 4, 4, 4))
    p.meta["applied_operations"] = G.expose_ops()
    d = T.ExposeTransformKeyMetad("image", "RandSpatialCrop", ["extra_info", "cropped"], "box_1")({"image": p})
    assert np.array_equal(np.asarray(d["box_1"]), GOLD["expose/named"])
    d = T.ExposeTransformKeyMetad("image", "RandSpatialCrop", ["extra_info", "cropped"])({"image": p})
    assert sorted(d.keys()) == list(GOLD["expose/default_keys"])
    assert np.array_equal(np.asarray(d["box_image"]), GOLD["expose/default_value"])


def test_oracle_intensity_restatements_equal_reference_vectors():
    hi, lo = G.volume(1, (1, 12, 10, 6), 2000.0), G.volume(2, (1, 12, 10, 6), 400.0)
    assert np.array_equal(M.conditional_rescaling(torch.from_numpy(hi), 500, 0.001).numpy(), GOLD["condrescale/hi"])
    assert np.array_equal(M.conditional_rescaling(torch.from_numpy(lo), 500, 0.001).numpy(), GOLD["condrescale/lo"])
    x = G.volume(3, (1, 9, 8, 7), 3.0, 0.25)
    assert np.array_equal(M.offset(torch.from_numpy(x), None).numpy(), GOLD["offset/min"])
    assert np.array_equal(M.offset(torch.from_numpy(x), 0.75).numpy(), GOLD["offset/fixed"])


@pytest.mark.gpu
def test_device_intensity_transforms_equal_reference_vectors():
    hi, lo = G.volume(1, (1, 12, 10, 6), 2000.0), G.volume(2, (1, 12, 10, 6), 400.0)
    d = F.ConditionalRescalingd(["hi", "lo"], 500, 0.001)({"hi": torch.from_numpy(hi).cuda(), "lo": torch.from_numpy(lo).cuda()})
    assert np.array_equal(_mat(d["hi"]), GOLD["condrescale/hi"]) and np.array_equal(_mat(d["lo"]), GOLD["condrescale/lo"])
    x = G.volume(3, (1, 9, 8, 7), 3.0, 0.25)
    assert np.array_equal(_mat(F.Offsetd(["x"], None)({"x": torch.from_numpy(x).cuda()})["x"]), GOLD["offset/min"])
    assert np.array_equal(_mat(F.Offsetd(["x"], 0.75)({"x": torch.from_numpy(x).cuda()})["x"]), GOLD["offset/fixed"])
