"""CropFromMaskd (/root/reference/adell_mri/utils/monai_transforms/labels.py:412-522).
GPU only (the bounding box is a device reduction).  Two layers:
* the windows must equal those the reference's own `compute_slices` produced for the fixtures of the
  reference's test file (tests/golden/crop_from_mask.json, made by tests/golden/make_golden_cropmask.py);
* the reference's tests themselves, mirrored (/root/reference/testing/test_crop_from_maskd.py:63-164,
  230-262): arange volumes with box masks, crop + inverse must preserve the content of the window; the
  classification factory with image_crop_from_mask=True yields (2, 16, 16, 16) with a binary mask channel."""

import importlib.util
import json
import os

import pytest
import torch

from adell_mri_b200 import collate, transform_factory as F, transforms as T

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden_cropmask", os.path.join(HERE, "golden", "make_golden_cropmask.py"))
G = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(G)
GOLD = json.load(open(os.path.join(HERE, "golden", "crop_from_mask.json")))
DEV = "cuda:0"


def _mat(x):
    return (x.tensor() if isinstance(x, T.Pending) else x).cpu()


@pytest.mark.parametrize("name", sorted(G.CASES))
def test_windows_equal_the_reference_compute_slices(name):
    spatial, box, osize = G.CASES[name]
    img, mask = G.make_sample(spatial, box)
    t = F.CropFromMaskd(keys=["image", "mask"], mask_key="mask", output_size=osize)
    sl = t.cropper.compute_slices((1, *spatial), mask.to(DEV))
    assert [[s.start, s.stop] for s in sl] == GOLD[name]["slices"]
    d = t({"image": img.to(DEV), "mask": mask.to(DEV)})
    got_img, got_mask = _mat(d["image"]), _mat(d["mask"])
    want = img[(slice(None), *[slice(a, b) for a, b in GOLD[name]["slices"]])]
    assert list(got_img.shape) == GOLD[name]["image_shape"] and torch.equal(got_img, want)
    assert float(got_img.double().sum()) == GOLD[name]["image_sum"] and float(got_mask.sum()) == GOLD[name]["mask_sum"]


def _assert_content_preserved(restored, original, mask_slices):
    s = (slice(None), *mask_slices)
    assert restored.shape == original.shape
    assert torch.allclose(restored[s], original[s])


@pytest.mark.parametrize("spatial,box,osize", [
    ((32, 32, 32), [(3, 13), (1, 15), (4, 7)], None),
    ((64, 64, 64), [(24, 40), (24, 40), (24, 40)], [16, 16, 16]),
    ((32, 32, 32), [(0, 4), (0, 4), (0, 4)], [16, 16, 16]),
])
def test_crop_then_inverse_preserves_window_content(spatial, box, osize):
    img, mask = G.make_sample(spatial, box)
    t = F.CropFromMaskd(keys=["image"], mask_key="mask", output_size=osize)
    restored = _mat(t.inverse(t({"image": img.to(DEV), "mask": mask.to(DEV)}))["image"])
    _assert_content_preserved(restored, img, [slice(a, b) for a, b in box])


def test_empty_mask_centre_crop_content():
    img, mask = G.make_sample((32, 32, 32), None)
    t = F.CropFromMaskd(keys=["image"], mask_key="mask", output_size=[16, 16, 16])
    restored = _mat(t.inverse(t({"image": img.to(DEV), "mask": mask.to(DEV)}))["image"])
    _assert_content_preserved(restored, img, [slice(8, 24)] * 3)


def test_multi_key_content():
    box = [(10, 22)] * 3
    img1, mask = G.make_sample((32, 32, 32), box)
    img2, _ = G.make_sample((32, 32, 32), box, n_channels=2)
    t = F.CropFromMaskd(keys=["img1", "img2"], mask_key="mask", output_size=[12, 12, 12])
    restored = t.inverse(t({"img1": img1.to(DEV), "img2": img2.to(DEV), "mask": mask.to(DEV)}))
    _assert_content_preserved(_mat(restored["img1"]), img1, [slice(a, b) for a, b in box])
    _assert_content_preserved(_mat(restored["img2"]), img2, [slice(a, b) for a, b in box])


def test_classification_transforms_mask_crop():
    """test_crop_from_maskd.py:230-262: ClassificationTransforms(image_crop_from_mask=True)."""
    img, mask = G.make_sample((64, 64, 64), [(20, 36)] * 3)
    tf = F.ClassificationTransforms(keys=("image",), adc_keys=(), clinical_feature_keys=(), target_spacing=None,
                                    crop_size=(16, 16, 16), pad_size=None, image_crop_from_mask=True, mask_key="mask")
    final = T.Compose([*tf.pre_transforms(), *tf.post_transforms()])({"image": img.to(DEV), "mask": mask.to(DEV)})
    restored = collate.safe_collate([final])["image"][0].cpu()
    assert restored.shape == (2, 16, 16, 16)
    assert restored[0].min() >= 0.0 and restored[0].max() <= 1.0
    assert torch.all((restored[1] == 0) | (restored[1] == 1))


# ------------------------------------------------------------------ label construction (labels.py:123-220)
import numpy as np

LAB = np.load(os.path.join(HERE, "golden", "label_ops.npz"))


@pytest.mark.parametrize("name", sorted(G.LABEL_CASES))
def test_combine_and_label_operator_equal_reference_classes(name):
    comb, possible, mode, positive = G.LABEL_CASES[name]
    a, b, c = G.label_inputs()
    d = {"a": (a > 0).float().to(DEV), "b": (b > 0).float().to(DEV), "c": (c > 0).float().to(DEV)}
    d = F.CombineBinaryLabelsd(["a", "b", "c"], comb, "mask")(d)
    d = F.LabelOperatorSegmentationd(["mask"], possible, mode=mode, positive_labels=positive)(d)
    assert np.array_equal(d["mask"].cpu().numpy(), LAB[name])


@pytest.mark.parametrize("name", sorted(G.OPERATOR_CASES))
def test_label_operator_equals_reference_class(name):
    possible, mode, positive = G.OPERATOR_CASES[name]
    a, _, _ = G.label_inputs()
    d = F.LabelOperatorSegmentationd(["a"], possible, mode=mode, positive_labels=positive)({"a": a.to(DEV)})
    assert np.array_equal(d["a"].cpu().numpy(), LAB["op_" + name])


def test_segmentation_factory_builds_the_mask_from_label_keys():
    """transforms.py:181-194: two label maps -> any -> binary "mask", then the image / mask batch."""
    R = np.random.RandomState(3)
    s = {"t2": torch.from_numpy(R.rand(1, 16, 12, 8).astype(np.float32) * 900).to(DEV),
         "l1": torch.from_numpy((R.rand(1, 16, 12, 8) > 0.8).astype(np.float32)).to(DEV),
         "l2": torch.from_numpy((R.rand(1, 16, 12, 8) > 0.8).astype(np.float32)).to(DEV)}
    tf = F.SegmentationTransforms(all_keys=["t2", "l1", "l2"], image_keys=["t2"], label_keys=["l1", "l2"], non_adc_keys=["t2"],
                                  adc_keys=[], possible_labels=[0, 1], positive_labels=[1], label_mode="binary")
    out = collate.safe_collate([tf.transforms()(s)])
    want = ((s["l1"] + s["l2"]) > 0).float()
    assert sorted(out.keys()) == ["image", "mask"]
    assert torch.equal(out["mask"][0], want)
    assert float(out["image"].min()) == 0.0 and float(out["image"].max()) == 1.0
