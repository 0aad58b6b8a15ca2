"""Generates tests/golden/crop_from_mask.json by running the REFERENCE's own `CropFromMask` /
`CropFromMaskd` (/root/reference/adell_mri/utils/monai_transforms/labels.py:412-522) on the fixtures of
the reference's test file (/root/reference/testing/test_crop_from_maskd.py:63-164: arange volumes with
box masks).  MONAI / scikit-image are not installed: `Crop` / `Cropd` are replaced by the two-line
slicing they perform (`img[:, *slices]` per key); `compute_slices` / `get_centre_extremes`, i.e. all
the window logic, are the reference's code.  Recorded: the slices the reference computes per case.

    python tests/golden/make_golden_cropmask.py        # needs /root/reference; not run on the GPU box
"""
import importlib.util
import json
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/adell_mri/utils/monai_transforms/labels.py"

#: name -> (spatial shape, mask box slices or None for an empty mask, output_size)
CASES = {
    "mask_driven": ((32, 32, 32), [(3, 13), (1, 15), (4, 7)], None),                 # test_3d_mask_driven_content
    "fixed_size": ((64, 64, 64), [(24, 40), (24, 40), (24, 40)], [16, 16, 16]),      # test_3d_fixed_size_content
    "fixed_size_clamped": ((32, 32, 32), [(0, 4), (0, 4), (0, 4)], [16, 16, 16]),    # test_3d_fixed_size_clamped_content
    "empty_mask_centre": ((32, 32, 32), None, [16, 16, 16]),                         # test_3d_empty_mask_centre_crop_content
    "multi_key": ((32, 32, 32), [(10, 22), (10, 22), (10, 22)], [12, 12, 12]),       # test_multi_key_content
    "clamped_high_odd": ((30, 29, 17), [(25, 30), (20, 29), (15, 17)], [9, 11, 5]),
    "classification_margin": ((64, 64, 64), [(20, 36), (20, 36), (20, 36)], [32, 32, 32]),  # crop_size 16 + 16 margin
}


def _random_cases(n=40, seed=11):
    """Seeded random windows: boxes touching the borders, single voxels, odd sizes, output sizes smaller and larger
    than the box, empty masks."""
    import numpy as np
    R = np.random.RandomState(seed)
    out = {}
    for i in range(n):
        spatial = tuple(int(R.randint(9, 48)) for _ in range(3))
        if R.rand() < 0.12:
            box = None
        else:
            box = []
            for s in spatial:
                a = int(R.randint(0, s)); b = int(R.randint(a + 1, s + 1))
                if R.rand() < 0.2: a = 0
                if R.rand() < 0.2: b = s
                box.append((a, b))
        osize = None if (box is not None and R.rand() < 0.3) else [int(R.randint(1, s + 1)) for s in spatial]
        out[f"random_{i:02d}"] = (spatial, box, osize)
    return out


CASES.update(_random_cases())


def load_reference():
    class Crop:
        def __init__(self, *a, **k):
            pass

        def __call__(self, img, slices, lazy=False):
            return img[(slice(None), *slices)]

    class Cropd:
        def __init__(self, keys, cropper, *a, **k):
            self.keys, self.cropper = keys, cropper

    class Base:
        def __init__(self, *a, **k):
            pass

    monai = types.ModuleType("monai")
    for name in ("monai.data", "monai.data.meta_tensor", "monai.transforms", "monai.transforms.inverse", "monai.utils",
                 "skimage", "skimage.morphology", "adell_mri", "adell_mri.utils"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["monai"] = monai
    monai.transforms = sys.modules["monai.transforms"]
    sys.modules["monai.data.meta_tensor"].MetaTensor = torch.Tensor
    tr = sys.modules["monai.transforms"]
    tr.Crop, tr.Cropd, tr.MapTransform, tr.TraceableTransform, tr.Transform = Crop, Cropd, Base, Base, Base
    sys.modules["monai.transforms.inverse"].InvertibleTransform = Base
    monai.utils = sys.modules["monai.utils"]
    monai.data = sys.modules["monai.data"]
    monai.utils.TraceKeys = object
    monai.utils.ensure_tuple = tuple
    monai.utils.TransformBackends = types.SimpleNamespace(NUMPY="numpy", TORCH="torch")
    sys.modules["skimage.morphology"].convex_hull_image = None
    spec = importlib.util.spec_from_file_location("adell_mri.custom_types", "/root/reference/adell_mri/custom_types.py")
    ct = importlib.util.module_from_spec(spec)
    sys.modules["adell_mri.custom_types"] = ct
    spec.loader.exec_module(ct)
    spec = importlib.util.spec_from_file_location("ref_labels", SRC)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_sample(spatial, box, n_channels=1):
    n = n_channels * int(torch.tensor(spatial).prod())
    img = torch.arange(n, dtype=torch.float32).reshape(n_channels, *spatial)
    mask = torch.zeros(1, *spatial)
    if box is not None:
        mask[(slice(None), *[slice(a, b) for a, b in box])] = 1.0
    return img, mask


def label_inputs():
    """Three label maps with values in {0, 1, 2, 3} (uint8-valued floats), deterministic."""
    import numpy as np
    R = np.random.RandomState(4)
    return [torch.from_numpy((R.rand(1, 9, 8, 7) * 4).astype(np.int64).astype(np.float32) * (R.rand(1, 9, 8, 7) > 0.5)) for _ in range(3)]


LABEL_CASES = {  # name -> (combine mode, possible_labels, label mode, positive_labels)
    "any_binary": ("any", [0, 1], "binary", [1]),
    "majority_binary": ("majority", [0, 1], "binary", [1]),
    "any_cat": ("any", [0, 1], "cat", [1]),
}
OPERATOR_CASES = {  # applied to the first raw map alone
    "binary_2_3": ([0, 1, 2, 3], "binary", [2, 3]),
    "cat_reordered": ([3, 1, 2], "cat", [1]),
    "passthrough": ([0, 1], None, [1]),
}


def main():
    L = load_reference()
    out = {}
    import numpy as np
    lab = {}
    for name, (comb, possible, mode, positive) in LABEL_CASES.items():
        a, b, c = label_inputs()
        d = L.CombineBinaryLabelsd(["a", "b", "c"], comb, "mask")({"a": (a > 0).float(), "b": (b > 0).float(), "c": (c > 0).float()})
        d = L.LabelOperatorSegmentationd(["mask"], possible, mode=mode, positive_labels=positive)(d)
        lab[name] = np.asarray(d["mask"], np.float32)
    for name, (possible, mode, positive) in OPERATOR_CASES.items():
        a, _, _ = label_inputs()
        d = L.LabelOperatorSegmentationd(["a"], possible, mode=mode, positive_labels=positive)({"a": a.numpy()})
        lab["op_" + name] = np.asarray(d["a"], np.float32)
    np.savez_compressed(os.path.join(HERE, "label_ops.npz"), **lab)
    for name, (spatial, box, osize) in CASES.items():
        img, mask = make_sample(spatial, box)
        t = L.CropFromMaskd(keys=["image", "mask"], mask_key="mask", output_size=osize)
        d = t({"image": img, "mask": mask})
        sl = t.cropper.compute_slices(img, mask)
        out[name] = {"slices": [[int(s.start), int(s.stop)] for s in sl], "image_shape": list(d["image"].shape),
                     "image_sum": float(d["image"].double().sum()), "mask_sum": float(d["mask"].sum())}
    json.dump(out, open(os.path.join(HERE, "crop_from_mask.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(out)[:400])


if __name__ == "__main__":
    main()
