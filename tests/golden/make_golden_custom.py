"""Generates tests/golden/custom_transforms.npz by running the REFERENCE's own custom transforms
(/root/reference/adell_mri/utils/monai_transforms/{image_intensity_ops,generic_data_ops,image_ops}.py:
ConditionalRescalingd, Offsetd, CopyEntryd, AdjustSizesd, GetAllCropsd) in the build container.

These classes are the reference's code, not MONAI's; they only subclass `monai.transforms.Transform`
/ `MapTransform` as markers.  MONAI and SimpleITK are not installed, so the three modules are loaded
from their files with two-line stand-ins for those base classes (no behaviour is supplied by the
stand-ins: every method that runs is the reference's).

    python tests/golden/make_golden_custom.py        # needs /root/reference; not run on the GPU box
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = "/root/reference/adell_mri"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    monai = types.ModuleType("monai")
    tr = types.ModuleType("monai.transforms")

    class Transform:  # marker base classes only
        pass

    class MapTransform(Transform):
        pass

    tr.Transform, tr.MapTransform, tr.InvertibleTransform, tr.RandomizableTransform = Transform, MapTransform, Transform, Transform
    monai.transforms = tr
    sys.modules["monai"], sys.modules["monai.transforms"] = monai, tr
    sys.modules["SimpleITK"] = types.ModuleType("SimpleITK")
    for name in ("adell_mri", "adell_mri.utils", "adell_mri.utils.monai_transforms"):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    mods = {}
    for name, path in (("adell_mri.custom_types", "custom_types.py"),
                       ("ref_intensity", "utils/monai_transforms/image_intensity_ops.py"),
                       ("ref_generic", "utils/monai_transforms/generic_data_ops.py"),
                       ("ref_image", "utils/monai_transforms/image_ops.py")):
        spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, path))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods["ref_intensity"], mods["ref_generic"], mods["ref_image"]


def volume(seed, shape, scale=1.0, offset=0.0):
    return (np.random.RandomState(seed).rand(*shape).astype(np.float32) * np.float32(scale) + np.float32(offset)).astype(np.float32)


#: (name, shapes per key) of the AdjustSizesd cases: odd and even size differences on every axis
ADJUST_CASES = [("even_odd", {"a": (1, 10, 9, 8), "b": (1, 5, 9, 6), "c": (2, 7, 12, 8)}),
                ("equal", {"a": (1, 6, 6, 4), "b": (1, 6, 6, 4)})]
CROP_CASES = [((2, 70, 61, 20), (32, 32, 8)), ((1, 64, 45, 17), (16, 16, 8)), ((1, 40, 36, 16), (32, 32, 8))]


def _random_cases(seed=21):
    """Seeded random shapes: size differences of every parity for AdjustSizesd, volumes smaller / larger than the
    crop size and non-multiples of it for GetAllCropsd."""
    R = np.random.RandomState(seed)
    adjust, crops = [], []
    for i in range(8):
        keys = ["a", "b", "c"][: int(R.randint(2, 4))]
        adjust.append((f"random_{i}", {k: (int(R.randint(1, 3)), int(R.randint(4, 14)), int(R.randint(4, 14)), int(R.randint(3, 10))) for k in keys}))
    for i in range(6):
        size = (int(R.choice([8, 12, 16])), int(R.choice([8, 12, 16])), int(R.choice([4, 6, 8])))
        shape = (int(R.randint(1, 3)), int(R.randint(5, 40)), int(R.randint(5, 40)), int(R.randint(3, 20)))
        crops.append((shape, size))
    return adjust, crops


_adj, _crp = _random_cases()
ADJUST_CASES += _adj
CROP_CASES += _crp


def ramp(shape, mod=251):
    """Position-coded volume (compresses well, every voxel of a crop identifies its origin)."""
    return (np.arange(int(np.prod(shape)), dtype=np.int64) % mod).astype(np.float32).reshape(shape)


def expose_ops():
    return [{"class": "SpatialPad", "extra_info": {"padded": [[0, 0], [1, 1]]}},
            {"class": "RandSpatialCrop", "extra_info": {"cropped": [3, 5, 0, 8, 2, 2]}},
            {"class": "RandFlip", "extra_info": {}},
            {"class": "RandSpatialCrop", "extra_info": {"cropped": [1, 4, 2, 6, 0, 3]}}]


def main():
    I, Gm, Im = load_reference()
    out = {}
    # ConditionalRescalingd(500, 0.001): one key above the threshold, one below (transforms.py:143-155)
    hi, lo = volume(1, (1, 12, 10, 6), 2000.0), volume(2, (1, 12, 10, 6), 400.0)
    d = I.ConditionalRescalingd(["hi", "lo"], 500, 0.001)({"hi": torch.from_numpy(hi.copy()), "lo": torch.from_numpy(lo.copy())})
    out["condrescale/hi"], out["condrescale/lo"] = d["hi"].numpy(), d["lo"].numpy()
    # Offsetd(None) (classification ADC branch) and a fixed offset
    x = volume(3, (1, 9, 8, 7), 3.0, 0.25)
    out["offset/min"] = I.Offsetd(["x"], None)({"x": torch.from_numpy(x.copy())})["x"].numpy()
    out["offset/fixed"] = I.Offsetd(["x"], 0.75)({"x": torch.from_numpy(x.copy())})["x"].numpy()
    # CopyEntryd
    d = Gm.CopyEntryd(["x"], {"x": "x_copy"})({"x": torch.from_numpy(x.copy()), "other": 3})
    out["copy/x_copy"] = d["x_copy"].numpy()
    out["copy/keys"] = np.array(sorted(d.keys()))
    # ExposeTransformKeyMetad on an object that carries MONAI-style applied_operations (two crops: the last wins)
    ops = expose_ops()
    holder = types.SimpleNamespace(applied_operations=ops)
    d = Gm.ExposeTransformKeyMetad("image", "RandSpatialCrop", ["extra_info", "cropped"], "box_1")({"image": holder})
    out["expose/named"] = np.asarray(d["box_1"])
    d = Gm.ExposeTransformKeyMetad("image", "RandSpatialCrop", ["extra_info", "cropped"])({"image": holder})
    out["expose/default_keys"] = np.array(sorted(d.keys()))
    out["expose/default_value"] = np.asarray(d["box_image"])
    # AdjustSizesd, both modes
    for name, shapes in ADJUST_CASES:
        for mode in ("crop", "pad"):
            data = {k: torch.from_numpy(volume(10 + i, s)) for i, (k, s) in enumerate(shapes.items())}
            d = Im.AdjustSizesd(list(shapes), ndim=3, mode=mode)(data)
            for k in shapes:
                out[f"adjust/{name}/{mode}/{k}"] = d[k].numpy()
    # GetAllCropsd
    for ci, (shape, size) in enumerate(CROP_CASES):
        img, mask = ramp(shape), (ramp((1, *shape[1:]), 7) > 3).astype(np.float32)
        crops = Im.GetAllCropsd(["image", "mask"], size)({"image": torch.from_numpy(img), "mask": torch.from_numpy(mask), "id": 7})
        out[f"crops/{ci}/n"] = np.array(len(crops))
        if crops:
            out[f"crops/{ci}/image"] = np.stack([np.asarray(c["image"]) for c in crops])
            out[f"crops/{ci}/mask"] = np.stack([np.asarray(c["mask"]) for c in crops])
    np.savez_compressed(os.path.join(HERE, "custom_transforms.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
