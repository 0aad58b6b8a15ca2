"""Generates tests/golden/builder_wiring.json by RUNNING THE REFERENCE's own augmentation builders
(/root/reference/adell_mri/transform_factory/augmentations.py: get_augmentations_unet / _class / _ssl,
/root/reference/adell_mri/modules/augmentations.py: AugmentationWorkhorsed, get_transform_d) in the
build container.  MONAI is not installed: `monai.transforms` is replaced by a RECORDER whose classes
only remember their name and constructor arguments, so what is captured is exactly which MONAI
transform the reference constructs, with which arguments, in which order — the wiring rows a7-a10 of
SURVEY.md §8 — produced by the reference's code, not restated.  The nested `flatten_box` of
get_augmentations_ssl is captured as the live function object and evaluated on sample boxes.

    python tests/golden/make_golden_wiring.py        # needs /root/reference; not run on the GPU box
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

ROOT = "/root/reference/adell_mri"
HERE = os.path.dirname(os.path.abspath(__file__))


class Rec:
    """Stand-in for any `monai.transforms.X`: remembers the class name and constructor arguments."""

    _name = "?"

    def __init__(self, *args, **kwargs):
        self.args, self.kwargs = args, kwargs


class _Base:  # base classes the reference subclasses
    def __init__(self, *a, **k):
        pass


def _recorder_module(name):
    mod = types.ModuleType(name)
    cache = {}

    def getattr_(attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        if attr in ("Transform", "MapTransform", "RandomizableTransform", "InvertibleTransform", "Randomizable"):
            return _Base
        if attr not in cache:
            cache[attr] = type(attr, (Rec,), {"_name": attr})
        return cache[attr]

    mod.__getattr__ = getattr_
    return mod


def load_reference():
    monai = types.ModuleType("monai")
    monai.transforms = _recorder_module("monai.transforms")
    sys.modules["monai"], sys.modules["monai.transforms"] = monai, monai.transforms
    for name in ("adell_mri", "adell_mri.utils", "adell_mri.modules", "adell_mri.transform_factory"):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    sys.modules["adell_mri.utils.monai_transforms"] = _recorder_module("adell_mri.utils.monai_transforms")
    mods = {}
    for name, path in (("adell_mri.custom_types", "custom_types.py"),
                       ("adell_mri.modules.augmentations", "modules/augmentations.py"),
                       ("adell_mri.transform_factory.augmentations", "transform_factory/augmentations.py"),
                       ("adell_mri.transform_factory.transforms", "transform_factory/transforms.py")):
        spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, path))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    return (mods["adell_mri.modules.augmentations"], mods["adell_mri.transform_factory.augmentations"],
            mods["adell_mri.transform_factory.transforms"])


FLATTEN_BOX_INPUTS = [[3, 5, 0, 8, 2, 2], [0, 0, 0, 0, 0, 0], [10, 4, 7, 1, 3, 9]]


def to_json(x, workhorse_cls):
    if isinstance(x, Rec):
        return {"cls": x._name, "args": [to_json(a, workhorse_cls) for a in x.args],
                "kwargs": {k: to_json(v, workhorse_cls) for k, v in sorted(x.kwargs.items())}}
    if isinstance(x, workhorse_cls):
        return {"cls": "AugmentationWorkhorsed", "augmentations": list(x.augmentations), "keys": list(x.keys),
                "mask_keys": list(x.mask_keys), "max_mult": x.max_mult, "N": x.N,
                "transforms": {k: to_json(v, workhorse_cls) for k, v in x.transforms.items()}}
    if callable(x) and getattr(x, "__name__", "") == "flatten_box":
        return {"fn": "flatten_box", "outputs": [[float(v) for v in x(b)] for b in FLATTEN_BOX_INPUTS]}
    if isinstance(x, (list, tuple)):
        return [to_json(v, workhorse_cls) for v in x]
    if isinstance(x, dict):
        return {str(k): to_json(v, workhorse_cls) for k, v in x.items()}
    if isinstance(x, (np.floating, float)):
        return float(x)
    if isinstance(x, (np.integer, int)) and not isinstance(x, bool):
        return int(x)
    if isinstance(x, np.ndarray):
        return to_json(x.tolist(), workhorse_cls)
    if x is None or isinstance(x, (bool, str)):
        return x
    return repr(x)


KEYS, ALL = ["t2", "adc", "dwi"], ["t2", "adc", "dwi", "mask"]
UNET_CASES = {
    "affine_flip": dict(augment=["affine", "flip"], all_keys=ALL, image_keys=KEYS, t2_keys=[], flip_axis=[0, 1, 2]),
    "affine_shear_flip": dict(augment=["affine", "shear", "flip"], all_keys=ALL, image_keys=KEYS, t2_keys=[], flip_axis=[0, 1, 2]),
    "trivial": dict(augment=["trivial", "affine", "shear", "flip"], all_keys=ALL, image_keys=KEYS, t2_keys=[], flip_axis=[0, 1]),
    "intensity": dict(augment=["intensity", "affine", "flip"], all_keys=ALL, image_keys=KEYS, t2_keys=[], flip_axis=[0, 1, 2]),
    "posneg_crops": dict(augment=["affine", "flip"], all_keys=ALL, image_keys=KEYS, t2_keys=[], random_crop_size=[128, 128, 24],
                         n_crops=2, flip_axis=[0, 1, 2]),
    "random_crops_no_label": dict(augment=["affine", "shear", "flip"], all_keys=KEYS, image_keys=KEYS, t2_keys=[],
                                  random_crop_size=[64, 64, 16], has_label=False, flip_axis=[0, 1]),
}
CLASS_CASES = {
    "flip_affine": dict(augment=["flip", "affine"], image_keys=KEYS, mask_key="mask", t2_keys=[], flip_axis=[0, 1]),
    "flip_affine_shear_p06": dict(augment=["flip", "affine", "shear"], image_keys=KEYS, mask_key="mask", t2_keys=[], flip_axis=[0, 1, 2], prob=0.6),
    "trivial": dict(augment=["trivial", "flip", "affine", "shear"], image_keys=KEYS, mask_key=None, t2_keys=[], n_transforms_trivial=2),
    "intensity": dict(augment=["intensity", "flip", "affine"], image_keys=KEYS, mask_key="mask", t2_keys=[]),
}
SSL_CASES = {
    "shared_crop": dict(all_keys=["image"], copied_keys=["image_copy"], scaled_crop_size=None, roi_size=[128, 128, 32], vicregl=False, different_crop=False),
    "different_crop": dict(all_keys=["image"], copied_keys=["image_copy"], scaled_crop_size=None, roi_size=[128, 128, 32], vicregl=False, different_crop=True, n_transforms=2),
    "vicregl": dict(all_keys=["image"], copied_keys=["image_copy"], scaled_crop_size=None, roi_size=[96, 96, 24], vicregl=True, different_crop=False),
    "scaled_crop": dict(all_keys=["image"], copied_keys=["image_copy"], scaled_crop_size=[160, 160, 40], roi_size=[128, 128, 32], vicregl=False, different_crop=False),
}


SEG_COMMON = dict(all_keys=ALL, image_keys=KEYS, label_keys=["mask"], non_adc_keys=["t2", "dwi"], adc_keys=["adc"],
                  target_spacing=None, intp=["area", "area", "area", "nearest"],
                  intp_resampling_augmentations=["bilinear", "bilinear", "bilinear", "nearest"], possible_labels=[0, 1],
                  positive_labels=[1], all_aux_keys=[], resize_keys=[], feature_keys=[], aux_key_net=None, feature_key_net=None,
                  resize_size=None, label_mode="binary", fill_missing=False, brunet=False)
FACTORY_CASES = {
    "seg_plain": ("SegmentationTransforms", dict(SEG_COMMON, crop_size=None, pad_size=None, random_crop_size=None)),
    "seg_pad_crop_randomcrop": ("SegmentationTransforms", dict(SEG_COMMON, crop_size=[256, 256, 32], pad_size=[256, 256, 32], random_crop_size=[128, 128, 24])),
    "seg_unlabelled_randomcrop": ("SegmentationTransforms", dict(SEG_COMMON, all_keys=KEYS, label_keys=None, intp=["area"] * 3,
                                                                  intp_resampling_augmentations=["bilinear"] * 3, crop_size=None,
                                                                  pad_size=[64, 64, 16], random_crop_size=[32, 32, 8])),
    "class_crop_mask": ("ClassificationTransforms", dict(keys=KEYS, adc_keys=["adc"], clinical_feature_keys=[], target_spacing=None,
                                                          crop_size=[192, 192, 48], pad_size=[192, 192, 48], image_masking=False,
                                                          image_crop_from_mask=False, mask_key="mask", branched=False,
                                                          possible_labels=[0, 1], positive_labels=[1], label_groups=None, label_key=None,
                                                          target_size=None, label_mode="binary", cat_confounder_keys=None,
                                                          cont_confounder_keys=None)),
    "class_crop_from_mask": ("ClassificationTransforms", dict(keys=KEYS, adc_keys=["adc"], clinical_feature_keys=[], target_spacing=None,
                                                               crop_size=[64, 64, 16], pad_size=[64, 64, 16], image_masking=False,
                                                               image_crop_from_mask=True, mask_key="mask", branched=False,
                                                               possible_labels=[0, 1], positive_labels=[1], label_groups=None, label_key=None,
                                                               target_size=None, label_mode="binary", cat_confounder_keys=None,
                                                               cont_confounder_keys=None)),
    "class_plain_branched": ("ClassificationTransforms", dict(keys=["t2"], adc_keys=[], clinical_feature_keys=[], target_spacing=None,
                                                               crop_size=None, pad_size=None, image_masking=False,
                                                               image_crop_from_mask=False, mask_key=None, branched=True,
                                                               possible_labels=[0, 1], positive_labels=[1], label_groups=None, label_key=None,
                                                               target_size=None, label_mode="binary", cat_confounder_keys=None,
                                                               cont_confounder_keys=None)),
    "ssl_crop_pad": ("SSLTransforms", dict(all_keys=["image"], copied_keys=["image_copy"], adc_keys=[], non_adc_keys=["image"],
                                            target_spacing=None, crop_size=[160, 160, 40], pad_size=[160, 160, 40], resize_size=None,
                                            in_channels=1, n_dim=3, skip_augmentations=False, jpeg_dataset=False)),
    "seg_resize": ("SegmentationTransforms", dict(SEG_COMMON, resize_keys=["t2", "adc", "dwi", "mask"], resize_size=[128, 128, 24],
                                                   crop_size=None, pad_size=[128, 128, 32], random_crop_size=None)),
    "ssl_resize": ("SSLTransforms", dict(all_keys=["image"], copied_keys=["image_copy"], adc_keys=[], non_adc_keys=["image"],
                                          target_spacing=None, crop_size=[160, 160, 40], pad_size=[160, 160, 40], resize_size=[96, 96, 24],
                                          in_channels=1, n_dim=3, skip_augmentations=False, jpeg_dataset=False)),
    "ssl_adc": ("SSLTransforms", dict(all_keys=["t2", "adc"], copied_keys=["t2_copy", "adc_copy"], adc_keys=["adc"], non_adc_keys=["t2"],
                                       target_spacing=None, crop_size=None, pad_size=None, resize_size=None, in_channels=2, n_dim=3,
                                       skip_augmentations=False, jpeg_dataset=False)),
}


def main():
    M, A, TF = load_reference()
    W = M.AugmentationWorkhorsed
    out = {"unet": {}, "class": {}, "ssl": {}}
    for name, kw in UNET_CASES.items():
        out["unet"][name] = to_json(A.get_augmentations_unet(**kw), W)
    for name, kw in CLASS_CASES.items():
        out["class"][name] = to_json(A.get_augmentations_class(**kw), W)
    for name, kw in SSL_CASES.items():
        out["ssl"][name] = to_json(A.get_augmentations_ssl(**kw), W)
    out["factories"] = {}
    for name, (cls, kw) in FACTORY_CASES.items():
        obj = getattr(TF, cls)(**kw)
        out["factories"][name] = {"pre": to_json(obj.pre_transforms(), W), "post": to_json(obj.post_transforms(), W)}
    out["ADC_FACTOR"] = TF.ADC_FACTOR
    out["member_lists"] = {"generic": list(M.generic_augments), "mri_specific": list(M.mri_specific_augments), "spatial": list(M.spatial_augments)}
    json.dump(out, open(os.path.join(HERE, "builder_wiring.json"), "w"), indent=1, sort_keys=True)
    print("wrote builder_wiring.json:", {k: len(v) for k, v in out.items() if hasattr(v, "__len__")})


if __name__ == "__main__":
    main()
