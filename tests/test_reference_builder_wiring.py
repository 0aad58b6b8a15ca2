"""Builder wiring pinned against the reference itself:
tests/golden/builder_wiring.json was produced by RUNNING the reference's get_augmentations_unet /
_class / _ssl and AugmentationWorkhorsed with a recorder in place of `monai.transforms`
(tests/golden/make_golden_wiring.py), i.e. it lists which MONAI transform the reference constructs,
with which arguments, in which order.  The product's builders must construct the same chain with the
same effective parameters (for the members on the fused path; MONAI defaults spelled out)."""

import json
import os

import numpy as np
import pytest

from adell_mri_b200 import transform_factory as F, transforms as T

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "builder_wiring.json")))
import importlib.util

_spec = importlib.util.spec_from_file_location("make_golden_wiring", os.path.join(HERE, "golden", "make_golden_wiring.py"))
G = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(G)


def _rng(x):
    """Range arguments as nested lists of floats (tuples == lists, ints == floats)."""
    if x is None:
        return None
    if isinstance(x, (list, tuple)):
        return [_rng(v) for v in x]
    return float(x)


def _pair(x):
    """MONAI's scalar-or-pair convention for offsets / factors: ``f -> (-f, f)`` sorted."""
    if isinstance(x, (list, tuple)):
        return [float(min(x)), float(max(x))]
    return [float(min(-x, x)), float(max(-x, x))]


# ------------------------------------------------------------------ reference records -> effective config
def ref_cfg(r):
    cls, a, k = r["cls"], r.get("args", []), r.get("kwargs", {})
    if cls in ("Compose", "OneOf", "SomeOf"):
        out = {"cls": cls, "children": [ref_cfg(c) for c in a[0]]}
        if cls == "SomeOf":
            out["num_transforms"] = k["num_transforms"]
        return out
    if cls == "AugmentationWorkhorsed":
        fused = [m for m in r["augmentations"] if m in F.FUSED_AUGMENTS]
        return {"cls": cls, "augmentations": fused, "keys": r["keys"], "max_mult": r["max_mult"], "N": r["N"],
                "transforms": {m: ref_cfg(r["transforms"][m]) for m in fused}}
    keys = list(a[0]) if isinstance(a[0], list) else [a[0]]
    if cls == "Identityd":
        return {"cls": cls, "keys": keys}
    if cls == "RandAffined":
        mode = k.get("mode", "bilinear")
        return {"cls": cls, "keys": keys, "prob": k.get("prob", 0.1), "mode": mode if isinstance(mode, list) else [mode] * len(keys),
                "padding_mode": k.get("padding_mode", "reflection"),   # MONAI default
                "rotate_range": _rng(k.get("rotate_range")), "shear_range": _rng(k.get("shear_range")),
                "translate_range": _rng(k.get("translate_range")), "scale_range": _rng(k.get("scale_range"))}
    if cls == "RandFlipd":
        return {"cls": cls, "keys": keys, "prob": k.get("prob", 0.1), "spatial_axis": list(k["spatial_axis"])}
    if cls == "RandSpatialCropd":
        roi = a[1] if len(a) > 1 else k["roi_size"]
        return {"cls": cls, "keys": keys, "roi_size": [int(x) for x in roi], "random_size": k.get("random_size", False)}
    if cls == "RandCropByPosNegLabeld":
        return {"cls": cls, "keys": keys, "label_key": a[1], "spatial_size": [int(x) for x in a[2]], "num_samples": k["num_samples"],
                "allow_smaller": k["allow_smaller"], "fg_indices_key": k["fg_indices_key"], "bg_indices_key": k["bg_indices_key"],
                "pos_ratio": 0.5}   # MONAI defaults pos=neg=1
    if cls == "CenterSpatialCropd":
        return {"cls": cls, "keys": keys, "roi_size": [int(x) for x in a[1]]}
    if cls == "SpatialPadd":
        return {"cls": cls, "keys": keys, "spatial_size": [int(x) for x in a[1]]}
    if cls == "Resized":   # MONAI default mode: "area"
        mode = k.get("mode", "area")
        return {"cls": cls, "keys": keys, "spatial_size": [int(x) for x in (a[1] if len(a) > 1 else k["spatial_size"])],
                "mode": [mode] * len(keys) if isinstance(mode, str) else list(mode)}
    if cls == "ExposeTransformKeyMetad":
        return {"cls": cls, "key": a[0], "transform_class": a[1], "nested_pattern": list(a[2]), "output_key": a[3]}
    if cls == "Lambdad":
        return {"cls": cls, "keys": keys, "flatten_box": a[1]["outputs"]}
    if cls == "RandGaussianNoised":
        return {"cls": cls, "keys": keys, "prob": k["prob"], "std": float(k["std"]), "mean": 0.0, "sample_std": True}
    if cls == "RandShiftIntensityd":
        return {"cls": cls, "keys": keys, "prob": k["prob"], "offsets": _pair(k["offsets"])}
    if cls == "RandScaleIntensityd":
        return {"cls": cls, "keys": keys, "prob": k["prob"], "factors": _pair(k["factors"])}
    if cls == "RandStdShiftIntensityd":
        return {"cls": cls, "keys": keys, "prob": k["prob"], "factors": _pair(k["factors"])}
    if cls == "RandAdjustContrastd":
        g = k["gamma"]   # MONAI: a number means U(0.5, gamma)
        return {"cls": cls, "keys": keys, "prob": k["prob"], "gamma": [0.5, float(g)] if isinstance(g, (int, float)) else [float(x) for x in g]}
    if cls == "RandRicianNoised":   # MONAI defaults: mean 0, sample_std True
        return {"cls": cls, "keys": keys, "prob": k["prob"], "std": float(k["std"]), "mean": 0.0, "sample_std": True}
    raise AssertionError(f"unexpected reference transform {cls}")


# ------------------------------------------------------------------ product objects -> effective config
def our_cfg(t, roi_size=None):
    if isinstance(t, F.AugmentationWorkhorsed):
        return {"cls": "AugmentationWorkhorsed", "augmentations": list(t.augmentations), "keys": list(t.keys), "max_mult": t.max_mult,
                "N": t.N, "transforms": {m: our_cfg(t.transforms[m]) for m in t.augmentations}}
    if isinstance(t, T.SomeOf):
        assert t.min_num_transforms == t.max_num_transforms
        return {"cls": "SomeOf", "children": [our_cfg(c, roi_size) for c in t.transforms], "num_transforms": t.max_num_transforms}
    if isinstance(t, T.OneOf):
        return {"cls": "OneOf", "children": [our_cfg(c, roi_size) for c in t.transforms]}
    if isinstance(t, T.Compose):
        return {"cls": "Compose", "children": [our_cfg(c, roi_size) for c in t.transforms]}
    cls = type(t).__name__
    if cls == "ExposeTransformKeyMetad":
        return {"cls": cls, "key": t.key, "transform_class": t.transform_class, "nested_pattern": t.nested_pattern, "output_key": t.output_key}
    keys = list(t.keys)
    if cls == "Identityd":
        return {"cls": cls, "keys": keys}
    if cls == "RandAffined":
        pm = set(t.padding_mode)
        assert len(pm) == 1
        s = t.sampler
        return {"cls": cls, "keys": keys, "prob": t.prob, "mode": list(t.mode), "padding_mode": pm.pop(),
                "rotate_range": _rng(s.rotate_range), "shear_range": _rng(s.shear_range),
                "translate_range": _rng(s.translate_range), "scale_range": _rng(s.scale_range)}
    if cls == "RandFlipd":
        return {"cls": cls, "keys": keys, "prob": t.prob, "spatial_axis": list(t.spatial_axis)}
    if cls == "RandSpatialCropd":
        return {"cls": cls, "keys": keys, "roi_size": [int(x) for x in t.roi_size], "random_size": t.random_size}
    if cls == "RandCropByPosNegLabeld":
        return {"cls": cls, "keys": keys, "label_key": t.label_key, "spatial_size": t.spatial_size, "num_samples": t.num_samples,
                "allow_smaller": t.allow_smaller, "fg_indices_key": t.fg_indices_key, "bg_indices_key": t.bg_indices_key,
                "pos_ratio": t.pos_ratio}
    if cls == "CenterSpatialCropd":
        return {"cls": cls, "keys": keys, "roi_size": t.roi_size}
    if cls == "SpatialPadd":
        return {"cls": cls, "keys": keys, "spatial_size": [int(x) for x in t.spatial_size]}
    if cls == "Resized":
        return {"cls": cls, "keys": keys, "spatial_size": [int(x) for x in t.spatial_size], "mode": list(t.mode)}
    if cls == "Lambdad":
        return {"cls": cls, "keys": keys, "flatten_box": [[float(v) for v in t.func(b)] for b in G.FLATTEN_BOX_INPUTS]}
    if cls in ("RandGaussianNoised", "RandRicianNoised"):
        return {"cls": cls, "keys": keys, "prob": t.prob, "std": float(t.std), "mean": float(t.mean), "sample_std": t.sample_std}
    if cls == "RandShiftIntensityd":
        return {"cls": cls, "keys": keys, "prob": t.prob, "offsets": [float(x) for x in t.offsets]}
    if cls in ("RandScaleIntensityd", "RandStdShiftIntensityd"):
        return {"cls": cls, "keys": keys, "prob": t.prob, "factors": [float(x) for x in t.factors]}
    if cls == "RandAdjustContrastd":
        return {"cls": cls, "keys": keys, "prob": t.prob, "gamma": [float(x) for x in t.gamma]}
    raise AssertionError(f"unexpected product transform {cls}")


def _assert_same(ours, ref, path="root"):
    if isinstance(ref, dict):
        assert isinstance(ours, dict) and sorted(ours) == sorted(ref), (path, sorted(ours) if isinstance(ours, dict) else ours, sorted(ref))
        for k in ref:
            _assert_same(ours[k], ref[k], f"{path}.{k}")
    elif isinstance(ref, list):
        assert isinstance(ours, (list, tuple)) and len(ours) == len(ref), (path, ours, ref)
        for i, (o, r) in enumerate(zip(ours, ref)):
            _assert_same(o, r, f"{path}[{i}]")
    elif isinstance(ref, float):
        assert ours == pytest.approx(ref, rel=1e-12, abs=0), (path, ours, ref)
    else:
        assert ours == ref, (path, ours, ref)


@pytest.mark.parametrize("name", sorted(G.UNET_CASES))
def test_unet_builder_constructs_what_the_reference_constructs(name):
    ours = our_cfg(F.get_augmentations_unet(**G.UNET_CASES[name]))
    _assert_same(ours, ref_cfg(GOLD["unet"][name]))


@pytest.mark.parametrize("name", sorted(G.CLASS_CASES))
def test_classification_builder_constructs_what_the_reference_constructs(name):
    ours = our_cfg(F.get_augmentations_class(**G.CLASS_CASES[name]))
    _assert_same(ours, ref_cfg(GOLD["class"][name]))


@pytest.mark.parametrize("name", sorted(G.SSL_CASES))
def test_ssl_builder_constructs_what_the_reference_constructs(name):
    ours = [our_cfg(t) for t in F.get_augmentations_ssl(**G.SSL_CASES[name])]
    ref = [ref_cfg(r) for r in GOLD["ssl"][name]]
    _assert_same(ours, ref)


def test_member_vocabulary_matches_reference():
    assert F.generic_augments == GOLD["member_lists"]["generic"]
    assert F.mri_specific_augments == GOLD["member_lists"]["mri_specific"]
    assert F.spatial_augments == GOLD["member_lists"]["spatial"]


# ------------------------------------------------------------------ *Transforms factories (rows a1-a6)
#: reference entries that belong to the cached loading stage (disk IO, orientation, label
#: construction): out of the hot path, the product's factories do not emit them
LOADING_STAGE = {"LoadImaged", "Orientationd", "Spacingd", "ResampleToMatchd", "SampleChannelDimd", "CreateImageAndWeightsd"}


def ref_stage_cfg(r):
    cls, a, k = r["cls"], r["args"], r["kwargs"]
    keys = list(a[0]) if isinstance(a[0], list) else [a[0]]
    if cls == "ScaleIntensityd":
        pos = list(a[1:]) + [None] * 3
        return {"cls": cls, "keys": keys, "minv": k.get("minv", pos[0] if len(a) > 1 else 0.0), "maxv": k.get("maxv", pos[1] if len(a) > 2 else 1.0),
                "factor": k.get("factor", pos[2] if len(a) > 3 else None)}
    if cls == "ConditionalRescalingd":
        return {"cls": cls, "keys": keys, "max_value": a[1], "scale": a[2]}
    if cls == "Offsetd":
        return {"cls": cls, "keys": keys, "offset": a[1] if len(a) > 1 else k.get("offset")}
    if cls == "SpatialPadd":
        return {"cls": cls, "keys": keys, "spatial_size": [int(x) for x in a[1]]}
    if cls == "CenterSpatialCropd":
        return {"cls": cls, "keys": keys, "roi_size": [int(x) for x in a[1]]}
    if cls == "EnsureTyped":
        return {"cls": cls, "keys": keys}
    if cls == "Resized":
        return ref_cfg(r)
    if cls == "AdjustSizesd":
        return {"cls": cls, "keys": keys, "mode": k["mode"]}
    if cls == "FgBgToIndicesd":
        return {"cls": cls, "keys": keys}
    if cls == "ConcatItemsd":
        return {"cls": cls, "keys": keys, "name": a[1]}
    if cls == "SelectItemsd":
        return {"cls": cls, "keys": keys}
    if cls == "ToTensord":
        return {"cls": cls, "keys": keys}
    if cls == "CopyEntryd":
        return {"cls": cls, "keys": keys, "out_keys": a[1]}
    if cls == "CropFromMaskd":
        return {"cls": cls, "keys": keys, "mask_key": k["mask_key"], "output_size": [int(x) for x in k["output_size"]]}
    if cls == "CombineBinaryLabelsd":
        return {"cls": cls, "keys": keys, "mode": a[1], "output_key": a[2]}
    if cls == "LabelOperatorSegmentationd":
        return {"cls": cls, "keys": keys, "possible_labels": list(a[1]), "mode": k["mode"], "positive_labels": list(k["positive_labels"])}
    raise AssertionError(f"unexpected reference stage transform {cls}")


def our_stage_cfg(t):
    cls = type(t).__name__
    keys = list(t.keys)
    if cls == "ScaleIntensityd":
        return {"cls": cls, "keys": keys, "minv": t.minv, "maxv": t.maxv, "factor": t.factor}
    if cls == "ConditionalRescalingd":
        return {"cls": cls, "keys": keys, "max_value": t.max_value, "scale": t.scale}
    if cls == "Offsetd":
        return {"cls": cls, "keys": keys, "offset": t.offset}
    if cls == "SpatialPadd":
        return {"cls": cls, "keys": keys, "spatial_size": [int(x) for x in t.spatial_size]}
    if cls == "CenterSpatialCropd":
        return {"cls": cls, "keys": keys, "roi_size": [int(x) for x in t.roi_size]}
    if cls in ("EnsureTyped", "FgBgToIndicesd", "SelectItemsd", "ToTensord"):
        return {"cls": cls, "keys": keys}
    if cls == "AdjustSizesd":
        return {"cls": cls, "keys": keys, "mode": t.mode}
    if cls == "Resized":
        return our_cfg(t)
    if cls == "ConcatItemsd":
        return {"cls": cls, "keys": keys, "name": t.name}
    if cls == "CopyEntryd":
        return {"cls": cls, "keys": keys, "out_keys": dict(t.out_keys)}
    if cls == "CropFromMaskd":
        return {"cls": cls, "keys": keys, "mask_key": t.mask_key, "output_size": [int(x) for x in t.output_size]}
    if cls == "CombineBinaryLabelsd":
        return {"cls": cls, "keys": keys, "mode": t.mode, "output_key": t.output_key}
    if cls == "LabelOperatorSegmentationd":
        return {"cls": cls, "keys": keys, "possible_labels": list(t.possible_labels), "mode": t.mode, "positive_labels": list(t.positive_labels)}
    raise AssertionError(f"unexpected product stage transform {cls}")


#: arguments of the reference dataclasses that only steer the loading stage / label handling
_DROP = {"SegmentationTransforms": (),
         "ClassificationTransforms": ("possible_labels", "positive_labels", "label_groups", "label_key", "label_mode",
                                      "cat_confounder_keys", "cont_confounder_keys"),
         "SSLTransforms": ()}


@pytest.mark.parametrize("name", sorted(G.FACTORY_CASES))
def test_factories_emit_the_reference_chain_on_the_hot_path(name):
    cls, kw = G.FACTORY_CASES[name]
    kw = {k: v for k, v in kw.items() if k not in _DROP[cls]}
    if cls == "SegmentationTransforms" and kw.get("label_keys") is not None:
        kw["all_keys"] = [k for k in kw["all_keys"]]   # the combined label arrives under "mask" (cached stage)
    obj = getattr(F, cls)(**kw)
    for stage in ("pre", "post"):
        ref = [ref_stage_cfg(r) for r in GOLD["factories"][name][stage] if r["cls"] not in LOADING_STAGE]
        ours = [our_stage_cfg(t) for t in getattr(obj, stage + "_transforms")()]
        _assert_same(ours, ref, f"{name}.{stage}")


def test_adc_factor_matches_reference():
    assert F.ADC_FACTOR == GOLD["ADC_FACTOR"]
