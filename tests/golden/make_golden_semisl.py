"""Generates tests/golden/semi_sl_wiring.json by RUNNING THE REFERENCE's own semi-supervised transform
builder (/root/reference/adell_mri/transform_factory/semi_sl_segmentation.py: get_semi_sl_transforms, with
/root/reference/adell_mri/modules/semi_supervised_segmentation/utils.py) in the build container, under the
recorder stand-in for `monai.transforms` of make_golden_wiring.py.  Also stores the outputs of the four
argument rewrites on the inputs of /root/reference/testing/test_semi_sl_utils.py.

    python tests/golden/make_golden_semisl.py        # needs /root/reference; not run on the GPU box
"""
import importlib.util
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_wiring as W  # noqa: E402

KEYS, ALL = ["t2", "adc"], ["t2", "adc", "mask"]
TRANSFORM_ARGUMENTS = dict(all_keys=ALL, image_keys=KEYS, label_keys=["mask"], non_adc_keys=["t2"], adc_keys=["adc"],
                           target_spacing=None, intp=["area", "area", "nearest"],
                           intp_resampling_augmentations=["bilinear", "bilinear", "nearest"], possible_labels=[0, 1],
                           positive_labels=[1], all_aux_keys=[], resize_keys=[], feature_keys=[], aux_key_net=None,
                           feature_key_net=None, resize_size=None, crop_size=[64, 64, 16], pad_size=[64, 64, 16],
                           random_crop_size=None, label_mode="binary", fill_missing=False, brunet=False)
AUGMENT_ARGUMENTS = dict(augment=["affine"], all_keys=ALL, image_keys=KEYS, t2_keys=[], random_crop_size=None, n_crops=1,
                         flip_axis=[0, 1, 2])
# the literal inputs of the reference's own unit test (testing/test_semi_sl_utils.py:12-47)
UNIT_TRANSFORM_ARGUMENTS = {"label_keys": ["mask"], "all_keys": ["image", "mask"], "image_keys": ["image"], "non_adc_keys": ["image"],
                            "adc_keys": [], "target_spacing": 1.0, "intp": ["area", "nearest"],
                            "intp_resampling_augmentations": ["bilinear", "nearest"], "possible_labels": 2, "positive_labels": None,
                            "all_aux_keys": [], "resize_keys": [], "feature_keys": [], "aux_key_net": None, "feature_key_net": None,
                            "resize_size": None, "pad_size": None, "crop_size": None, "random_crop_size": None, "label_mode": "binary",
                            "fill_missing": False, "brunet": False}
UNIT_AUGMENT_ARGUMENTS = {"augment": ["affine"], "all_keys": ["image", "mask"], "image_keys": ["image"], "t2_keys": [],
                          "random_crop_size": None, "n_crops": 1, "flip_axis": [0, 1, 2]}


def main():
    M, A, TF = W.load_reference()
    mods = {}
    for name, path in (("adell_mri.modules.semi_supervised_segmentation", None),
                       ("adell_mri.modules.semi_supervised_segmentation.utils", "modules/semi_supervised_segmentation/utils.py"),
                       ("adell_mri.transform_factory.semi_sl_segmentation", "transform_factory/semi_sl_segmentation.py")):
        if path is None:
            import types
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
            continue
        spec = importlib.util.spec_from_file_location(name, os.path.join(W.ROOT, path))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    U = mods["adell_mri.modules.semi_supervised_segmentation.utils"]
    S = mods["adell_mri.transform_factory.semi_sl_segmentation"]
    compose = S.get_semi_sl_transforms(dict(TRANSFORM_ARGUMENTS), dict(AUGMENT_ARGUMENTS), list(KEYS))
    out = {"pipeline": W.to_json(compose, M.AugmentationWorkhorsed),
           "inputs": {"transform_arguments": TRANSFORM_ARGUMENTS, "augment_arguments": AUGMENT_ARGUMENTS, "keys": KEYS},
           "unit_inputs": {"transform_arguments": UNIT_TRANSFORM_ARGUMENTS, "augment_arguments": UNIT_AUGMENT_ARGUMENTS},
           "rewrites": {
               "pre": U.convert_arguments_pre(UNIT_TRANSFORM_ARGUMENTS, ["image"]),
               "post_2": U.convert_arguments_post(UNIT_TRANSFORM_ARGUMENTS, 2, ["image"]),
               "augment_all": U.convert_arguments_augment_all(UNIT_AUGMENT_ARGUMENTS, ["image"]),
               "augment_individual_2": U.convert_arguments_augment_individual(UNIT_AUGMENT_ARGUMENTS, 2, ["image"]),
               "pre_two_keys": U.convert_arguments_pre(TRANSFORM_ARGUMENTS, KEYS),
               "augment_all_two_keys": U.convert_arguments_augment_all(AUGMENT_ARGUMENTS, KEYS),
           }}
    json.dump(out, open(os.path.join(HERE, "semi_sl_wiring.json"), "w"), indent=1, sort_keys=True)
    kids = out["pipeline"]["args"][0]
    print("wrote semi_sl_wiring.json:", [k["cls"] for k in kids])


if __name__ == "__main__":
    main()
