"""Semi-supervised two-view segmentation pipeline: mirrors of
/root/reference/adell_mri/modules/semi_supervised_segmentation/utils.py:9-174 (the four argument
rewrites) and /root/reference/adell_mri/transform_factory/semi_sl_segmentation.py:17-58
(``get_semi_sl_transforms``).

The reference builds, per unlabelled sample: ``SegmentationTransforms(pre).pre_transforms()``, two
``CopyEntryd`` (``k -> k_aug_1 / k_aug_2``), ONE joint geometric augmentation over all copies
(``["affine", "shear", "flip"]``: both views stay aligned), one individual augmentation per view
(``["intensity", "noise", "rbf", "blur", "trivial"]``), the two ``post_transforms`` chains writing
``semi_sl_image_1 / 2`` and ``SelectItemsd``.  The joint part, ``intensity`` and ``trivial`` run on
the fused path; ``noise`` (Gibbs: FFT), ``rbf`` (polynomial bias field) and ``blur`` (stencil) are
outside it, so the reference's individual list raises ``NotImplementedError`` like
any other request for those tokens and ``individual_augment`` selects the subset to run.
"""

from __future__ import annotations

from itertools import product
from typing import Any, Sequence

from . import transforms as T
from .transform_factory import CopyEntryd, SegmentationTransforms, get_augmentations_unet

#: semi_supervised_segmentation/utils.py:93-97 and :140-146
JOINT_AUGMENT = ["affine", "shear", "flip"]
INDIVIDUAL_AUGMENT = ["intensity", "noise", "rbf", "blur", "trivial"]
#: the members of INDIVIDUAL_AUGMENT that the fused path implements
FUSED_INDIVIDUAL_AUGMENT = ["intensity", "trivial"]


def convert_arguments_pre(transform_arguments: dict[str, Any], image_keys: list[str]) -> dict[str, Any]:
    """utils.py:9-41: no labels, the image keys only, interpolation lists cut to their number."""
    out = dict(transform_arguments)
    out["label_keys"] = None
    out["all_keys"] = image_keys
    out["image_keys"] = image_keys
    out["intp"] = out["intp"][: len(image_keys)]
    out["intp_resampling_augmentations"] = out["intp_resampling_augmentations"][: len(image_keys)]
    return out


def convert_arguments_post(transform_arguments: dict[str, Any], idx: int, image_keys: list[str]) -> dict[str, Any]:
    """utils.py:44-72: view ``idx`` reads ``k_aug_idx`` and writes ``semi_sl_image_idx``."""
    out = dict(transform_arguments)
    out["label_keys"] = None
    out["all_keys"] = image_keys
    out["track_meta"] = True
    out["output_image_key"] = f"semi_sl_image_{idx}"
    out["image_keys"] = [f"{k}_aug_{idx}" for k in image_keys]
    return out


def convert_arguments_augment_all(augment_arguments: dict[str, Any], image_keys: list[str]) -> dict[str, Any]:
    """utils.py:75-110: the joint geometric augmentation over every copy of both views (key order:
    ``product(image_keys, [1, 2])``), whatever ``augment`` the user gave."""
    out = dict(augment_arguments)
    out["augment"] = list(JOINT_AUGMENT)
    out["all_keys"] = [f"{k}_aug_{idx}" for k, idx in product(image_keys, [1, 2])]
    out["image_keys"] = [f"{k}_aug_{idx}" for k, idx in product(image_keys, [1, 2])]
    out["has_label"] = False
    return out


def convert_arguments_augment_individual(augment_arguments: dict[str, Any], idx: int, image_keys: list[str]) -> dict[str, Any]:
    """utils.py:113-174: the per-view intensity-type augmentation of view ``idx``."""
    out = dict(augment_arguments)
    out["augment"] = list(INDIVIDUAL_AUGMENT)
    out["all_keys"] = [f"{k}_aug_{idx}" for k in image_keys]
    out["image_keys"] = [f"{k}_aug_{idx}" for k in image_keys]
    out["has_label"] = False
    return out


def get_semi_sl_transforms(transform_arguments: dict[str, Any], augment_arguments: dict[str, Any], keys: list[str],
                           individual_augment: Sequence[str] | None = None) -> T.Compose:
    """semi_sl_segmentation.py:17-58.  ``individual_augment=None`` keeps the reference's per-view
    token list (which raises: it names transforms outside the fused path);
    ``FUSED_INDIVIDUAL_AUGMENT`` runs its fused members."""
    pre = convert_arguments_pre(transform_arguments, keys)
    post = [convert_arguments_post(transform_arguments, idx, keys) for idx in (1, 2)]
    aug_all = convert_arguments_augment_all(augment_arguments, keys)
    aug_ind = [convert_arguments_augment_individual(augment_arguments, image_keys=keys, idx=idx) for idx in (1, 2)]
    if individual_augment is not None:
        bad = [a for a in individual_augment if a not in INDIVIDUAL_AUGMENT]
        if bad:
            raise ValueError(f"individual_augment may only hold members of {INDIVIDUAL_AUGMENT}, got {bad}")
        for a in aug_ind:
            a["augment"] = [x for x in INDIVIDUAL_AUGMENT if x in individual_augment]
    return T.Compose([
        *SegmentationTransforms(**pre).pre_transforms(),
        CopyEntryd(keys, {k: f"{k}_aug_1" for k in keys}),
        CopyEntryd(keys, {k: f"{k}_aug_2" for k in keys}),
        get_augmentations_unet(**aug_all),
        get_augmentations_unet(**aug_ind[0]),
        get_augmentations_unet(**aug_ind[1]),
        *SegmentationTransforms(**post[0]).post_transforms(),
        *SegmentationTransforms(**post[1]).post_transforms(),
        T.SelectItemsd(["semi_sl_image_1", "semi_sl_image_2"]),
    ])
