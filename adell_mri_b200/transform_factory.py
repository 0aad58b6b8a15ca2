"""Mirror of ``adell_mri/transform_factory`` for the fused GPU hot path.

Same names, keys and argument meaning as the reference's pipeline builders
(/root/reference/adell_mri/transform_factory/transforms.py:53-67,70-263,367-557,706-820;
/root/reference/adell_mri/transform_factory/augmentations.py:19-178,181-320,391-516;
/root/reference/adell_mri/modules/augmentations.py:10-256), built from the lazy dictionary
transforms of :mod:`adell_mri_b200.transforms`, so an entrypoint can swap

    from adell_mri.transform_factory import SegmentationTransforms, get_augmentations_unet
    from adell_mri.utils.utils import safe_collate

for the same names from this package.  What differs by design:

* samples enter as channel-first device tensors (the device-resident analogue of what
  ``LoadImaged`` + ``Orientationd`` hand on); file loading, re-orientation and ``Spacingd``
  are outside the hot path and raise if requested (``Resized`` runs on the device: K5);
* members of the augmentation vocabulary that are not gathers / pointwise maps raise
  ``NotImplementedError`` instead of silently doing something else.
"""

from __future__ import annotations

import itertools
from dataclasses import dataclass
from typing import Sequence

import numpy as np
import torch

from . import _lib, stats
from . import transforms as T
from .transforms import Pending

ADC_FACTOR = -2 / 3  # transforms.py:25
IMAGE_INTERPOLATION = "bilinear"


# --------------------------------------------------------------------------- custom transforms
class _DeviceIntensity(T.MapTransform):
    """Cached-stage intensity op evaluated by the exact intensity-program kernel
    (``y = ((x*m0 - a)/d)*m1*m2 + b``, every step rounded to fp32, identity steps skipped)."""

    def coefs(self, mm: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def __call__(self, data):
        d = dict(data)
        for k in self.key_iterator(d):
            x = d[k].tensor() if isinstance(d[k], Pending) else d[k]
            flat = [x.contiguous().reshape(-1)]
            mm = stats.minmax(flat)
            d[k] = stats.intensity_map(flat, self.coefs(mm))[0].reshape(x.shape)
        return d


class ConditionalRescalingd(_DeviceIntensity):
    """/root/reference/adell_mri/utils/monai_transforms/image_intensity_ops.py:49-103:
    ``if X.max() > max_value: X = X * scale``."""

    def __init__(self, keys, max_value: float, scale: float):
        super().__init__(keys, allow_missing_keys=True)
        self.max_value, self.scale = max_value, scale

    def coefs(self, mm):
        c = torch.tensor([[1.0, 0.0, 1.0, 1.0, 1.0, 0.0]], dtype=torch.float32, device=mm.device).repeat(mm.shape[0], 1)
        c[:, 0] = torch.where(mm[:, 1] > self.max_value, float(np.float32(self.scale)), 1.0)
        return c


class Offsetd(_DeviceIntensity):
    """image_intensity_ops.py:106-143: ``X - offset`` (the array minimum when ``offset`` is None)."""

    def __init__(self, keys, offset: float | None = None):
        super().__init__(keys, allow_missing_keys=False)
        self.offset = offset

    def coefs(self, mm):
        c = torch.tensor([[1.0, 0.0, 1.0, 1.0, 1.0, 0.0]], dtype=torch.float32, device=mm.device).repeat(mm.shape[0], 1)
        c[:, 1] = mm[:, 0] if self.offset is None else float(np.float32(self.offset))
        return c


class CopyEntryd(T.Transform):
    """/root/reference/adell_mri/utils/monai_transforms/generic_data_ops.py:7-26 (``out_keys`` is
    an ``input_key -> output_key`` dict); the recorded chain is cloned, the voxels are shared."""

    def __init__(self, keys, out_keys):
        self.keys = list(keys)
        self.out_keys = dict(out_keys) if isinstance(out_keys, dict) else dict(zip(keys, out_keys))

    def __call__(self, data):
        d = dict(data)
        for k in list(d.keys()):
            if k in self.keys and k in self.out_keys:
                d[k] = T.as_pending(d[k])
                d[self.out_keys[k]] = d[k].clone()
        return d


class AdjustSizesd(T.MapTransform):
    """/root/reference/adell_mri/utils/monai_transforms/image_ops.py:368-438: ``mode="crop"`` crops every
    key to the per-axis minimum size over the keys, window start ``(size - target) // 2`` (NOT MONAI's
    ``CenterSpatialCrop`` start ``size // 2 - target // 2``: they differ when ``size`` is even and
    ``target`` odd — pinned by tests/golden/custom_transforms.npz); ``mode="pad"`` pads to the per-axis
    maximum with ``pad // 2`` zeros before and the rest after."""

    def __init__(self, keys, ndim: int = 3, mode: str = "pad"):
        super().__init__(keys)
        if ndim != 3:
            raise NotImplementedError("AdjustSizesd: the fused path handles 3-D volumes")
        self.ndim, self.mode = ndim, mode

    def __call__(self, data):
        d = dict(data)
        for k in self.key_iterator(d):
            d[k] = T.as_pending(d[k])
        sizes = np.array([d[k].spatial_shape for k in self.keys])
        if self.mode == "pad":
            target = np.max(sizes, axis=0)
            for k in self.keys:
                pad = [int(t - s) for t, s in zip(target, d[k].spatial_shape)]
                before = [p // 2 for p in pad]
                d[k].plan.pad(before, [p - b for p, b in zip(pad, before)])
        else:
            target = np.min(sizes, axis=0)
            for k in self.keys:
                start = [int(s - t) // 2 for s, t in zip(d[k].spatial_shape, target)]
                d[k].plan.crop(start, [int(t) for t in target])
        return d


# --------------------------------------------------------------------------- workhorse (SSL)
generic_augments = ["gaussian_noise", "shift_intensity", "scale_intensity", "contrast", "gaussian_smooth_x",
                    "gaussian_smooth_y", "gaussian_smooth_z", "gaussian_sharpen_x", "gaussian_sharpen_y",
                    "gaussian_sharpen_z", "coarse_dropout"]
mri_specific_augments = ["rbf", "gibbs_noise", "spike_noise", "rician_noise"]
spatial_augments = ["rotate_x", "rotate_y", "rotate_z", "translate_x", "translate_y", "translate_z",
                    "shear_x", "shear_y", "shear_z", "scale_x", "scale_y", "scale_z"]
#: members of the reference vocabulary that the fused path implements
FUSED_AUGMENTS = ["gaussian_noise", "shift_intensity", "scale_intensity", "contrast", "rician_noise", *spatial_augments]


class CombineBinaryLabelsd(T.Transform):
    """/root/reference/adell_mri/utils/monai_transforms/labels.py:189-220: voxel-wise ``any`` (sum > 0)
    or ``majority`` (mean > 0.5) of the binary label maps under ``keys``, written as a float32 map to
    ``output_key`` (default: the first key).  One pointwise device pass (``adell_label_map``); a
    cached-stage transform: pending inputs are materialised."""

    def __init__(self, keys, mode: str = "any", output_key: str | None = None):
        self.keys, self.mode = list(keys), mode
        self.output_key = self.keys[0] if output_key is None else output_key

    def __call__(self, X):
        from . import stats

        d = dict(X)
        vols = [(d[k].tensor() if isinstance(d[k], T.Pending) else d[k]).contiguous() for k in self.keys]
        if self.mode in ("any", "majority"):
            d[self.output_key] = stats.label_map(vols, self.mode, None, [])
        else:   # the reference leaves the stacked maps untouched for other modes
            d[self.output_key] = torch.stack(vols, -1)
        return d


class LabelOperatorSegmentationd(T.Transform):
    """labels.py:123-186: ``mode="binary"``: 1 where the value is one of ``positive_labels``;
    ``mode="cat"``: the index of the value in ``possible_labels`` (0 when absent); any other mode:
    unchanged.  One pointwise device pass (``adell_label_map``)."""

    def __init__(self, keys, possible_labels, mode: str = "cat", positive_labels=[1], output_keys={}):
        self.keys, self.possible_labels, self.mode = list(keys), list(possible_labels), mode
        self.positive_labels, self.output_keys = list(positive_labels), dict(output_keys)

    def __call__(self, data):
        from . import stats

        d = dict(data)
        for key in self.keys:
            out_key = self.output_keys.get(key, key)
            x = (d[key].tensor() if isinstance(d[key], T.Pending) else d[key]).contiguous()
            if self.mode == "cat":
                d[out_key] = stats.label_map([x], None, "cat", self.possible_labels)
            elif self.mode == "binary":
                d[out_key] = stats.label_map([x], None, "binary", self.positive_labels)
            else:
                d[out_key] = x
        return d


class CropFromMask:
    """``adell_mri.utils.monai_transforms.CropFromMask``
    (/root/reference/adell_mri/utils/monai_transforms/labels.py:412-477): the window is either the
    bounding box of the mask's non-zero voxels or, when ``output_size`` is given (or the mask is empty),
    an ``output_size`` box centred on the bounding box (on the volume centre for an empty mask) and
    shifted back inside the volume.  The bounding box is one reduction on the device
    (``adell_mask_bbox``); the crop itself is a recorded integer window."""

    def __init__(self, output_size=None, lazy: bool = False):
        self.output_size = None if output_size is None else [int(x) for x in output_size]

    def get_centre_extremes(self, mask):
        from . import stats

        m = mask.tensor() if isinstance(mask, T.Pending) else mask
        if m.shape[0] != 1:
            raise NotImplementedError("CropFromMask: the mask must have one channel (the reference's slicing fails otherwise)")
        box = stats.mask_bbox([m[0].contiguous()])[0].tolist()          # cached-stage transform: one small D2H
        if box[1] > 0:
            extremes = [(box[2 * a], box[2 * a + 1]) for a in range(3)]
            centre = [(e[1] + e[0]) // 2 for e in extremes]
        else:
            centre = [int(c) // 2 for c in m.shape[1:]]
            extremes = None
        return centre, extremes

    def compute_slices(self, shape, mask):
        centres, extremes = self.get_centre_extremes(mask)
        min_shape = [int(x) for x in shape[1:]]
        if (self.output_size is not None) or (extremes is None):
            if self.output_size is None:
                raise TypeError("CropFromMask: an empty mask needs output_size (as in the reference)")
            if any(o > s for o, s in zip(self.output_size, min_shape)):
                raise NotImplementedError("CropFromMask: output_size larger than the volume (pad first, as the reference pipelines do)")
            half_size = [x // 2 for x in self.output_size]
            extremes = [(c - h, c + (o - h)) for c, h, o in zip(centres, half_size, self.output_size)]
            for i in range(len(extremes)):
                if extremes[i][0] < 0:
                    extremes[i] = (0, self.output_size[i])
                if extremes[i][1] > min_shape[i]:
                    extremes[i] = (min_shape[i] - self.output_size[i], min_shape[i])
        return [slice(int(e[0]), int(e[1])) for e in extremes]


class CropFromMaskd(T.MapTransform):
    """Dictionary version (labels.py:480-522; used by ``ClassificationTransforms(image_crop_from_mask=True)``,
    transform_factory/transforms.py:474-481).  ``inverse`` zero-pads the entries back to their size
    before the crop (MONAI ``Cropd.inverse`` †)."""

    def __init__(self, keys, mask_key: str, output_size=None, lazy: bool = False):
        super().__init__([keys] if isinstance(keys, str) else keys)
        self.mask_key, self.output_size = mask_key, output_size
        self.cropper = CropFromMask(output_size=output_size)

    def __call__(self, data, lazy: bool = False):
        d = dict(data)
        mask = d[self.mask_key]
        for k in self.keys:
            p = T.as_pending(d[k]).clone() if isinstance(d[k], T.Pending) else T.as_pending(d[k])
            full = p.spatial_shape
            sl = self.cropper.compute_slices(p.shape, mask)
            start = [s.start for s in sl]
            size = [s.stop - s.start for s in sl]
            p.plan.crop(start, size)
            p.meta.setdefault("crop_from_mask", []).append((start, size, full))
            d[k] = p
        return d

    def inverse(self, data):
        d = dict(data)
        for k in self.keys:
            p = d[k]
            if not isinstance(p, T.Pending) or not p.meta.get("crop_from_mask"):
                raise RuntimeError(f"CropFromMaskd.inverse: no crop recorded for key {k}")
            p = p.clone()
            start, size, full = p.meta["crop_from_mask"].pop()
            p.plan.pad(start, [f - s - z for f, s, z in zip(full, start, size)])
            d[k] = p
        return d


class GetAllCrops:
    """``adell_mri.utils.monai_transforms.GetAllCrops``
    (/root/reference/adell_mri/utils/monai_transforms/image_ops.py:257-331): every ``size``-shaped crop
    of a volume on the regular grid, after a symmetric zero pad of the axes whose missing remainder is
    smaller than half a crop.  Literal reference quirks kept: the pad amounts are derived from the
    shape shifted by the channel axis (see ``get_pad_size``), the loops run over the UN-padded extent
    and a crop is emitted only if it ends inside it, so the pad never creates an extra crop.  Here each
    crop is a recorded integer pad + crop on a :class:`Pending` entry: nothing is copied until
    ``safe_collate_crops`` gathers all crops of the batch in one K1 launch of box copies."""

    def __init__(self, size):
        self.size = None if size is None else [int(x) for x in size]
        self.ndim = None if size is None else len(self.size)

    def get_pad_size(self, sh):
        # literal reference behaviour (image_ops.py:275-284): ``sh`` is the FULL shape, channel axis
        # included, so the remainders are taken from (C, H, W) against size[0..2] but applied to
        # (H, W, D) — pinned by tests/golden/custom_transforms.npz, produced by the reference itself
        remainder = [(y - (x % y)) if x > y else 0 for x, y in zip(sh, self.size)]
        remainder = [x if x < (y // 2) else 0 for x, y in zip(remainder, self.size)]
        return [(0, 0), *[(x // 2, x - x // 2) for x in remainder]]

    def crop_windows(self, sh):
        """``(pad_before, pad_after, [start, ...])`` for a ``[C, H, W, D]`` shape."""
        if self.ndim != 3:
            raise NotImplementedError("GetAllCrops: the fused path handles 3-D volumes")
        pads = self.get_pad_size(sh)[1:]
        ext = list(sh[1:])
        starts = [(i, j, k)
                  for i in range(0, ext[0], self.size[0]) for j in range(0, ext[1], self.size[1]) for k in range(0, ext[2], self.size[2])
                  if i + self.size[0] < ext[0] + 1 and j + self.size[1] < ext[1] + 1 and k + self.size[2] < ext[2] + 1]
        return [p[0] for p in pads], [p[1] for p in pads], starts

    def __call__(self, X):
        if self.size is None:
            return X
        base = T.as_pending(X)
        before, after, starts = self.crop_windows(base.shape)
        out = []
        for st in starts:
            p = base.clone()
            p.plan.pad(before, after)
            p.plan.crop(st, self.size)
            out.append(p)
        return out


class GetAllCropsd(T.MapTransform):
    """Dictionary version (image_ops.py:334-365): a list with one dict per crop, the other entries
    shared.  Used by the validation pipelines of the segmentation entrypoints
    (/root/reference/adell_mri/entrypoints/segmentation/train.py:359-371)."""

    def __init__(self, keys, size):
        super().__init__(keys)
        self.size = size
        self.tr = GetAllCrops(self.size)

    def __call__(self, X):
        crops = {k: list(self.tr(X[k])) for k in self.keys}
        outputs = []
        for elements in zip(*[crops[k] for k in self.keys]):
            output = {k: e for k, e in zip(self.keys, elements)}
            for k in X:
                if k not in output:
                    output[k] = X[k]
            outputs.append(output)
        return outputs


def _aug_param_dict():
    """modules/augmentations.py:103-129 (a fresh copy: the reference mutates its dict in place)."""
    d = {"gaussian_noise": {"std": 1}, "shift_intensity": {"offsets": 0.5}, "scale_intensity": {"factors": 0.5},
         "contrast": {"gamma": 3}, "rician_noise": {"std": 0.2}}
    for c in ["x", "y", "z"]:
        t = 30 if c != "z" else 5
        a = np.pi / 6 if c != "z" else np.pi / 16
        d["rotate_" + c] = {"rotate_range": a}
        d["translate_" + c] = {"translate_range": t}
        d["shear_" + c] = {"shear_range": 0.5}
        d["scale_" + c] = {"scale_range": 0.3}
    return d


def _axis_tuple(c: str, value):
    i = "xyz".index(c)
    out = [0, 0, 0]
    out[i] = value
    return tuple(out)


def _param_correction(name: str, x):
    """modules/augmentations.py:131-162."""
    kind, _, c = name.partition("_")
    if kind in ("rotate", "translate") and c in "xyz" and c:
        return _axis_tuple(c, (-x, x))
    if kind in ("shear", "scale") and c in "xyz" and c and name not in ("scale_intensity",):
        return _axis_tuple(c, (1 - x, 1 + x))
    return x


def get_transform_d(keys, transform_str: str, params: dict, mask_keys=()):
    """modules/augmentations.py:165-186 for the fused members (``prob=1.0`` each)."""
    if transform_str not in FUSED_AUGMENTS:
        raise NotImplementedError(f"workhorse member '{transform_str}' is outside the fused GPU hot path")
    params = {k: (_param_correction(transform_str, v) if transform_str in spatial_augments else v) for k, v in params.items()}
    if transform_str == "gaussian_noise":
        return T.RandGaussianNoised(keys, prob=1.0, **params)
    if transform_str == "shift_intensity":
        return T.RandShiftIntensityd(keys, prob=1.0, **params)
    if transform_str == "scale_intensity":
        return T.RandScaleIntensityd(keys, prob=1.0, **params)
    if transform_str == "contrast":       # AUG_PARAM_CORRECTION["contrast"]: x + 0.51, gamma ~ U(0.5, x + 0.51)
        return T.RandAdjustContrastd(keys, prob=1.0, **{k: v + 0.51 for k, v in params.items()})
    if transform_str == "rician_noise":
        return T.RandRicianNoised(keys, prob=1.0, **params)
    mode = ["bilinear" if k not in mask_keys else "nearest" for k in keys]
    return T.RandAffined(keys, prob=1.0, mode=mode, padding_mode="zeros", **params)


class AugmentationWorkhorsed(T.RandomizableTransform):
    """modules/augmentations.py:189-256: ``np.random.choice(augmentations, N, replace=False)`` on
    the GLOBAL numpy stream (as the reference does), then the chosen members in that order.  The
    members' own streams are seeded from this transform's ``set_random_state`` (the reference
    leaves them unseeded, i.e. not reproducible; seeding them is a superset of that behaviour)."""

    def __init__(self, augmentations, keys=None, mask_keys=(), max_mult: float = 1.0, N: int = 5,
                 aug_param_dict=None, dropout_size=(32, 32, 2)):
        super().__init__(1.0)
        self.augmentations = list(augmentations)
        self.keys, self.mask_keys, self.max_mult, self.N = keys, list(mask_keys), max_mult, N
        bad = [k for k in self.augmentations if k not in FUSED_AUGMENTS]
        if bad:
            raise NotImplementedError(f"workhorse members {bad} are outside the fused GPU hot path")
        base = aug_param_dict if aug_param_dict is not None else _aug_param_dict()
        self.param_dict = {k: {kk: base[k][kk] * max_mult for kk in base[k]} for k in self.augmentations}
        self.transforms = {k: get_transform_d(keys, k, self.param_dict[k], self.mask_keys) for k in self.param_dict}

    def set_random_state(self, seed=None, state=None):
        super().set_random_state(seed, state)
        for k in self.augmentations:
            self.transforms[k].set_random_state(seed=int(self.R.randint(T.MAX_SEED, dtype="uint32")))
        return self

    def __call__(self, X):
        t_list = np.random.choice(self.augmentations, self.N, replace=False)
        self.last_choice = [str(t) for t in t_list]
        for t in t_list:
            X = self.transforms[t](X)
        return X


# --------------------------------------------------------------------------- augmentation builders
_UNET_VALID = ["intensity", "noise", "rbf", "affine", "shear", "flip", "blur", "distort", "lowres", "trivial"]
_CLASS_VALID = ["intensity", "noise", "rbf", "affine", "shear", "flip", "blur", "lowres", "distort", "trivial"]
_FUSED_TOKENS = {"affine", "shear", "flip", "trivial", "intensity"}


def _check_tokens(augment, valid):
    for a in augment:
        if a not in valid:
            raise NotImplementedError("augment can only contain {}".format(valid))
    bad = [a for a in augment if a not in _FUSED_TOKENS]
    if bad:
        raise NotImplementedError(f"augmentations {bad} are outside the fused GPU hot path (DESIGN.md, out of scope)")


def get_augmentations_unet(augment, all_keys, image_keys, t2_keys, random_crop_size: list[int] = None,
                           has_label: bool = True, n_crops: int = 1, flip_axis: list[int] = [0, 1]):
    """augmentations.py:19-178."""
    _check_tokens(augment, _UNET_VALID)
    interpolation = ["bilinear" if k in image_keys else "nearest" for k in all_keys]
    augments = []
    prob = 0.2
    if "trivial" in augment:
        augments.append(T.Identityd(image_keys))
        prob = 1.0
    if "intensity" in augment:   # augmentations.py:66-76 (listed before the spatial members)
        augments.extend([T.RandAdjustContrastd(image_keys, gamma=(0.5, 1.5), prob=prob),
                         T.RandStdShiftIntensityd(image_keys, factors=0.1, prob=prob)])
    if "affine" in augment:
        augments.append(T.RandAffined(all_keys, rotate_range=[np.pi / 8, np.pi / 8, np.pi / 16], prob=prob, mode=interpolation))
    if "shear" in augment:
        augments.append(T.RandAffined(all_keys, shear_range=((0.9, 1.1), (0.9, 1.1), (0.9, 1.1)), prob=prob, mode=interpolation))
    flip_transform = [T.RandFlipd(all_keys, spatial_axis=[axis], prob=0.25) for axis in flip_axis] if "flip" in augment else []
    if "trivial" in augment:
        augments = T.Compose([T.OneOf(augments), *flip_transform])
    else:
        augments = T.Compose([*augments, *flip_transform])
    if random_crop_size is not None:
        pre_final_size = [int(i * 1.10) for i in random_crop_size]
        new_augments = []
        if has_label is True:
            new_augments.append(T.RandCropByPosNegLabeld(
                [*image_keys, "mask"], "mask", pre_final_size, allow_smaller=True, num_samples=n_crops,
                fg_indices_key="mask_fg_indices", bg_indices_key="mask_bg_indices"))
        else:
            new_augments.append(T.RandSpatialCropd(image_keys, pre_final_size))
        augments = T.Compose([*new_augments, augments,
                              T.CenterSpatialCropd([*image_keys, "mask"] if has_label else image_keys, random_crop_size)])
    return augments


def get_augmentations_class(augment, image_keys, mask_key, t2_keys, flip_axis: list[int] = [0, 1], prob: float = 0.1,
                            n_transforms_trivial: int = 1):
    """augmentations.py:181-320."""
    _check_tokens(augment, _CLASS_VALID)
    all_keys_with_mask = [k for k in image_keys]
    if mask_key is not None:
        all_keys_with_mask.append(mask_key)
    intp = ["bilinear" if k != mask_key else "nearest" for k in all_keys_with_mask]
    augments = []
    if "trivial" in augment:
        augments.append(T.Identityd(image_keys))
        prob = 1.0
    if "intensity" in augment:   # augmentations.py:219-232
        augments.extend([T.RandAdjustContrastd(image_keys, gamma=(0.5, 1.5), prob=prob),
                         T.RandStdShiftIntensityd(image_keys, factors=0.1, prob=prob),
                         T.RandShiftIntensityd(image_keys, offsets=0.1, prob=prob)])
    if "flip" in augment:
        if isinstance(flip_axis, int):
            flip_axis = [flip_axis]
        flips = []
        for i in range(len(flip_axis)):
            for axis_to_flip in itertools.combinations(flip_axis, i + 1):
                flips.append(T.RandFlipd(all_keys_with_mask, prob=prob, spatial_axis=axis_to_flip))
        augments.append(T.OneOf(flips))
    if "affine" in augment:
        augments.append(T.RandAffined(all_keys_with_mask, translate_range=[4, 4, 1], rotate_range=[np.pi / 16],
                                      scale_range=[0.1, 0.1, 0.05], prob=prob, mode=intp, padding_mode="zeros"))
    if "shear" in augment:
        augments.append(T.RandAffined(all_keys_with_mask, shear_range=((0.9, 1.1), (0.9, 1.1), (0.9, 1.1)), prob=prob,
                                      mode=intp, padding_mode="zeros"))
    if "trivial" in augment:
        return T.SomeOf(augments, num_transforms=n_transforms_trivial)
    return T.Compose(augments)


def flatten_box(box, roi_size):
    """augmentations.py:402-406."""
    box1 = np.array(box[::2])
    box2 = np.array(roi_size) - np.array(box[1::2])
    return np.concatenate([box1, box2]).astype(np.float32)


def get_augmentations_ssl(all_keys, copied_keys, scaled_crop_size, roi_size, vicregl: bool, different_crop: bool,
                          n_transforms=3, n_dim: int = 3, skip_augmentations: bool = False, aug_list=None):
    """augmentations.py:391-516.  ``aug_list`` defaults to the fused members of the reference's
    list (after its own removals); passing the full reference list raises for the first member
    that is not on the fused path."""
    roi_size = tuple(int(x) for x in roi_size)
    all_keys, copied_keys = list(all_keys), list(copied_keys)
    transforms_to_remove = []
    if vicregl is True:
        transforms_to_remove.extend(spatial_augments)
    if n_dim == 2:
        transforms_to_remove.extend(["rotate_z", "translate_z", "shear_z", "scale_z"])
    else:
        transforms_to_remove.extend(["gaussian_sharpen_x", "gaussian_sharpen_y", "gaussian_sharpen_z"])
    if aug_list is None:
        aug_list = [x for x in generic_augments + mri_specific_augments + spatial_augments if x in FUSED_AUGMENTS]
    aug_list = [x for x in aug_list if x not in transforms_to_remove]
    cropping_strategy = []
    if scaled_crop_size is not None:   # augmentations.py:427-444
        scaled_crop_size = tuple(int(x) for x in scaled_crop_size)
        small_crop_size = [x // 2 for x in scaled_crop_size]
        cropping_strategy.extend([
            T.SpatialPadd(all_keys + copied_keys, small_crop_size),
            T.RandSpatialCropd(all_keys + copied_keys, roi_size=small_crop_size, random_size=True),
            T.Resized(all_keys + copied_keys, scaled_crop_size),
        ])
    if skip_augmentations is True:
        return cropping_strategy
    if vicregl is True:
        cropping_strategy.extend([
            T.RandSpatialCropd(all_keys, roi_size=roi_size, random_size=False),
            T.RandSpatialCropd(copied_keys, roi_size=roi_size, random_size=False),
            T.ExposeTransformKeyMetad(all_keys[0], "RandSpatialCrop", ["extra_info", "cropped"], "box_1"),
            T.ExposeTransformKeyMetad(copied_keys[0], "RandSpatialCrop", ["extra_info", "cropped"], "box_2"),
            T.Lambdad(["box_1", "box_2"], lambda box: flatten_box(box, roi_size)),
        ])
    elif different_crop is True:
        cropping_strategy.extend([
            T.RandSpatialCropd(all_keys, roi_size=roi_size, random_size=False),
            T.RandSpatialCropd(copied_keys, roi_size=roi_size, random_size=False),
        ])
    else:
        cropping_strategy.append(T.RandSpatialCropd(all_keys + copied_keys, roi_size=roi_size, random_size=False))
    dropout_size = tuple(x // 10 for x in roi_size)
    out = [*cropping_strategy,
           AugmentationWorkhorsed(augmentations=aug_list, keys=all_keys, mask_keys=[], max_mult=0.5, N=n_transforms,
                                  dropout_size=dropout_size)]
    if len(copied_keys) > 0:
        out.append(AugmentationWorkhorsed(augmentations=aug_list, keys=copied_keys, mask_keys=[], max_mult=0.5,
                                          N=n_transforms, dropout_size=dropout_size))
    return out


# --------------------------------------------------------------------------- pipeline builders
@dataclass
class TransformMixin:
    """transforms.py:45-67."""

    def pre_transforms(self):
        raise NotImplementedError("pre_transform must be implemented")

    def post_transforms(self):
        raise NotImplementedError("post_transform must be implemented")

    def transforms(self, augmentations=None, final_transforms=None):
        transforms = [*self.pre_transforms()]
        if augmentations:
            if isinstance(augmentations, T.Transform):
                transforms.append(augmentations)
            else:
                transforms.extend(augmentations)
        transforms.extend(self.post_transforms())
        if final_transforms is not None:
            transforms.extend(final_transforms)
        return T.Compose(transforms)


def _reject(name, value):
    if value is not None and value is not False and value != []:
        raise NotImplementedError(f"{name} belongs to the cached loading stage (file IO / clinical tables): outside the fused hot path")


def _intensity_stage(non_adc_keys, adc_keys, offset_adc: bool):
    """transforms.py:143-155 (seg), 430-443 (class, with Offsetd), 772-786 (ssl)."""
    out = []
    if len(non_adc_keys) > 0:
        out.append(T.ScaleIntensityd(list(non_adc_keys), minv=0.0, maxv=1.0))
    if len(adc_keys) > 0:
        out.append(ConditionalRescalingd(list(adc_keys), 500, 0.001))
        if offset_adc:
            out.append(Offsetd(list(adc_keys), None))
        out.append(T.ScaleIntensityd(list(adc_keys), None, None, ADC_FACTOR))
    return out


@dataclass
class SegmentationTransforms(TransformMixin):
    """transforms.py:70-263 (in-scope fields; the others must stay at their defaults)."""

    all_keys: Sequence[str]
    image_keys: Sequence[str]
    label_keys: Sequence[str] | None
    non_adc_keys: Sequence[str]
    adc_keys: Sequence[str]
    target_spacing: Sequence[float] | None = None
    intp: Sequence[str] | None = None
    intp_resampling_augmentations: Sequence[str] | None = None
    output_image_key: str = "image"
    possible_labels: Sequence[int] = (0, 1)
    positive_labels: Sequence[int] = (1,)
    all_aux_keys: Sequence[str] = ()
    resize_keys: Sequence[str] = ()
    feature_keys: Sequence[str] = ()
    aux_key_net: str | None = None
    feature_key_net: str | None = None
    resize_size: Sequence[int] | None = None
    crop_size: Sequence[int] | None = None
    pad_size: Sequence[int] | None = None
    random_crop_size: Sequence[int] | None = None
    label_mode: str | None = None
    fill_missing: bool = False
    brunet: bool = False
    track_meta: bool = False
    convert_to_tensor: bool = True

    def __post_init__(self):
        _reject("fill_missing", self.fill_missing)
        _reject("brunet", self.brunet)
        _reject("all_aux_keys", list(self.all_aux_keys))
        _reject("feature_keys", list(self.feature_keys))
        self.all_keys, self.image_keys = list(self.all_keys), list(self.image_keys)
        self.transform_keys = [self.output_image_key]
        self.mask_key = ["mask"] if self.label_keys is not None else []

    def pre_transforms(self):
        transforms = []
        if self.target_spacing is not None:   # transforms.py:133-140 (after Orientationd, before the intensity scalers)
            transforms.append(T.Spacingd(keys=self.all_keys, pixdim=self.target_spacing, mode=self.intp_resampling_augmentations))
        transforms += _intensity_stage(self.non_adc_keys, self.adc_keys, offset_adc=False)
        if self.resize_size is not None and self.resize_keys:   # transforms.py:157-167
            intp_ = [k for k, kk in zip(self.intp, self.all_keys) if kk in self.resize_keys]
            transforms.append(T.Resized(list(self.resize_keys), tuple(self.resize_size), mode=intp_))
        if self.pad_size is not None:
            transforms.append(T.SpatialPadd(self.all_keys, self.pad_size))
        if self.crop_size is not None:
            transforms.append(T.CenterSpatialCropd(self.all_keys, self.crop_size))
        transforms.append(T.EnsureTyped(self.all_keys, dtype=torch.float32))
        if self.label_keys is not None:   # transforms.py:181-194: the label maps become the 0/1 "mask"
            transforms.extend([
                CombineBinaryLabelsd(list(self.label_keys), "any", "mask"),
                LabelOperatorSegmentationd(["mask"], list(self.possible_labels), mode=self.label_mode,
                                           positive_labels=list(self.positive_labels)),
            ])
        if self.random_crop_size is not None:
            if self.label_keys is not None:
                transforms.append(AdjustSizesd([*self.image_keys, "mask"], mode="crop"))
                transforms.append(T.FgBgToIndicesd("mask"))
            else:
                transforms.append(AdjustSizesd(self.image_keys, mode="crop"))
        return transforms

    def post_transforms(self):
        transforms = [T.ConcatItemsd(self.image_keys, self.output_image_key)]
        if self.convert_to_tensor is True:
            transforms.append(T.ToTensord([self.output_image_key] + self.mask_key, track_meta=self.track_meta, dtype=torch.float32))
        if not self.track_meta:
            transforms.append(T.SelectItemsd(self.transform_keys + self.mask_key))
        return transforms


@dataclass
class ClassificationTransforms(TransformMixin):
    """transforms.py:367-557 (image part; label / tabular / confounder entries pass through)."""

    keys: Sequence[str]
    adc_keys: Sequence[str]
    clinical_feature_keys: Sequence[str] = ()
    target_spacing: Sequence[float] | None = None
    crop_size: Sequence[int] | None = None
    pad_size: Sequence[int] | None = None
    image_masking: bool = False
    image_crop_from_mask: bool = False
    mask_key: str | None = None
    branched: bool = False
    target_size: Sequence[int] | None = None

    def __post_init__(self):
        _reject("target_size", self.target_size)
        _reject("image_masking", self.image_masking)
        self.keys = list(self.keys)
        self.non_adc_keys = [k for k in self.keys if k not in self.adc_keys]
        self.all_keys = [k for k in self.keys]
        if self.mask_key is not None:
            self.all_keys.append(self.mask_key)
        self.crop_size_with_margin = [int(j) + 16 for j in self.crop_size] if self.crop_size is not None else None
        self.crop_size_final = [int(j) for j in self.crop_size] if self.crop_size is not None else None

    def pre_transforms(self):
        transforms = _intensity_stage(self.non_adc_keys, self.adc_keys, offset_adc=True)
        if self.target_spacing is not None:   # transforms.py:444-454 (after the scalers here; IMAGE_INTERPOLATION / nearest for the mask)
            interpolation = ["bilinear" if k != self.mask_key else "nearest" for k in self.all_keys]
            transforms.append(T.Spacingd(self.all_keys, pixdim=self.target_spacing, dtype=torch.float32, mode=interpolation))
        if self.pad_size is not None:
            transforms.append(T.SpatialPadd(self.all_keys, self.crop_size_with_margin))
        if self.image_crop_from_mask is True:
            transforms.append(CropFromMaskd(self.all_keys, mask_key=self.mask_key, output_size=self.crop_size_with_margin))
        elif self.crop_size is not None:
            transforms.append(T.CenterSpatialCropd(self.all_keys, self.crop_size_with_margin))
        transforms.append(T.EnsureTyped(self.all_keys))
        return transforms

    def post_transforms(self):
        transforms = []
        if self.crop_size is not None:
            transforms.append(T.CenterSpatialCropd(self.all_keys, self.crop_size_final))
        if self.branched is not True:
            transforms.append(T.ConcatItemsd(self.all_keys, "image"))
        return transforms


@dataclass
class SSLTransforms(TransformMixin):
    """transforms.py:706-820."""

    all_keys: Sequence[str]
    copied_keys: Sequence[str]
    adc_keys: Sequence[str]
    non_adc_keys: Sequence[str]
    target_spacing: Sequence[float] | None = None
    crop_size: Sequence[int] | None = None
    pad_size: Sequence[int] | None = None
    resize_size: Sequence[int] | None = None
    in_channels: int = 1
    n_dim: int = 3
    skip_augmentations: bool = False
    jpeg_dataset: bool = False

    def __post_init__(self):
        _reject("jpeg_dataset", self.jpeg_dataset)
        if self.n_dim != 3:
            raise NotImplementedError("the fused hot path is volumetric (n_dim=3)")
        self.all_keys, self.copied_keys = list(self.all_keys), list(self.copied_keys)
        self.output_keys = ["image"] if self.skip_augmentations else ["augmented_image_1", "augmented_image_2"]
        self.concat_keys = [self.all_keys] if self.skip_augmentations else [self.all_keys, self.copied_keys]

    def pre_transforms(self):
        transforms = []
        if self.target_spacing is not None:   # transforms.py:761-771 (IMAGE_INTERPOLATION for every key)
            transforms.append(T.Spacingd(keys=self.all_keys, pixdim=self.target_spacing, mode=["bilinear" for _ in self.all_keys]))
        transforms += _intensity_stage(self.non_adc_keys, self.adc_keys, offset_adc=False)
        if self.crop_size is not None:
            transforms.append(T.CenterSpatialCropd(self.all_keys, [int(j) for j in self.crop_size]))
        if self.pad_size is not None:
            transforms.append(T.SpatialPadd(self.all_keys, [int(j) for j in self.pad_size]))
        if self.resize_size is not None:   # transforms.py:799-804 (MONAI's default mode: "area")
            transforms.append(T.Resized(self.all_keys, [int(j) for j in self.resize_size]))
        transforms.append(T.EnsureTyped(self.all_keys))
        if self.skip_augmentations is False:
            transforms.append(CopyEntryd(self.all_keys, {k: kk for k, kk in zip(self.all_keys, self.copied_keys)}))
        return transforms

    def post_transforms(self):
        transforms = [T.ConcatItemsd(keys, output_key) for keys, output_key in zip(self.concat_keys, self.output_keys)]
        transforms.append(T.ToTensord(self.output_keys, track_meta=False))
        return transforms
