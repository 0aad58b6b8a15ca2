"""Chain cases shared by the CPU (C restatement) and GPU (CUDA) parity tests.

Each case builds (a) the reference result by executing the reference op sequence one op at a
time with the literal torch oracle and (b) a recorded BatchPlan for the same ops.  ``exact``
says whether the comparison must be bit-exact (integer work, nearest masks, strict trilinear)
or within the stated tolerance (rtol=1e-4 of the reference, atol=1e-4 of its dynamic range)."""

from __future__ import annotations

import numpy as np
import torch

from adell_mri_b200.plan import BatchPlan
from oracle import monai_restated as M
from tests.helpers import rand_affine_matrix


def _vol(R, shape, kind="f32"):
    if kind == "mask":
        return torch.from_numpy((R.rand(1, *shape) > 0.6).astype(np.float32))
    if kind == "i16":
        return torch.from_numpy(R.randint(-200, 4096, size=(1, *shape)).astype(np.int16))
    if kind == "u8":
        return torch.from_numpy(R.randint(0, 256, size=(1, *shape)).astype(np.uint8))
    return torch.from_numpy(R.rand(1, *shape).astype(np.float32))


def _apply(plan: BatchPlan, ops):
    for op in ops:
        if op[0] == "flip":
            plan.flip(np.array([a in op[1] for a in range(3)]))
        elif op[0] == "pad":
            plan.spatial_pad(op[1])
        elif op[0] == "center_crop":
            plan.center_crop(op[1])
        elif op[0] == "crop":
            plan.crop(op[1], op[2])
    return plan


def chain_case(seed, shape, pre_ops, affine, mode, padding, post_ops, kind="f32", device="cpu", strict=True,
               post_scale=None, post_offset=None, noise=False):
    R = np.random.RandomState(seed)
    img = _vol(R, shape, kind)
    A = rand_affine_matrix(R) if affine else None
    nz = None
    ref = M.canonical_item(img, pre_ops=pre_ops, affine=A, mode=mode, padding_mode=padding, post_ops=post_ops)
    if post_scale is not None:
        ref = ref * torch.tensor(post_scale, dtype=torch.float32)
    if post_offset is not None:
        ref = ref + torch.tensor(post_offset, dtype=torch.float32)
    if noise:
        nz = torch.from_numpy(R.normal(0, 0.1, size=tuple(ref.shape[1:])).astype(np.float32))
        ref = ref + nz
    plan = BatchPlan([img[0].to(device)], strict=strict)
    _apply(plan, pre_ops)
    if A is not None:
        plan.affine(A.numpy(), mode, padding)
    _apply(plan, post_ops)
    if post_scale is not None:
        plan.intensity(scale=post_scale)
    if post_offset is not None:
        plan.intensity(offset=post_offset)
    if noise:
        plan.add_noise([nz.to(device)])
    return plan, ref[0].contiguous()


# (name, kwargs) — every case is bit-exact unless strict=False
CHAIN_CASES = [
    ("unet_affine_flip_crop", dict(shape=(40, 36, 16), pre_ops=[], affine=True, mode="bilinear", padding="reflection",
                                   post_ops=[("flip", [0]), ("flip", [2]), ("center_crop", (32, 28, 12))])),
    ("unet_mask_nearest", dict(shape=(40, 36, 16), pre_ops=[], affine=True, mode="nearest", padding="reflection",
                               post_ops=[("flip", [1]), ("center_crop", (32, 28, 12))], kind="mask")),
    ("class_flip_then_affine_zeros", dict(shape=(48, 48, 24), pre_ops=[("flip", [0, 1])], affine=True, mode="bilinear",
                                          padding="zeros", post_ops=[("center_crop", (32, 32, 8))])),
    ("class_mask_flip_then_affine", dict(shape=(48, 48, 24), pre_ops=[("flip", [2])], affine=True, mode="nearest",
                                         padding="zeros", post_ops=[("center_crop", (32, 32, 8))], kind="mask")),
    ("pad_crop_then_affine", dict(shape=(20, 30, 10), pre_ops=[("pad", (32, 32, 16)), ("center_crop", (28, 28, 14))],
                                  affine=True, mode="bilinear", padding="zeros", post_ops=[])),
    ("pad_then_affine_reflection", dict(shape=(20, 30, 10), pre_ops=[("pad", (32, 32, 16))], affine=True,
                                        mode="bilinear", padding="reflection", post_ops=[])),
    ("crop_pad_then_affine_border", dict(shape=(30, 30, 20), pre_ops=[("crop", (3, 5, 2), (20, 18, 12)), ("pad", (24, 24, 16)), ("flip", [1])],
                                         affine=True, mode="nearest", padding="border", post_ops=[("flip", [0, 2])])),
    ("affine_then_pad", dict(shape=(20, 20, 8), pre_ops=[], affine=True, mode="bilinear", padding="border",
                             post_ops=[("center_crop", (12, 12, 6)), ("pad", (16, 16, 8)), ("flip", [1])])),
    ("identity_flip_crop_pad", dict(shape=(22, 18, 9), pre_ops=[("flip", [0, 2]), ("crop", (2, 1, 0), (16, 16, 8)), ("pad", (20, 20, 12)), ("flip", [1])],
                                    affine=False, mode="bilinear", padding="zeros", post_ops=[])),
    ("identity_i16", dict(shape=(16, 16, 16), pre_ops=[("flip", [1]), ("center_crop", (12, 12, 8))], affine=False,
                          mode="bilinear", padding="zeros", post_ops=[], kind="i16")),
    ("affine_i16_nearest", dict(shape=(24, 24, 12), pre_ops=[], affine=True, mode="nearest", padding="zeros", post_ops=[], kind="i16")),
    ("affine_u8_trilinear", dict(shape=(24, 24, 12), pre_ops=[], affine=True, mode="bilinear", padding="border", post_ops=[], kind="u8")),
    ("odd_shape_trilinear", dict(shape=(19, 23, 7), pre_ops=[], affine=True, mode="bilinear", padding="reflection",
                                 post_ops=[("center_crop", (17, 17, 5))])),
    ("ssl_scale_shift_noise", dict(shape=(32, 32, 16), pre_ops=[("crop", (4, 2, 1), (24, 24, 12))], affine=True,
                                   mode="bilinear", padding="zeros", post_ops=[], post_scale=1.13, post_offset=-0.2, noise=True)),
    ("noise_only", dict(shape=(16, 16, 8), pre_ops=[], affine=False, mode="bilinear", padding="zeros", post_ops=[], noise=True)),
    ("wide_d_48", dict(shape=(12, 12, 48), pre_ops=[], affine=True, mode="bilinear", padding="zeros", post_ops=[])),
    ("wide_d_40", dict(shape=(12, 12, 40), pre_ops=[], affine=True, mode="nearest", padding="reflection", post_ops=[])),
]


def two_resample_case(seed, device="cpu"):
    """["affine","shear"]: two sequential RandAffined firings (parity mode = two passes)."""
    R = np.random.RandomState(seed)
    img = _vol(R, (28, 24, 12))
    A1 = rand_affine_matrix(R)
    A2 = M.compose_affine(shear=[R.uniform(0.9, 1.1) for _ in range(3)])
    ref = M.affine_resample(img, A1, "bilinear", "reflection")
    ref = M.affine_resample(ref, A2, "bilinear", "reflection")
    ref = M.flip(ref, [0])
    plan = BatchPlan([img[0].to(device)], strict=True)
    plan.affine(A1.numpy(), "bilinear", "reflection").affine(A2.numpy(), "bilinear", "reflection")
    plan.flip(np.array([True, False, False]))
    return plan, ref[0].contiguous()


def noise_then_resample_case(seed, device="cpu"):
    """SSL workhorse order [gaussian_noise, rotate, shift_intensity]: noise is resampled."""
    R = np.random.RandomState(seed)
    img = _vol(R, (24, 24, 12))
    nz = torch.from_numpy(R.normal(0, 0.2, size=(24, 24, 12)).astype(np.float32))
    A = rand_affine_matrix(R)
    ref = img + nz
    ref = M.affine_resample(ref, A, "bilinear", "zeros")
    ref = ref + torch.tensor(0.1, dtype=torch.float32)
    plan = BatchPlan([img[0].to(device)], strict=True)
    plan.add_noise([nz.to(device)]).affine(A.numpy(), "bilinear", "zeros").intensity(offset=0.1)
    return plan, ref[0].contiguous()
