#!/usr/bin/env python
"""Generates tests/golden/k1_golden.npz with the literal torch oracle (oracle/monai_restated.py:
MONAI operation order + the ATen CPU grid_sample kernel).  MONAI itself cannot be imported in
this image, so these vectors pin the oracle's behaviour at generation time
(torch 2.11.0 CPU, MKL) — any later drift of the oracle, the C restatement or the CUDA path
shows up against them.  Run from the repo root:  python tests/golden/make_golden.py"""

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import monai_restated as M  # noqa: E402

CASES = []
R = np.random.RandomState(20261018)


def add(name, shape, mode, padding, rotate, translate, scale, shear=None, pre_flip=(), post_flip=(), crop=None, mask=False):
    img = R.rand(1, *shape).astype(np.float32)
    if mask:
        img = (img > 0.55).astype(np.float32)
    p = M.rand_affine_grid_params(R, rotate_range=rotate, shear_range=shear, translate_range=translate, scale_range=scale)
    A = M.compose_affine(p["rotate"], p["shear"], p["translate"], p["scale"])
    pre = [("flip", list(pre_flip))] if pre_flip else []
    post = ([("flip", list(post_flip))] if post_flip else []) + ([("center_crop", crop)] if crop else [])
    out = M.canonical_item(torch.from_numpy(img), pre_ops=pre, affine=A, mode=mode, padding_mode=padding, post_ops=post)
    CASES.append(dict(name=name, img=img[0], A=A.numpy(), mode=mode, padding=padding, pre_flip=list(pre_flip),
                      post_flip=list(post_flip), crop=list(crop) if crop else [], out=out[0].numpy()))


add("unet_trilinear_reflection", (20, 18, 8), "bilinear", "reflection", [np.pi / 8, np.pi / 8, np.pi / 16], None, None,
    post_flip=(0, 2), crop=(16, 14, 6))
add("unet_mask_nearest_reflection", (20, 18, 8), "nearest", "reflection", [np.pi / 8, np.pi / 8, np.pi / 16], None, None,
    post_flip=(1,), crop=(16, 14, 6), mask=True)
add("class_trilinear_zeros", (24, 24, 12), "bilinear", "zeros", [np.pi / 16], [4, 4, 1], [0.1, 0.1, 0.05],
    pre_flip=(0, 1), crop=(16, 16, 8))
add("class_mask_nearest_zeros", (24, 24, 12), "nearest", "zeros", [np.pi / 16], [4, 4, 1], [0.1, 0.1, 0.05],
    pre_flip=(2,), crop=(16, 16, 8), mask=True)
add("benchmark_affine_border", (16, 16, 8), "bilinear", "border", [np.pi / 6] * 3, [10, 10, 3], [0.1] * 3)
add("shear_nearest_border", (14, 12, 10), "nearest", "border", None, None, None, shear=((0.9, 1.1),) * 3)
add("ssl_scale_quirk_zeros", (16, 16, 8), "bilinear", "zeros", None, None, ((0.85, 1.15), 0, 0))

out = {}
for i, c in enumerate(CASES):
    for k, v in c.items():
        out[f"{i}:{k}"] = np.asarray(v)
out["n"] = np.asarray(len(CASES))
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "k1_golden.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes,", len(CASES), "cases")
