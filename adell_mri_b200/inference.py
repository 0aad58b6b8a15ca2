"""Test-time augmentation around the model call: flipped inference.

Mirror of ``adell_mri.utils.inference.FlippedInference``
(/root/reference/adell_mri/utils/inference.py:303-392): same constructor, the same
``flip`` over tensors / arrays / one-level dicts, lists and tuples (``flip_keys`` filter), the same
``output += flip(inference(flip(X)))`` accumulation and final division.  The reference flips with
``torch.flip(deepcopy(X), axis)``; here a float32 CUDA batch ``[B, C, H, W, D]`` (or ``[C, H, W, D]``) whose flip
axes are spatial is flipped by ONE K1 launch of box-copy items (TMA load -> TMA store for axes 0 / 1, a register
quad reversal for the contiguous axis): a bit-exact integer index transform, like every other flip on the path.
Anything else (numpy arrays, other dtypes, CPU tensors, non-spatial axes) takes ``np.flip`` / ``torch.flip``
exactly as the reference does — those are not voxel volumes on the device, so nothing is "fallen back" from.
"""

from __future__ import annotations

from copy import deepcopy
from typing import Callable, List, Sequence

import numpy as np
import torch

from . import _lib, engine
from .plan import BatchPlan


def flip_volumes(x: torch.Tensor, axis: Sequence[int]) -> torch.Tensor:
    """``torch.flip(x, axis)`` for a contiguous float32 CUDA tensor ``[..., H, W, D]`` with every flip axis among the
    last three: all leading-index volumes in one K1 launch, written into a new tensor."""
    nd = x.dim()
    axis = tuple(a % nd for a in axis)
    if x.device.type != "cuda" or x.dtype != torch.float32 or nd < 3 or any(a < nd - 3 for a in axis):
        raise ValueError("flip_volumes: float32 CUDA tensor with spatial flip axes expected")
    x = x.contiguous()
    out = torch.empty_like(x)
    if x.numel() == 0:
        return out
    shape = x.shape[-3:]
    vox = int(np.prod(shape))
    n = x.numel() // vox
    ptr = (x.data_ptr() + 4 * vox * np.arange(n, dtype=np.int64)).astype(np.uint64)
    dptr = (out.data_ptr() + 4 * vox * np.arange(n, dtype=np.int64)).astype(np.uint64)
    stride = np.tile(np.asarray([shape[1] * shape[2], shape[2], 1], np.int64), (n, 1))
    plan = BatchPlan.from_arrays(ptr, stride, np.full(n, _lib.F32, np.uint8), np.tile(np.asarray(shape, np.int64), (n, 1)),
                                 x.device, [x])
    plan.flip(np.array([(nd - 3 + a) in axis for a in range(3)]))
    engine.execute_ptrs(plan, dptr, stride, keep=[out])
    return out


class FlippedInference:
    """inference.py:303-392: runs ``inference_function`` on the input and on every flip of it, un-flips each result
    and averages."""

    def __init__(self, inference_function: Callable, flips: List[List[int]], flip_keys: List[str] = None, ndim: int = 3,
                 inference_batch_size: int = 1):
        self.inference_function = inference_function
        self.flips = flips
        self.flip_keys = flip_keys
        self.ndim = ndim

    def flip_array(self, X, axis):
        if isinstance(X, np.ndarray):
            return np.flip(deepcopy(X), axis)
        elif isinstance(X, torch.Tensor):
            nd = X.dim()
            ax = tuple(a % nd for a in axis) if nd else ()
            if X.device.type == "cuda" and X.dtype == torch.float32 and nd >= 3 and ax and all(a >= nd - 3 for a in ax):
                return flip_volumes(X, ax)
            return torch.flip(deepcopy(X), axis)

    def flip(self, X, axis: List[int]):
        axis = tuple(axis)
        if isinstance(X, (np.ndarray, torch.Tensor)):
            return self.flip_array(X, axis)
        elif isinstance(X, dict):
            X_out = dict(X)   # (the reference deep-copies the structure and then replaces the flipped entries)
            for k in X:
                if self.flip_keys is not None:
                    if k in self.flip_keys:
                        X_out[k] = self.flip_array(X[k], axis)
                    else:
                        X_out[k] = deepcopy(X[k])
                else:
                    X_out[k] = self.flip_array(X[k], axis)
            return X_out
        elif isinstance(X, (tuple, list)):
            X_out = list(X)
            for k in range(len(X)):
                if self.flip_keys is not None:
                    if k in self.flip_keys:
                        X_out[k] = self.flip_array(X[k], axis)
                    else:
                        X_out[k] = deepcopy(X[k])
                else:
                    X_out[k] = self.flip_array(X[k], axis)
            return X_out if isinstance(X, list) else tuple(X_out)
        else:
            raise NotImplementedError("Supported inputs are np.ndarray, dict, tuple, list")

    def __call__(self, X, *args, **kwargs):
        flips = self.flips
        output = self.inference_function(X, *args, **kwargs)
        for flip in flips:
            output += self.flip(self.inference_function(self.flip(X, flip), *args, **kwargs), flip)
        output = output / (len(flips) + 1)
        return output
