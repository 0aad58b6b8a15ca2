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


# ----------------------------------------------------------------------------- sliding window
def _example(X):
    if isinstance(X, (np.ndarray, torch.Tensor)):
        return X
    if isinstance(X, dict):
        return X[list(X.keys())[0]]
    if isinstance(X, (tuple, list)):
        return X[0]
    raise NotImplementedError("Supported inputs are np.ndarray, dict, tuple, list")


def gather_windows(x: torch.Tensor, coords: Sequence, batched: bool) -> torch.Tensor:
    """The windows ``x[..., x1:x2, y1:y2, z1:z2]`` of a float32 CUDA tensor for every entry of ``coords``, already
    concatenated (``batched``: ``x`` is ``[B, C, H, W, D]``, windows are concatenated along axis 0) or stacked
    (``x`` is ``[C, H, W, D]``) the way ``multi_format_stack_or_cat`` would: ONE K1 launch of box-copy items
    writes all of them into the inference batch (the reference slices, then copies with ``torch.cat``)."""
    lead = x.shape[:-3]
    shape = x.shape[-3:]
    n_lead = int(np.prod(lead)) if len(lead) else 1
    size = tuple(c[1] - c[0] for c in coords[0])
    out = torch.empty((len(coords) * lead[0], *lead[1:], *size) if batched else (len(coords), *lead, *size), dtype=torch.float32, device=x.device)
    x = x.contiguous()
    vox, ovox = int(np.prod(shape)), int(np.prod(size))
    n = len(coords) * n_lead
    src = x.data_ptr() + 4 * vox * np.tile(np.arange(n_lead, dtype=np.int64), len(coords))
    dst = out.data_ptr() + 4 * ovox * np.arange(n, dtype=np.int64)
    stride = np.tile(np.asarray([shape[1] * shape[2], shape[2], 1], np.int64), (n, 1))
    ostride = np.tile(np.asarray([size[1] * size[2], size[2], 1], np.int64), (n, 1))
    plan = BatchPlan.from_arrays(src.astype(np.uint64), stride, np.full(n, _lib.F32, np.uint8), np.tile(np.asarray(shape, np.int64), (n, 1)),
                                 x.device, [x])
    starts = np.repeat(np.array([[c[0] for c in cc] for cc in coords], np.int64), n_lead, axis=0)
    plan.crop(starts, np.asarray(size, np.int64))
    engine.execute_ptrs(plan, dst.astype(np.uint64), ostride, keep=[out])
    return out


class SlidingWindowSegmentation:
    """Mirror of ``adell_mri.utils.inference.SlidingWindowSegmentation``
    (/root/reference/adell_mri/utils/inference.py:394-786): the same window grid (stride, windows that would leave
    the image moved back inside it), the same batching of ``inference_batch_size`` windows per model call, the same
    sum / count accumulation and final division.  For a float32 CUDA tensor input (or a dict / list / tuple of
    them) the windows of one inference batch are gathered by one K1 launch (:func:`gather_windows`)."""

    def __init__(self, sliding_window_size, inference_function: Callable, n_classes: int, stride=None, inference_batch_size: int = 1):
        self.sliding_window_size = sliding_window_size
        self.inference_function = inference_function
        self.n_classes = n_classes
        self.stride = stride if stride is not None else sliding_window_size
        self.inference_batch_size = inference_batch_size
        self.ndim = len(sliding_window_size)

    def adjust_if_necessary(self, x1: int, x2: int, M: int, a: int):
        if x2 > M:
            x1, x2 = M - a, M
        return x1, x2

    def all_coords(self, sh):
        """Window bounds in the reference's iteration order (get_all_crops_2d / _3d)."""
        ranges = [range(0, sh[a], self.stride[a]) for a in range(self.ndim)]
        import itertools

        for idx in itertools.product(*ranges):
            yield tuple(self.adjust_if_necessary(i, i + self.sliding_window_size[a], sh[a], self.sliding_window_size[a])
                        for a, i in enumerate(idx))

    def extract_patch_from_array(self, X, coords):
        sl = tuple(slice(a, b) for a, b in coords)
        return X[(..., *sl)]

    def _windows(self, x, coords_list, batched):
        """One inference batch of windows of one array / tensor."""
        if isinstance(x, torch.Tensor) and x.device.type == "cuda" and x.dtype == torch.float32 and self.ndim == 3:
            return gather_windows(x, coords_list, batched)
        patches = [self.extract_patch_from_array(x, c) for c in coords_list]
        if isinstance(x, np.ndarray):
            return np.concatenate(patches, 0) if batched else np.stack(patches, 0)
        return torch.cat(patches, 0) if batched else torch.stack(patches, 0)

    def extract_batch(self, X, coords_list, batched):
        if isinstance(X, (np.ndarray, torch.Tensor)):
            return self._windows(X, coords_list, batched)
        if isinstance(X, dict):
            return {k: self._windows(X[k], coords_list, batched) for k in X if isinstance(X[k], (np.ndarray, torch.Tensor))}
        if isinstance(X, (tuple, list)):
            return [self._windows(x, coords_list, batched) for x in X if isinstance(x, (np.ndarray, torch.Tensor))]
        raise NotImplementedError("Supported inputs are np.ndarray, torch.Tensor, dict, tuple, list")

    def update_output(self, output_array, output_denominator, tmp_out, coords):
        sl = (..., *(slice(a, b) for a, b in coords))
        output_array[sl] += tmp_out.squeeze(0).squeeze(0)
        output_denominator[sl] += 1.0
        return output_array, output_denominator

    def __call__(self, X, *args, **kwargs):
        ex = _example(X)
        output_size = list(ex.shape)
        batched = len(output_size) == self.ndim + 2
        if batched:
            output_size[1] = self.n_classes
        elif len(output_size) < self.ndim + 2:
            output_size[0] = self.n_classes
        else:
            raise Exception("length of input array shape should be <= self.ndim+2")
        if isinstance(ex, np.ndarray):
            output_array, output_denominator = np.zeros(output_size), np.zeros(output_size)
            split = lambda t, n: np.split(t, t.shape[0] // n, 0)
        else:
            output_array = torch.zeros(output_size, device=ex.device)
            output_denominator = torch.zeros(output_size, device=ex.device)
            split = lambda t, n: torch.split(t, n, 0)
        original_batch_size = output_size[0]
        pending = []

        def flush():
            batch = self.extract_batch(X, pending, batched)
            with torch.no_grad():
                batch_out = self.inference_function(batch, *args, **kwargs)
            for out, coords in zip(split(batch_out, original_batch_size), pending):
                self.update_output(output_array, output_denominator, out, coords)
            pending.clear()

        for coords in self.all_coords(ex.shape[-self.ndim:]):
            pending.append(coords)
            if len(pending) == self.inference_batch_size:
                flush()
        if pending:
            flush()
        return output_array / output_denominator
