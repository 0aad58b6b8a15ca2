"""Batch-level mixing right after collation: label smoothing and (partial) mixup.

Drop-in mirror of ``adell_mri.utils.batch_preprocessing``
(/root/reference/adell_mri/utils/batch_preprocessing.py:15-181, wired into the training step in
/root/reference/adell_mri/utils/network_factories.py:201-212): same function names, arguments and
random draws (``numpy.random.Generator``: ``beta`` then ``permutation``; partial mixup: ``binomial``,
``beta``, ``permutation``), so results are bit-identical to the reference's for the same seed.  The
batch itself is mixed by ONE pass of the ``adell_mixup`` kernel (12 B per element instead of the
three elementwise passes + gathered copy torch makes); the handful of labels is mixed with the same
fp32 torch expressions as the reference.  CUDA tensors only — there is no CPU fallback.
"""

from __future__ import annotations

import ctypes as C
from typing import Tuple

import numpy as np
import torch

from . import _lib


def label_smoothing(y: torch.Tensor, smooth_factor: float) -> torch.Tensor:
    """batch_preprocessing.py:15-28 (binary labels)."""
    return torch.where(y < 0.5, y + smooth_factor, y - smooth_factor)


def _mix_batch(x: torch.Tensor, factor: np.ndarray, perm: np.ndarray, sel: np.ndarray | None) -> torch.Tensor:
    """``out[b] = x[b]*f[b] + x[perm[b]]*(1-f[b])`` where ``sel[b]`` (all if None), else ``x[b]``.

    The kernel computes in fp32.  The reference mixes any floating dtype in that dtype: fp16 / bf16 / fp64
    batches are cast to fp32, mixed and cast back (for the half types that is MORE precise than the
    reference's three half-precision roundings, documented deviation; fp64 loses precision and is refused)."""
    if x.device.type != "cuda":
        raise RuntimeError("adell_mri_b200.batch_preprocessing runs on CUDA tensors only (no CPU fallback)")
    if x.dtype == torch.float64 or not x.dtype.is_floating_point:
        raise TypeError(f"mixup expects a float32 / float16 / bfloat16 batch, got {x.dtype}")
    in_dtype = x.dtype
    x = x.contiguous() if in_dtype == torch.float32 else x.float().contiguous()
    B = x.shape[0]
    per = x.numel() // max(B, 1)
    out = torch.empty_like(x)
    # per-sample parameters {factor fp32 (cast like the reference: to x.dtype), perm int32, selected u8}:
    # one upload through the engine's pinned staging ring
    from . import engine

    stride = (4 * B + 127) // 128 * 128
    buf = np.zeros(3 * stride, np.uint8)
    buf[: 4 * B].view(np.float32)[:] = torch.as_tensor(np.asarray(factor, np.float64), dtype=in_dtype).float().numpy()
    buf[stride : stride + 4 * B].view(np.int32)[:] = np.asarray(perm, np.int32)
    if sel is not None:
        buf[2 * stride : 2 * stride + B] = np.asarray(sel, np.uint8)
    with torch.cuda.device(x.device):
        dev = engine._stage(buf, x.device)
        base = dev.data_ptr()
        stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        _lib.check(_lib.load().adell_mixup(x.data_ptr(), out.data_ptr(), base, base + stride,
                                           None if sel is None else base + 2 * stride, B, per, stream), "adell_mixup")
    # (`dev` may be freed now: torch's allocator reuses it in stream order, after the kernel above)
    return out if in_dtype == torch.float32 else out.to(in_dtype)


def mixup(x: torch.Tensor, y: torch.Tensor, mixup_alpha: float, g: np.random.Generator = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """batch_preprocessing.py:31-65."""
    batch_size = y.shape[0]
    if g is None:
        g = np.random.default_rng()
    factor = g.beta(mixup_alpha, mixup_alpha, batch_size)
    mixup_factor = torch.as_tensor(factor, dtype=x.dtype, device=x.device)
    mixup_perm = g.permutation(batch_size)
    x = _mix_batch(x, factor, mixup_perm, None)
    y = y * mixup_factor + y[mixup_perm] * (1.0 - mixup_factor)
    return x, y


def partial_mixup(x: torch.Tensor, y: torch.Tensor, mixup_alpha: float, mixup_fraction: float = 0.5,
                  g: np.random.Generator = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """batch_preprocessing.py:68-118.  Like the reference, ``x`` and ``y`` are updated IN PLACE (the
    reference assigns ``x[mxu_i] = ...``) and returned: a caller that keeps its own handle on the batch
    sees the mixed data."""
    batch_size = y.shape[0]
    if g is None:
        g = np.random.default_rng()
    mxu_i = g.binomial(1, mixup_fraction, batch_size).astype(bool)
    sel_factor = g.beta(mixup_alpha, mixup_alpha, mxu_i.sum())
    mixup_factor = torch.as_tensor(sel_factor, dtype=x.dtype, device=x.device)
    mixup_perm = g.permutation(batch_size)
    factor = np.ones(batch_size, np.float64)
    factor[mxu_i] = sel_factor
    x.copy_(_mix_batch(x, factor, mixup_perm, mxu_i))
    y[mxu_i] = torch.add(y[mxu_i] * mixup_factor, y[mixup_perm][mxu_i] * (1 - mixup_factor))
    return x, y


class BatchPreprocessing:
    """batch_preprocessing.py:121-181: label smoothing, then (partial) mixup."""

    def __init__(self, label_smoothing: float = None, mixup_alpha: float = None, partial_mixup: float = None, seed: int = 42):
        self.label_smoothing = label_smoothing
        self.mixup_alpha = mixup_alpha
        self.partial_mixup = partial_mixup
        self.seed = seed
        if self.mixup_alpha is not None:
            self.g = np.random.default_rng(seed)

    def __call__(self, X: torch.Tensor, y: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        if self.label_smoothing is not None:
            y = label_smoothing(y, self.label_smoothing)
        if self.mixup_alpha is not None:
            initial_y_dtype = y.dtype
            y = y.float()
            if self.partial_mixup is not None:
                X, y = partial_mixup(X, y, self.mixup_alpha, self.partial_mixup, self.g)
            else:
                X, y = mixup(X, y, self.mixup_alpha, self.g)
            y = y.to(initial_y_dtype)
        return X, y
