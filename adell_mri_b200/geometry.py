"""Host-side affine composition, batched, in MONAI's exact fp32 operation order.

MONAI's ``AffineGrid`` builds ``eye(4) @ rotate @ shear @ translate @ scale`` with torch fp32
ops (sin/cos of the fp32-cast angle; 3-D rotate = Rx @ Ry @ Rz with only the leading
rotations if fewer than three angles are given).  This module does the same for a whole
batch at once (``[B, 4, 4]`` torch CPU tensors) so that the matrices handed to K1 are the
ones the reference's RandAffined would have produced for the same parameter draws
(wiring: /root/reference/adell_mri/transform_factory/augmentations.py:98-116,279-301;
/root/reference/adell_mri/modules/augmentations.py:131-186).
"""

from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from . import _lib


def _eye(b: int) -> np.ndarray:
    return np.tile(np.eye(4, dtype=np.float32), (b, 1, 1))


def _as_param(x, b: int) -> np.ndarray | None:
    """``[B, k]`` float32 array (k = number of parameters given) or None when empty.  The cast to
    fp32 is the one ``torch.as_tensor(th, dtype=torch.float32)`` performs in MONAI."""
    if x is None:
        return None
    a = np.asarray(x, dtype=np.float64)
    if a.size == 0:
        return None
    if a.ndim == 1:
        a = np.broadcast_to(a, (b, a.shape[0]))
    return np.ascontiguousarray(a).astype(np.float32)


def rotate_factors(radians: np.ndarray) -> list[np.ndarray]:
    """The Rx, Ry, Rz factors (leading ones only) — multiplied left to right by the caller.
    sin / cos are evaluated by torch in fp32 exactly as MONAI's ``create_rotate`` does."""
    b, k = radians.shape
    t = torch.from_numpy(radians)
    s, c = torch.sin(t).numpy(), torch.cos(t).numpy()
    out = []
    if k >= 1:
        m = _eye(b)
        m[:, 1, 1], m[:, 1, 2] = c[:, 0], -s[:, 0]
        m[:, 2, 1], m[:, 2, 2] = s[:, 0], c[:, 0]
        out.append(m)
    if k >= 2:
        m = _eye(b)
        m[:, 0, 0], m[:, 0, 2] = c[:, 1], s[:, 1]
        m[:, 2, 0], m[:, 2, 2] = -s[:, 1], c[:, 1]
        out.append(m)
    if k >= 3:
        m = _eye(b)
        m[:, 0, 0], m[:, 0, 1] = c[:, 2], -s[:, 2]
        m[:, 1, 0], m[:, 1, 1] = s[:, 2], c[:, 2]
        out.append(m)
    return out


def shear_matrices(coefs: np.ndarray) -> np.ndarray:
    b, k = coefs.shape
    c = np.zeros((b, 6), np.float32)
    c[:, : min(k, 6)] = coefs[:, :6]
    out = _eye(b)
    out[:, 0, 1], out[:, 0, 2] = c[:, 0], c[:, 1]
    out[:, 1, 0], out[:, 1, 2] = c[:, 2], c[:, 3]
    out[:, 2, 0], out[:, 2, 1] = c[:, 4], c[:, 5]
    return out


def translate_matrices(shift: np.ndarray) -> np.ndarray:
    b, k = shift.shape
    out = _eye(b)
    out[:, : min(k, 3), 3] = shift[:, :3]
    return out


def scale_matrices(factors: np.ndarray) -> np.ndarray:
    b, k = factors.shape
    out = _eye(b)
    for i in range(min(k, 3)):
        out[:, i, i] = factors[:, i]
    return out


def compose_affine(rotate=None, shear=None, translate=None, scale=None, batch: int | None = None) -> np.ndarray:
    """``[B, 4, 4]`` float32 MONAI AffineGrid matrices for a batch of parameter sets.
    Every argument is ``[B, k]`` (or ``[k]``, broadcast) or None/empty."""
    if batch is None:
        batch = 1
        for x in (rotate, shear, translate, scale):
            if x is not None and np.asarray(x).ndim == 2:
                batch = np.asarray(x).shape[0]
    r, sh, t, sc = (_as_param(x, batch) for x in (rotate, shear, translate, scale))
    out = np.empty((batch, 4, 4), np.float32)
    if batch == 0:
        return out
    sin_r = cos_r = None
    if r is not None:
        if r.shape[1] > 3:
            r = np.ascontiguousarray(r[:, :3])
        # fp32, evaluated by torch exactly as MONAI's create_rotate does.  In slices of <= 1024 elements: above a few
        # thousand elements torch hands the call to a threaded vector-math region, whose worker threads cost milliseconds
        # to wake when the host's cores are shared (8 ranks per box; 20-50 ms per call measured in a CPU-limited
        # container) — per element the result is the same.
        tr = torch.from_numpy(np.ascontiguousarray(r)).reshape(-1)
        sin_t, cos_t = torch.empty_like(tr), torch.empty_like(tr)
        for o in range(0, tr.numel(), 1024):
            torch.sin(tr[o:o + 1024], out=sin_t[o:o + 1024])
            torch.cos(tr[o:o + 1024], out=cos_t[o:o + 1024])
        sin_r, cos_r = sin_t.numpy().reshape(r.shape), cos_t.numpy().reshape(r.shape)
    p = lambda a: None if a is None else a.ctypes.data
    k = lambda a: 0 if a is None else a.shape[1]
    _lib.check(_lib.load().adell_affine_compose(p(sin_r), p(cos_r), k(r), p(sh), k(sh), p(t), k(t), p(sc), k(sc), batch, out.ctypes.data),
               "adell_affine_compose")
    return out
