"""Host composer: turns a chain of reference transforms into canonical ``adell_item`` arrays.

The reference executes every MONAI dictionary transform eagerly, one full-volume copy per
op (/root/reference/adell_mri/transform_factory/augmentations.py:98-176,255-301,427-515).
Here the ops are *recorded* on a :class:`BatchPlan` (n independent volumes, all state held in
``[n, 3]`` numpy arrays so one call composes a whole batch) and collapsed into the canonical
form consumed by K1 (see ``include/adell_b200.h``)::

    parent -> integer ops (crop / pad / flip)  -> resample domain S
           -> ONE affine resample              -> grid G
           -> integer ops                      -> output O  -> post intensity map -> noise

Flip / crop / pad stay integer index transforms on either side of the resample (never folded
into the fp32 matrix), which keeps masks bit-exact.  A second resample, or any op after a
noise tensor, cannot be represented in one pass: the current pass is *closed* (it will be
materialised into a scratch volume by one extra launch) — the reference's own sequential
semantics, i.e. "parity mode".  ``fast=True`` composes consecutive affines into one matrix
instead (documented deviation: no double interpolation).
"""

from __future__ import annotations

import ctypes
from typing import Sequence

import numpy as np
import torch

from . import _lib
from ._lib import Item

ITEM_DTYPE = np.dtype(Item)
assert ITEM_DTYPE.itemsize == _lib.ITEM_SIZE

_TORCH_TO_ADELL = {torch.float32: _lib.F32, torch.int16: _lib.I16, torch.uint8: _lib.U8}
_ELSIZE = {_lib.F32: 4, _lib.I16: 2, _lib.U8: 1}


def _i3(x, n):
    a = np.asarray(x, dtype=np.int64)
    if a.ndim == 1:
        a = np.broadcast_to(a, (n, 3))
    return a.astype(np.int64, copy=True)


def _where(where, n):
    if where is None:
        return np.ones(n, dtype=bool)
    w = np.asarray(where, dtype=bool)
    if w.ndim == 0:
        w = np.broadcast_to(w, (n,))
    return w.copy()


class IntMap:
    """Batch of per-axis integer index maps: index ``t`` of the *current* space ``[0,size)``
    reads underlying index ``off + sign*t`` and is valid iff ``vlo <= t < vhi`` (else literal 0).
    """

    __slots__ = ("size", "off", "sign", "vlo", "vhi")

    def __init__(self, size: np.ndarray):
        self.size = size.astype(np.int64, copy=True)
        self.off = np.zeros_like(self.size)
        self.sign = np.ones_like(self.size)
        self.vlo = np.zeros_like(self.size)
        self.vhi = self.size.copy()

    def take(self, idx):
        m = IntMap.__new__(IntMap)
        for f in self.__slots__:
            setattr(m, f, getattr(self, f)[idx].copy())
        return m

    def put(self, idx, other: "IntMap"):
        for f in self.__slots__:
            getattr(self, f)[idx] = getattr(other, f)

    @staticmethod
    def concat(maps: Sequence["IntMap"]):
        m = IntMap.__new__(IntMap)
        for f in IntMap.__slots__:
            setattr(m, f, np.concatenate([getattr(x, f) for x in maps], axis=0))
        return m

    def crop(self, start, new_size, w):
        start = np.where(w[:, None], start, 0)
        new_size = np.where(w[:, None], new_size, self.size)
        self.off = self.off + self.sign * start
        self.vlo = np.clip(self.vlo - start, 0, new_size)
        self.vhi = np.clip(self.vhi - start, 0, new_size)
        self.size = new_size

    def pad(self, before, after, w):
        before = np.where(w[:, None], before, 0)
        after = np.where(w[:, None], after, 0)
        self.off = self.off - self.sign * before
        self.vlo = self.vlo + before
        self.vhi = self.vhi + before
        self.size = self.size + before + after

    def flip(self, mask):
        # t = size-1-t'
        self.off = np.where(mask, self.off + self.sign * (self.size - 1), self.off)
        new_vlo = np.where(mask, self.size - self.vhi, self.vlo)
        new_vhi = np.where(mask, self.size - self.vlo, self.vhi)
        self.sign = np.where(mask, -self.sign, self.sign)
        self.vlo, self.vhi = new_vlo, new_vhi

    def has_invalid(self):
        return ((self.vlo > 0) | (self.vhi < self.size)).any(axis=1)


class _Stage:
    """State arrays of one (open or closed) pass for a batch of volumes."""

    _ARR = (
        "has_affine", "A", "interp", "padding", "pre_s", "pre_o", "clip", "clip_lo", "clip_hi",
        "post_s", "post_o", "noise_ptr", "philox_std", "philox_seed", "philox_off", "strict", "pre_dev", "grid",
        "win_dev", "win_parent",
    )

    def __init__(self, size: np.ndarray):
        n = size.shape[0]
        self.pre = IntMap(size)
        self.post = IntMap(size)  # meaningful only where has_affine
        self.has_affine = np.zeros(n, bool)
        self.A = np.tile(np.eye(4, dtype=np.float32)[:3], (n, 1, 1))
        self.interp = np.full(n, _lib.TRILINEAR, np.uint8)
        self.padding = np.full(n, _lib.PAD_ZEROS, np.uint8)
        self.pre_s = np.ones(n, np.float64)
        self.pre_o = np.zeros(n, np.float64)
        self.clip = np.zeros(n, bool)
        self.clip_lo = np.zeros(n, np.float32)
        self.clip_hi = np.zeros(n, np.float32)
        self.post_s = np.ones(n, np.float64)
        self.post_o = np.zeros(n, np.float64)
        self.noise_ptr = np.zeros(n, np.uint64)
        self.philox_std = np.zeros(n, np.float32)
        self.philox_seed = np.zeros(n, np.uint64)
        self.philox_off = np.zeros(n, np.uint64)
        self.strict = np.zeros(n, bool)
        self.pre_dev = np.zeros(n, np.uint64)
        self.grid = size.astype(np.int64, copy=True)  # output grid of the resample (where has_affine): the source size for
        #                                               RandAffined(spatial_size=None), the new size for a Spacingd resample
        self.win_dev = np.zeros(n, np.uint64)         # device int32[3]: start of a crop window chosen on the device
        self.win_parent = np.zeros((n, 3), np.int64)  # extents of the volume that window lives in
        self.keep = []  # tensors that must outlive the launch (noise)

    def take(self, idx):
        s = _Stage.__new__(_Stage)
        s.pre, s.post = self.pre.take(idx), self.post.take(idx)
        for f in self._ARR:
            setattr(s, f, getattr(self, f)[idx].copy())
        s.keep = list(self.keep)
        return s

    def reset(self, idx, size):
        fresh = _Stage(size)
        self.pre.put(idx, fresh.pre)
        self.post.put(idx, fresh.post)
        for f in self._ARR:
            getattr(self, f)[idx] = getattr(fresh, f)

    @staticmethod
    def concat(stages: Sequence["_Stage"]):
        s = _Stage.__new__(_Stage)
        s.pre = IntMap.concat([x.pre for x in stages])
        s.post = IntMap.concat([x.post for x in stages])
        for f in _Stage._ARR:
            setattr(s, f, np.concatenate([getattr(x, f) for x in stages], axis=0))
        s.keep = [k for x in stages for k in x.keep]
        return s

    def out_size(self):
        return np.where(self.has_affine[:, None], self.post.size, self.pre.size)


class BatchPlan:
    """``n`` volumes with their pending (recorded, not yet executed) transform chains."""

    @classmethod
    def from_arrays(cls, ptr, stride, dtype, shape, device, keep, fast=False, strict=False) -> "BatchPlan":
        """Build from precomputed per-volume metadata (``ptr`` uint64 [n], ``stride`` int64 [n,3],
        ``dtype`` uint8 [n], ``shape`` int64 [n,3]) — skips the per-tensor Python work."""
        self = cls.__new__(cls)
        self.n = int(ptr.shape[0])
        self.fast = fast
        self.keep = list(keep)
        self.parent_ptr = np.asarray(ptr, np.uint64).copy()
        self.parent_stride = np.asarray(stride, np.int64).copy()
        self.parent_dtype = np.asarray(dtype, np.uint8).copy()
        self.device = device
        self.st = _Stage(np.asarray(shape, np.int64))
        if strict:
            self.st.strict[:] = True
        self.default_strict = strict
        self.passes = []
        return self

    def __init__(self, parents: Sequence[torch.Tensor], fast: bool = False, strict: bool = False):
        n = len(parents)
        self.n = n
        self.fast = fast
        self.keep = list(parents)
        self.parent_ptr = np.zeros(n, np.uint64)
        self.parent_stride = np.zeros((n, 3), np.int64)
        self.parent_dtype = np.zeros(n, np.uint8)
        size = np.zeros((n, 3), np.int64)
        dev = None
        for i, p in enumerate(parents):
            if p.dim() != 3:
                raise ValueError("BatchPlan parents must be [H, W, D] volumes")
            if p.dtype not in _TORCH_TO_ADELL:
                raise TypeError(f"unsupported source dtype {p.dtype}")
            dev = p.device if dev is None else dev
            if p.device != dev:
                raise ValueError("all parents of a plan must live on one device")
            self.parent_ptr[i] = p.data_ptr()
            self.parent_stride[i] = p.stride()
            self.parent_dtype[i] = _TORCH_TO_ADELL[p.dtype]
            size[i] = p.shape
        self.device = dev
        self.st = _Stage(size)
        if strict:
            self.st.strict[:] = True
        self.default_strict = strict
        # closed passes: (idx, stage, parent_ptr, parent_stride, parent_dtype); a zero parent_ptr is
        # rewired to the scratch volume of the same volume's previous pass at execute time
        self.passes: list[tuple] = []

    # ------------------------------------------------------------------ bookkeeping
    @property
    def shape(self) -> np.ndarray:
        """Current logical spatial size ``[n, 3]``."""
        return self.st.out_size()

    @staticmethod
    def concat(plans: Sequence["BatchPlan"]) -> "BatchPlan":
        out = BatchPlan.__new__(BatchPlan)
        out.n = sum(p.n for p in plans)
        out.fast = any(p.fast for p in plans)
        out.default_strict = all(p.default_strict for p in plans)
        out.keep = [k for p in plans for k in p.keep]
        out.device = plans[0].device
        for p in plans:
            if p.device != out.device:
                raise ValueError("cannot collate plans living on different devices")
        out.parent_ptr = np.concatenate([p.parent_ptr for p in plans])
        out.parent_stride = np.concatenate([p.parent_stride for p in plans])
        out.parent_dtype = np.concatenate([p.parent_dtype for p in plans])
        out.st = _Stage.concat([p.st for p in plans])
        offs = np.cumsum([0] + [p.n for p in plans])
        depth = max(len(p.passes) for p in plans)
        out.passes = []
        for lvl in range(depth):
            parts = [(o, p.passes[lvl]) for o, p in zip(offs, plans) if lvl < len(p.passes)]
            out.passes.append(
                (
                    np.concatenate([q[0] + o for o, q in parts]),
                    _Stage.concat([q[1] for _, q in parts]),
                    np.concatenate([q[2] for _, q in parts]),
                    np.concatenate([q[3] for _, q in parts]),
                    np.concatenate([q[4] for _, q in parts]),
                )
            )
        return out

    def permuted(self, perm: np.ndarray) -> "BatchPlan":
        """The same plan with its volumes reordered: volume ``i`` of the result is volume ``perm[i]`` of this one."""
        perm = np.asarray(perm, np.int64)
        inv = np.empty_like(perm)
        inv[perm] = np.arange(perm.shape[0])
        out = BatchPlan.__new__(BatchPlan)
        out.n, out.fast, out.default_strict, out.device = self.n, self.fast, self.default_strict, self.device
        out.keep = list(self.keep)
        out.parent_ptr, out.parent_stride, out.parent_dtype = self.parent_ptr[perm], self.parent_stride[perm], self.parent_dtype[perm]
        out.st = self.st.take(perm)
        out.passes = [(inv[idx], st, a, b, c) for idx, st, a, b, c in self.passes]
        return out

    def _close(self, mask: np.ndarray):
        """Close the open pass of the volumes in ``mask``: it will be materialised into a
        scratch fp32 volume which becomes their new parent."""
        idx = np.nonzero(mask)[0]
        if idx.size == 0:
            return
        st = self.st.take(idx)
        # parent meta as of this pass (ptr == 0: scratch written by the volume's previous pass)
        self.passes.append(
            (idx, st, self.parent_ptr[idx].copy(), self.parent_stride[idx].copy(), self.parent_dtype[idx].copy())
        )
        size = st.out_size()
        self.st.reset(idx, size)
        if self.default_strict:
            self.st.strict[idx] = True
        self.parent_ptr[idx] = 0  # filled in at execute time
        self.parent_stride[idx] = _contig_stride(size)
        self.parent_dtype[idx] = _lib.F32

    def _has_noise(self):
        return (self.st.noise_ptr != 0) | (self.st.philox_std != 0)

    # ------------------------------------------------------------------ integer ops
    def _int_op(self, w, fn_name, *args):
        self._close(w & self._has_noise())
        st = self.st
        getattr(st.pre, fn_name)(*args, w & ~st.has_affine)
        getattr(st.post, fn_name)(*args, w & st.has_affine)

    def crop(self, start, size, where=None):
        """SpatialCrop: keep ``[start, start+size)`` (clipped to the current extent)."""
        w = _where(where, self.n)
        cur = self.shape
        start = np.clip(_i3(start, self.n), 0, cur)
        size = np.minimum(_i3(size, self.n), cur - start)
        self._int_op(w, "crop", start, size)
        return self

    def crop_from_device(self, win_dev: torch.Tensor, size, where=None):
        """SpatialCrop whose START lives in device memory (``win_dev``: contiguous ``[n, 3]`` int32, one row per
        volume; e.g. written by ``adell_posneg_starts``): the volumes become windows of ``size`` whose position K1
        reads when it fetches their items (``ADELL_F_WIN_DEV``).  Must be the first recorded op of those volumes."""
        w = _where(where, self.n)
        if isinstance(win_dev, torch.Tensor):
            if win_dev.shape != (self.n, 3) or win_dev.dtype != torch.int32 or not win_dev.is_contiguous():
                raise ValueError("win_dev must be a contiguous [n, 3] int32 tensor")
            self.keep.append(win_dev)
            ptrs = win_dev.data_ptr() + 12 * np.arange(self.n, dtype=np.uint64)
        else:   # per-volume device addresses of int32[3] rows (the caller keeps the tensors alive)
            ptrs = np.asarray(win_dev, np.uint64).reshape(self.n)
        st = self.st
        touched = st.has_affine | st.pre.has_invalid() | (st.pre.off != 0).any(1) | (st.pre.sign != 1).any(1) | (st.win_dev != 0)
        if (w & touched).any():
            raise ValueError("a device-side crop window must be the first recorded op")
        cur = st.pre.size
        size = np.minimum(_i3(size, self.n), cur)
        st.win_parent = np.where(w[:, None], cur, st.win_parent)
        st.win_dev = np.where(w, ptrs, st.win_dev).astype(np.uint64)
        st.pre.crop(np.zeros_like(size), size, w)
        return self

    def center_crop(self, roi, where=None):
        """CenterSpatialCrop †: start = max(⌊size/2⌋ - ⌊roi/2⌋, 0)."""
        cur = self.shape
        roi = _i3(roi, self.n)
        roi = np.where(roi <= 0, cur, np.minimum(roi, cur))
        start = np.maximum(cur // 2 - roi // 2, 0)
        return self.crop(start, roi, where)

    def pad(self, before, after, where=None):
        w = _where(where, self.n)
        # zeros injected by a pad must stay literal zeros: a pending additive post offset forbids that
        self._close(w & (self.st.post_o != 0))
        self._int_op(w, "pad", _i3(before, self.n), _i3(after, self.n))
        return self

    def spatial_pad(self, spatial_size, where=None):
        """SpatialPad(symmetric, constant 0) †: before=⌊w/2⌋, after=w-⌊w/2⌋."""
        cur = self.shape
        want = _i3(spatial_size, self.n)
        width = np.where(want > 0, np.maximum(want - cur, 0), 0)
        return self.pad(width // 2, width - width // 2, where)

    def flip(self, axes_mask, where=None):
        """``torch.flip`` over the axes set in ``axes_mask`` ([n,3] or [3] bool)."""
        w = _where(where, self.n)
        m = np.asarray(axes_mask, dtype=bool)
        if m.ndim == 1:
            m = np.broadcast_to(m, (self.n, 3))
        m = m & w[:, None]
        self._close(m.any(axis=1) & self._has_noise())
        st = self.st
        st.pre.flip(m & ~st.has_affine[:, None])
        st.post.flip(m & st.has_affine[:, None])
        return self

    # ------------------------------------------------------------------ resample
    def affine(self, A, mode="bilinear", padding_mode="reflection", where=None, out_size=None):
        """One RandAffine firing: ``A`` is the fp32 4x4 MONAI matrix ([4,4] or [n,4,4]); the
        output grid is the current size (``spatial_size=None`` as the reference always uses) unless
        ``out_size`` ([3] or [n,3]) names another one (:meth:`resample_to`)."""
        w = _where(where, self.n)
        A = np.asarray(A, dtype=np.float32)
        if A.ndim == 2:
            A = np.broadcast_to(A, (self.n, 4, 4))
        if isinstance(mode, str):
            interp = np.full(self.n, _lib.INTERP_MODES[mode], np.uint8)
        else:
            interp = np.asarray([_lib.INTERP_MODES[m] for m in mode], np.uint8)
        padding = np.uint8(_lib.PADDING_MODES[padding_mode])
        st = self.st
        if self.fast and out_size is None:
            # compose into the pending matrix: M = M_prev @ A (only legal if nothing integer sits between)
            comp = w & st.has_affine & ~self._has_noise()
            untouched = (
                (st.post.off == 0).all(1) & (st.post.sign == 1).all(1) & (st.post.size == st.post.vhi).all(1)
                & (st.post.vlo == 0).all(1) & (st.post_s == 1) & (st.post_o == 0)
                # a crop that starts at 0 only shrinks post.size: the next affine then acts about the centre
                # of the CROPPED grid, which a matrix product on the uncropped grid does not express
                & (st.post.size == st.pre.size).all(1)
            )
            comp &= untouched
            if comp.any():
                M = np.tile(np.eye(4, dtype=np.float64), (self.n, 1, 1))
                M[:, :3] = st.A
                M = M @ A.astype(np.float64)
                st.A[comp] = M[comp, :3].astype(np.float32)
                w = w & ~comp
        self._close(w & (st.has_affine | self._has_noise()))
        # a pending post map precedes this resample: fold it into the per-tap pre map
        foldable = w & ((st.post_s != 1) | (st.post_o != 0))
        # (the kernel applies pre -> clip -> post per tap: folding post into pre would move it before the clip)
        bad = foldable & ((st.pre.has_invalid() & (st.post_o != 0)) | (st.pre_dev != 0) | st.clip)
        self._close(bad)
        foldable &= ~bad
        st.pre_o = np.where(foldable, st.pre_o * st.post_s + st.post_o, st.pre_o)
        st.pre_s = np.where(foldable, st.pre_s * st.post_s, st.pre_s)
        st.post_s = np.where(foldable, 1.0, st.post_s)
        st.post_o = np.where(foldable, 0.0, st.post_o)
        st.has_affine = st.has_affine | w
        st.A[w] = A[w, :3]
        st.interp = np.where(w, interp, st.interp).astype(np.uint8)
        st.padding = np.where(w, padding, st.padding).astype(np.uint8)
        grid = st.pre.size if out_size is None else _i3(out_size, self.n)
        st.grid = np.where(w[:, None], grid, st.grid)
        fresh = IntMap(grid)
        idx = np.nonzero(w)[0]
        st.post.put(idx, fresh.take(idx))
        return self

    def resample_to(self, step, out_size, mode="bilinear", padding_mode="border", where=None):
        """Axis-aligned resample onto a grid of another size: output voxel ``o`` reads source index ``step * o``
        per axis (``step`` [3] or [n,3] float64) — ``monai.transforms.Spacingd`` on an axis-aligned volume †
        (``step = pixdim / spacing``; SpatialResample keeps voxel 0 in place and maps index to index with
        ``align_corners=False``).  Expressed as a MONAI-style affine for K1: with the centred grid coordinate
        ``c = o - (G-1)/2`` the kernel computes ``u = A c + t + (S-1)/2``, hence ``A = diag(step)`` and
        ``t = step (G-1)/2 - (S-1)/2``."""
        cur = self.shape.astype(np.float64)
        G = _i3(out_size, self.n).astype(np.float64)
        step = np.asarray(step, np.float64)
        if step.ndim == 1:
            step = np.broadcast_to(step, (self.n, 3))
        A = np.tile(np.eye(4, dtype=np.float32), (self.n, 1, 1))
        for a in range(3):
            A[:, a, a] = step[:, a].astype(np.float32)
            A[:, a, 3] = (step[:, a] * (G[:, a] - 1) / 2 - (cur[:, a] - 1) / 2).astype(np.float32)
        return self.affine(A, mode, padding_mode, where=where, out_size=G.astype(np.int64))

    # ------------------------------------------------------------------ intensity
    def intensity(self, scale=1.0, offset=0.0, where=None):
        """``v*scale + offset`` applied to the current (logical) volume."""
        w = _where(where, self.n)
        scale = np.broadcast_to(np.asarray(scale, np.float64), (self.n,))
        offset = np.broadcast_to(np.asarray(offset, np.float64), (self.n,))
        # a device-side pre map ({scale, offset} read from pre_dev by the kernel) cannot absorb a
        # host-side one: materialise that pass first
        self._close(w & (self._has_noise() | (self.st.pre_dev != 0)))
        st = self.st
        to_post = w & (st.has_affine | st.pre.has_invalid() | st.clip)
        to_pre = w & ~to_post
        st.pre_o = np.where(to_pre, st.pre_o * scale + offset, st.pre_o)
        st.pre_s = np.where(to_pre, st.pre_s * scale, st.pre_s)
        st.post_o = np.where(to_post, st.post_o * scale + offset, st.post_o)
        st.post_s = np.where(to_post, st.post_s * scale, st.post_s)
        return self

    def intensity_from_device(self, pre_dev: torch.Tensor):
        """Per-volume ``{scale, offset}`` computed on the device by the statistics kernels
        (``pre_dev`` is ``[n, 2]`` fp32); only legal on untouched volumes."""
        if pre_dev.shape != (self.n, 2) or pre_dev.dtype != torch.float32 or not pre_dev.is_contiguous():
            raise ValueError("pre_dev must be a contiguous [n, 2] float32 tensor")
        st = self.st
        if st.has_affine.any() or st.pre.has_invalid().any() or (st.pre_s != 1).any() or (st.pre_o != 0).any():
            raise ValueError("device-side intensity must be the first recorded op")
        self.keep.append(pre_dev)
        st.pre_dev[:] = pre_dev.data_ptr() + 8 * np.arange(self.n, dtype=np.uint64)
        return self

    def clip(self, lo, hi, where=None):
        w = _where(where, self.n)
        st = self.st
        self._close(w & (st.has_affine | st.pre.has_invalid() | self._has_noise() | (st.post_s != 1) | (st.post_o != 0)))
        st.clip |= w
        st.clip_lo = np.where(w, np.float32(lo), st.clip_lo).astype(np.float32)
        st.clip_hi = np.where(w, np.float32(hi), st.clip_hi).astype(np.float32)
        return self

    def add_noise(self, noise: Sequence[torch.Tensor | None]):
        """Add injected fp32 noise tensors (contiguous, current logical shape) — parity path of
        RandGaussianNoised (the reference draws it from RandomState.normal on the host)."""
        if len(noise) != self.n:
            raise ValueError("one noise tensor (or None) per volume")
        cur = self.shape
        self._close(np.array([t is not None for t in noise]) & self._has_noise())
        for i, t in enumerate(noise):
            if t is None:
                continue
            if tuple(t.shape[-3:]) != tuple(int(x) for x in cur[i]) or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("noise must be contiguous float32 of the current volume shape")
            if t.device != self.device:
                raise ValueError("noise must live on the plan's device")
            self.st.noise_ptr[i] = t.data_ptr()
            self.st.keep.append(t)
        return self

    def add_philox_noise(self, std, seed, offset=0, where=None):
        w = _where(where, self.n)
        self._close(w & self._has_noise())
        st = self.st
        st.philox_std = np.where(w, np.broadcast_to(np.asarray(std, np.float32), (self.n,)), st.philox_std).astype(np.float32)
        st.philox_seed = np.where(w, np.uint64(seed), st.philox_seed).astype(np.uint64)
        st.philox_off = np.where(w, np.broadcast_to(np.asarray(offset, np.uint64), (self.n,)), st.philox_off).astype(np.uint64)
        return self

    # ------------------------------------------------------------------ items
    @staticmethod
    def _fill_items(st: _Stage, parent_ptr, parent_stride, parent_dtype, dst_ptr, dst_stride) -> np.ndarray:
        n = parent_ptr.shape[0]
        it = np.zeros(n, ITEM_DTYPE)
        elsize = np.array([_ELSIZE[int(d)] for d in parent_dtype], np.int64) if n else np.zeros(0, np.int64)
        byte_off = (st.pre.off * parent_stride).sum(axis=1) * elsize
        it["src"] = (parent_ptr.astype(np.int64) + byte_off).astype(np.uint64)
        it["src_stride"] = st.pre.sign * parent_stride
        it["src_shape"] = st.pre.size
        it["src_vlo"] = st.pre.vlo
        win = st.win_dev != 0
        if win.any():
            if (st.pre.has_invalid() & win).any():
                raise ValueError("a device-side crop window cannot be padded before the resample")
            # ADELL_F_WIN_DEV: src_vhi carries the parent's extent counted from the window's lowest element at start 0
            lo = np.where(st.pre.sign > 0, st.pre.off, st.pre.off - (st.pre.size - 1))
            it["src_vhi"] = np.where(win[:, None], st.win_parent - lo, st.pre.vhi)
        else:
            it["src_vhi"] = st.pre.vhi
        it["win_dev"] = st.win_dev
        ha = st.has_affine[:, None]
        it["out_shape"] = np.where(ha, st.post.size, st.pre.size)
        it["grid_shape"] = np.where(ha, st.grid, st.pre.size)  # spatial_size=None: the affine grid has the source size
        it["grid_off"] = np.where(ha, st.post.off, 0)
        it["grid_sign"] = np.where(ha, st.post.sign, 1)
        it["out_vlo"] = np.where(ha, st.post.vlo, 0)
        it["out_vhi"] = np.where(ha, st.post.vhi, st.pre.size)
        it["A"] = st.A.reshape(n, 12)
        it["nrm"] = (2.0 / np.maximum(2, st.pre.size)).astype(np.float32)
        it["pre_scale"] = st.pre_s.astype(np.float32)
        it["pre_offset"] = st.pre_o.astype(np.float32)
        it["clip_lo"], it["clip_hi"] = st.clip_lo, st.clip_hi
        it["post_scale"] = st.post_s.astype(np.float32)
        it["post_offset"] = st.post_o.astype(np.float32)
        it["noise"] = st.noise_ptr
        it["noise_std"] = st.philox_std
        it["philox_seed"], it["philox_offset"] = st.philox_seed, st.philox_off
        it["pre_dev"] = st.pre_dev
        it["src_dtype"] = parent_dtype
        it["interp"], it["padding"] = st.interp, st.padding
        flags = np.zeros(n, np.uint8)
        flags |= np.where(~st.has_affine, _lib.F_IDENTITY, 0).astype(np.uint8)
        flags |= np.where(st.clip, _lib.F_CLIP, 0).astype(np.uint8)
        flags |= np.where(st.strict, _lib.F_STRICT, 0).astype(np.uint8)
        flags |= np.where(st.philox_std != 0, _lib.F_PHILOX, 0).astype(np.uint8)
        flags |= np.where(st.pre_dev != 0, _lib.F_PRE_DEV, 0).astype(np.uint8)
        flags |= np.where(win, _lib.F_WIN_DEV, 0).astype(np.uint8)
        it["flags"] = flags
        it["dst"] = dst_ptr
        it["dst_stride"] = dst_stride
        return it

    def build_launches(self, dst_ptr: np.ndarray, dst_stride: np.ndarray, alloc):
        """Resolve scratch volumes and return ``[items_array, ...]``, one entry per launch, in
        execution order.  ``alloc(n_elements) -> fp32 tensor`` provides scratch on the plan's device;
        the returned tensors are appended to ``self.keep``.

        Closed passes are regrouped by LEVEL: a volume's k-th closed pass goes into launch k, whatever
        other volumes were closed with it — volumes are independent, only the order of one volume's own
        passes matters.  A batch whose samples close passes at different moments (the SSL workhorse: every
        sample draws its own members) then needs as many launches as its deepest chain, not one per
        ``_close`` call."""
        cur_ptr = self.parent_ptr.copy()
        per_level: dict[int, list] = {}
        if self.passes:
            if alloc is None:
                raise ValueError("this plan needs scratch volumes (multi-pass) but no allocator was given")
            level = np.zeros(self.n, np.int64)
            # one scratch buffer for all passes; every scratch volume starts on a 256-byte boundary
            vox = [p[1].out_size().prod(axis=1) for p in self.passes]
            padded = [(v + 63) // 64 * 64 for v in vox]
            buf = alloc(int(sum(int(v.sum()) for v in padded)))
            self.keep.append(buf)
            base, off = buf.data_ptr(), 0
            for (idx, st, pptr, pstride, pdtype), pv in zip(self.passes, padded):
                offs = off + np.concatenate([[0], np.cumsum(pv)])
                tptr = (base + 4 * offs[:-1]).astype(np.uint64)
                src_ptr = np.where(pptr != 0, pptr, cur_ptr[idx])
                items = self._fill_items(st, src_ptr, pstride, pdtype, tptr, _contig_stride(st.out_size()))
                lv = level[idx]
                lo, hi = int(lv.min()), int(lv.max())
                if lo == hi:
                    per_level.setdefault(lo, []).append(items)
                else:
                    for l in range(lo, hi + 1):
                        m = lv == l
                        if m.any():
                            per_level.setdefault(l, []).append(items[m])
                level[idx] += 1
                cur_ptr[idx] = tptr
                off = int(offs[-1])
        # (concatenated as raw bytes: numpy re-derives a structured dtype on every concatenate)
        launches = [np.concatenate([x.view(np.uint8) for x in per_level[l]]).view(ITEM_DTYPE) if len(per_level[l]) > 1 else per_level[l][0]
                    for l in sorted(per_level)]
        launches.append(self._fill_items(self.st, cur_ptr, self.parent_stride, self.parent_dtype,
                                         np.asarray(dst_ptr, np.uint64), np.asarray(dst_stride, np.int64)))
        self.keep.extend(self.st.keep)
        for p in self.passes:
            self.keep.extend(p[1].keep)
        return launches


def _contig_stride(size: np.ndarray) -> np.ndarray:
    return np.stack([size[:, 1] * size[:, 2], size[:, 2], np.ones_like(size[:, 2])], axis=1)
