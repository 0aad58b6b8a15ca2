"""Batch-level mirrors of the reference's per-sample augmentation chains.

Each class reproduces, for a whole batch at once, the random chain that the reference
assembles per task and runs per sample in DataLoader workers, then collates:

* :class:`SegmentationBatchAugmenter` —
  ``SegmentationTransforms(...).transforms(get_augmentations_unet(...))`` + ``safe_collate[_crops]``
  (/root/reference/adell_mri/transform_factory/augmentations.py:19-178,
  /root/reference/adell_mri/transform_factory/transforms.py:206-263,
  /root/reference/adell_mri/entrypoints/segmentation/train.py:321-368,604-615)
* :class:`ClassificationBatchAugmenter` — ``get_augmentations_class`` + the ``crop+16`` margin /
  final centre crop / concat-with-mask of ``ClassificationTransforms``
  (augmentations.py:181-320, transforms.py:399-408,482-509)
* :class:`SSLBatchAugmenter` — ``get_augmentations_ssl`` two-view chain for the in-scope
  workhorse members (augmentations.py:391-516, /root/reference/adell_mri/modules/augmentations.py:10-256)

Random parameters are drawn on the host from ``numpy.random.RandomState`` streams in the
reference's draw order (see :mod:`adell_mri_b200.sampling`); geometry is composed on the
host into one canonical item per (sample, key) and the whole batch is ONE K1 launch per
resample pass, writing straight into the collated ``[B, C, H, W, D]`` output.
"""

from __future__ import annotations

import itertools
from typing import Sequence

import numpy as np
import torch

from . import engine, geometry
from .plan import BatchPlan
from .sampling import RandAffineSampler, child_seeds

UNET_AUGMENTS = ["intensity", "noise", "rbf", "affine", "shear", "flip", "blur", "distort", "lowres", "trivial"]
_GPU_AUGMENTS = {"affine", "shear", "flip"}


def _check_augment(augment, valid):
    for a in augment:
        if a not in valid:
            raise NotImplementedError("augment can only contain {}".format(valid))
    out_of_scope = [a for a in augment if a not in _GPU_AUGMENTS and a != "trivial"]
    if out_of_scope:
        raise NotImplementedError(
            f"augmentations {out_of_scope} are outside the fused GPU hot path (see DESIGN.md, out of scope)"
        )


class SegmentationBatchAugmenter:
    """``get_augmentations_unet`` chain for a batch of cached samples.

    ``samples`` passed to :meth:`__call__` are dicts ``key -> [1,H,W,D]`` device tensors
    (already intensity-scaled, i.e. what ``CacheDataset`` holds).  Output: ``{"image":
    [B, len(image_keys), ...], "mask": [B, 1, ...]}``.
    """

    def __init__(
        self,
        augment: Sequence[str],
        all_keys: Sequence[str],
        image_keys: Sequence[str],
        random_crop_size: Sequence[int] | None = None,
        has_label: bool = True,
        flip_axis: Sequence[int] = (0, 1),
        mask_key: str = "mask",
        output_image_key: str = "image",
        strict: bool = False,
        fast: bool = False,
    ):
        _check_augment(augment, UNET_AUGMENTS)
        if "trivial" in augment:
            raise NotImplementedError("'trivial' (OneOf) is handled by the dict-transform surface, not the batch fast path")
        self.image_keys = list(image_keys)
        self.mask_key = mask_key if has_label else None
        self.keys = list(self.image_keys) + ([mask_key] if has_label else [])
        self.modes = ["bilinear" if k in self.image_keys else "nearest" for k in self.keys]
        self.output_image_key = output_image_key
        self.random_crop_size = None if random_crop_size is None else [int(x) for x in random_crop_size]
        self.has_label = has_label
        self.strict, self.fast = strict, fast
        if self.random_crop_size is not None and has_label:
            raise NotImplementedError(
                "RandCropByPosNegLabeld (label-guided crop centres) is served by the dict-transform surface; "
                "the batch fast path implements the has_label=False RandSpatialCropd sandwich"
            )
        prob = 0.2
        # Randomizable children in Compose order: [crop], affine, shear, flips...
        self.samplers = []
        if "affine" in augment:
            self.samplers.append(RandAffineSampler(prob=prob, rotate_range=[np.pi / 8, np.pi / 8, np.pi / 16]))
        if "shear" in augment:
            self.samplers.append(RandAffineSampler(prob=prob, shear_range=((0.9, 1.1), (0.9, 1.1), (0.9, 1.1))))
        self.flip_axis = list(flip_axis) if "flip" in augment else []
        self.flip_R = [np.random.RandomState() for _ in self.flip_axis]
        self.crop_R = np.random.RandomState()
        self._meta = {}
        self._dst_cache = {}
        self.set_random_state(None)

    def set_random_state(self, seed=None):
        """Seed fan-out of ``Compose.set_random_state`` over the Randomizable children †."""
        n = (1 if self.random_crop_size is not None else 0) + len(self.samplers) + len(self.flip_axis)
        seeds = child_seeds(seed, n) if seed is not None else [None] * n
        i = 0
        if self.random_crop_size is not None:
            self.crop_R = np.random.RandomState(seeds[i]); i += 1
        for s in self.samplers:
            s.set_random_state(seeds[i]); i += 1
        for j in range(len(self.flip_axis)):
            self.flip_R[j] = np.random.RandomState(seeds[i]); i += 1
        return self

    def draw(self, batch: int, shape):
        """All host-side random parameters of one batch, sample by sample in stream order."""
        nk = len(self.keys)
        fired = np.zeros((len(self.samplers), batch), bool)
        mats = np.tile(np.eye(4, dtype=np.float32), (len(self.samplers), batch, 1, 1))
        flips = np.zeros((batch, 3), bool)
        starts = None
        if self.random_crop_size is not None:
            pre = [int(i * 1.10) for i in self.random_crop_size]
            pre = [min(p, s) for p, s in zip(pre, shape)]
            starts = np.zeros((batch, 3), np.int64)
        if starts is not None:
            # RandSpatialCropd(random_size=False): one randint per axis per sample, in order
            for b in range(batch):
                starts[b] = [self.crop_R.randint(s - p + 1) for s, p in zip(shape, pre)]
        for si, smp in enumerate(self.samplers):
            f, p = smp.draw_batch(batch, n_keys=nk)
            fired[si] = f
            if f.any():  # one batched composition per RandAffined (all draws share a parameter layout)
                mats[si, f] = geometry.compose_affine(p["rotate"], p["shear"], p["translate"], p["scale"],
                                                      batch=int(f.sum()))
        for j, ax in enumerate(self.flip_axis):
            flips[:, ax] ^= self.flip_R[j].random_sample(batch) < 0.25
        return dict(fired=fired, mats=mats, flips=flips, starts=starts)

    def _sample_meta(self, s: dict):
        """Per-sample volume metadata, cached on the sample dict's identity (a device-resident cache
        hands the same dicts back every epoch)."""
        m = self._meta.get(id(s))
        if m is None or m[0] is not s:
            vols = [s[k][0] for k in self.keys]
            plan = BatchPlan(vols)
            m = (s, plan.parent_ptr, plan.parent_stride, plan.parent_dtype, plan.shape, vols)
            self._meta[id(s)] = m
        return m

    def plan(self, samples: Sequence[dict], params=None) -> BatchPlan:
        B, nk = len(samples), len(self.keys)
        metas = [self._sample_meta(s) for s in samples]
        shape = tuple(int(x) for x in metas[0][4][0])
        if params is None:
            params = self.draw(B, shape)
        plan = BatchPlan.from_arrays(
            np.concatenate([m[1] for m in metas]), np.concatenate([m[2] for m in metas]),
            np.concatenate([m[3] for m in metas]), np.concatenate([m[4] for m in metas]),
            metas[0][5][0].device, [m[5] for m in metas], fast=self.fast, strict=self.strict)
        rep = lambda x: np.repeat(x, nk, axis=0)
        modes = self.modes * B
        if self.random_crop_size is not None:
            pre = [min(int(i * 1.10), s) for i, s in zip(self.random_crop_size, shape)]
            plan.crop(rep(params["starts"]), pre)
        for si in range(len(self.samplers)):
            # MONAI RandAffined default padding_mode is "reflection" (augmentations.py:98-116 sets none)
            plan.affine(rep(params["mats"][si]), modes, "reflection", where=rep(params["fired"][si]))
        plan.flip(rep(params["flips"]))
        if self.random_crop_size is not None:
            plan.center_crop(self.random_crop_size)
        return plan

    def _alloc_out(self, B, oshape, dev):
        out = {self.output_image_key: torch.empty((B, len(self.image_keys), *oshape), dtype=torch.float32, device=dev)}
        if self.has_label:
            out["mask"] = torch.empty((B, 1, *oshape), dtype=torch.float32, device=dev)
        return out

    def _dst(self, out, B):
        """Destination pointers / strides of every (sample, key) volume inside the collated batch."""
        key = (id(out), B) + tuple(t.data_ptr() for t in out.values())
        hit = self._dst_cache.get(key)
        if hit is None:
            if len(self._dst_cache) > 64:
                self._dst_cache.clear()
            hit = self._dst_uncached(out, B)
            self._dst_cache[key] = hit
        return hit

    def _dst_uncached(self, out, B):
        ni = len(self.image_keys)
        img = out[self.output_image_key]
        bi = np.arange(B, dtype=np.int64)[:, None]
        ci = np.arange(ni, dtype=np.int64)[None, :]
        p_img = img.data_ptr() + 4 * (bi * img.stride(0) + ci * img.stride(1))
        s_img = np.broadcast_to(np.asarray(img.stride()[2:], np.int64), (B, ni, 3))
        if self.has_label:
            m = out["mask"]
            p_m = m.data_ptr() + 4 * (bi * m.stride(0))
            dst_ptr = np.concatenate([p_img, p_m], axis=1).reshape(-1)
            s_m = np.broadcast_to(np.asarray(m.stride()[2:], np.int64), (B, 1, 3))
            dst_stride = np.concatenate([s_img, s_m], axis=1).reshape(-1, 3)
        else:
            dst_ptr, dst_stride = p_img.reshape(-1), s_img.reshape(-1, 3)
        return dst_ptr.astype(np.uint64), dst_stride

    def __call__(self, samples: Sequence[dict], params=None, out: dict | None = None) -> dict:
        plan = self.plan(samples, params)
        B = len(samples)
        if out is None:
            out = self._alloc_out(B, tuple(int(x) for x in plan.shape[0]), plan.device)
        dst_ptr, dst_stride = self._dst(out, B)
        engine.execute_ptrs(plan, dst_ptr, dst_stride, keep=list(out.values()))
        return out

    def prepare_steps(self, batches: Sequence[Sequence[dict]], outs: Sequence[dict]) -> "engine.PreparedSteps":
        """Draw and compose several consecutive steps at once (same RandomState order as calling
        the augmenter step by step); ``outs[k]`` receives step ``k`` when ``prepared.run(k)`` is
        called.  Amortises host composition over the steps."""
        samples = [s for b in batches for s in b]
        plan = self.plan(samples)
        ptrs, strides = [], []
        for b, out in zip(batches, outs):
            p, st = self._dst(out, len(b))
            ptrs.append(p); strides.append(st)
        nk = len(self.keys)
        return engine.prepare_steps(plan, np.concatenate(ptrs), np.concatenate(strides), [len(b) * nk for b in batches],
                                    keep=[t for o in outs for t in o.values()])
