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

from . import _lib, engine, geometry
from .plan import BatchPlan
from .sampling import RandAffineSampler, child_seeds, draw_used_uniforms

CHAIN_DTYPE = engine.CHAIN_DTYPE
_FLIP_BITS = np.array([1, 2, 4], np.uint8)


def _chain_template(ptr, stride, dtype, shape, dst_ptr, dst_stride, modes, padding: str, strict: bool) -> np.ndarray:
    """Static part of the ``adell_chain`` descriptors of a batch (one per volume): where the volumes live, where
    they go, how they are interpolated.  The per-step draws (matrix, flips, crop window) are written into a copy."""
    n = ptr.shape[0]
    ch = np.zeros(n, CHAIN_DTYPE)
    ch["src"], ch["src_stride"], ch["src_dtype"], ch["src_shape"] = ptr, stride, dtype, shape
    ch["dst"], ch["dst_stride"] = dst_ptr, dst_stride
    ch["interp"] = [_lib.INTERP_MODES[m] for m in modes]
    ch["padding"] = _lib.PADDING_MODES[padding]
    ch["pre_scale"] = ch["post_scale"] = 1.0
    ch["flags"] = _lib.CHAIN_STRICT if strict else 0
    return ch

UNET_AUGMENTS = ["intensity", "noise", "rbf", "affine", "shear", "flip", "blur", "distort", "lowres", "trivial"]
_GPU_AUGMENTS = {"affine", "shear", "flip"}


def _check_augment(augment, valid):
    for a in augment:
        if a not in valid:
            raise NotImplementedError("augment can only contain {}".format(valid))
    out_of_scope = [a for a in augment if a not in _GPU_AUGMENTS and a != "trivial"]
    if out_of_scope:
        raise NotImplementedError(
            f"augmentations {out_of_scope} are not on the batch fast path ('intensity' is served by the "
            "dictionary-transform surface; the others are out of scope, see DESIGN.md)"
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
        n_crops: int = 1,
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
        # has_label + random_crop_size: the label-guided sandwich (RandCropByPosNegLabeld(1.1x, num_samples=n_crops) ->
        # augments per crop -> CenterSpatialCropd, augmentations.py:142-176): the samples must carry the device-resident
        # index lists of FgBgToIndicesd (``mask_fg_indices`` / ``mask_bg_indices``); the crop centres are selected on
        # the device (adell_posneg_starts) and K1 reads the window starts from device memory.  Output batch = B * n_crops.
        self.posneg = self.random_crop_size is not None and has_label
        self.n_crops = int(n_crops) if self.posneg else 1
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
        self._tmpl_cache = {}
        self.set_random_state(None)

    def set_random_state(self, seed=None, nested: bool = False):
        """Seed fan-out of ``Compose.set_random_state`` over the Randomizable children †.  ``seed`` is the seed of
        the augmentation ``Compose`` (= ``get_augmentations_unet(...).set_random_state(seed)``).  The reference
        nests that Compose inside the pipeline's outer ``Compose`` and seeds the OUTER one
        (``transforms_train.set_random_state(args.seed)``, entrypoints/segmentation/train.py:449), which hands its
        only Randomizable child the first draw of ``RandomState(args.seed)``: ``nested=True`` reproduces that."""
        if nested and seed is not None:
            seed = child_seeds(seed, 1)[0]
        n = len(self.samplers) + len(self.flip_axis)
        if self.random_crop_size is not None:
            # crop sandwich (augmentations.py:142-176): Compose([RandSpatialCropd, Compose(augments + flips),
            # CenterSpatialCropd]) — the outer Compose seeds the crop and the INNER Compose, which fans out again
            outer = child_seeds(seed, 2) if seed is not None else [None, None]
            self.crop_R = np.random.RandomState(outer[0])
            seed = outer[1]
        seeds = child_seeds(seed, n) if seed is not None else [None] * n
        i = 0
        for s in self.samplers:
            s.set_random_state(seeds[i]); i += 1
        for j in range(len(self.flip_axis)):
            self.flip_R[j] = np.random.RandomState(seeds[i]); i += 1
        return self

    def draw_picks(self, samples: Sequence[dict], shape):
        """The two host draws of every crop of ``RandCropByPosNegLabeld`` (foreground or background list, entry), sample
        by sample from the crop transform's stream: ``[(index list tensor, entry), ...]`` in ``[sample, crop]`` order."""
        picks = []
        for s in samples:
            fg, bg = s[f"{self.mask_key}_fg_indices"], s[f"{self.mask_key}_bg_indices"]
            pos_ratio = 0.5
            if len(fg) == 0 or len(bg) == 0:
                if len(fg) == 0 and len(bg) == 0:
                    raise ValueError("No sampling location available.")
                pos_ratio = 0 if len(fg) == 0 else 1
            for _ in range(self.n_crops):
                lst = fg if self.crop_R.rand() < pos_ratio else bg
                picks.append((lst, int(self.crop_R.randint(len(lst)))))
        return picks

    def draw(self, batch: int, shape):
        """All host-side random parameters of one batch, sample by sample in stream order (``batch`` counts CROPS for the
        label-guided sandwich: every crop of a sample runs the inner chain on its own)."""
        nk = len(self.keys)
        fired = np.zeros((len(self.samplers), batch), bool)
        mats = np.zeros((len(self.samplers), batch, 4, 4), np.float32)
        mats[..., 0, 0] = mats[..., 1, 1] = mats[..., 2, 2] = mats[..., 3, 3] = 1.0
        flips = np.zeros((batch, 3), bool)
        starts = None
        if self.random_crop_size is not None and not self.posneg:
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

    def plan(self, samples: Sequence[dict], params=None, pre_dev: torch.Tensor | None = None) -> BatchPlan:
        """``pre_dev``: optional ``[B * len(keys), 2]`` fp32 device tensor of per-volume ``{scale, offset}`` (sample-major,
        key order = ``self.keys``) computed by the statistics kernels — the intensity normalisation of raw cached
        volumes (transforms.py:143-155 of the reference) folded into the gather instead of a pass of its own."""
        B, nk = len(samples), len(self.keys)
        metas = [self._sample_meta(s) for s in samples]
        shape = tuple(int(x) for x in metas[0][4][0])
        nc = self.n_crops
        if params is None:
            picks = self.draw_picks(samples, shape) if self.posneg else None
            params = self.draw(B * nc, shape)
            params["picks"] = picks
        # volume order [sample, crop, key]: every crop of a sample reads the sample's volumes
        cat = lambda i: np.concatenate([np.tile(m[i], (nc,) + (1,) * (m[i].ndim - 1)) for m in metas])
        plan = BatchPlan.from_arrays(cat(1), cat(2), cat(3), cat(4), metas[0][5][0].device, [m[5] for m in metas],
                                     fast=self.fast, strict=self.strict)
        if pre_dev is not None:
            plan.intensity_from_device(pre_dev)
        rep = lambda x: np.repeat(x, nk, axis=0)
        B = B * nc                              # from here on one "sample" per crop
        modes = self.modes * B
        if self.posneg:
            from .transforms import _posneg_starts

            pre = [min(int(i * 1.10), s) for i, s in zip(self.random_crop_size, shape)]
            win = params.get("win")
            if win is None:
                win = params["win"] = _posneg_starts(params["picks"], shape, pre, plan.device)   # [B, 3] int32 crop starts, on the device
            plan.keep.append(win)
            plan.crop_from_device(np.repeat(win.data_ptr() + 12 * np.arange(B, dtype=np.uint64), nk), pre)
        elif self.random_crop_size is not None:
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

    def __call__(self, samples: Sequence[dict], params=None, out: dict | None = None, pre_dev: torch.Tensor | None = None) -> dict:
        B = len(samples) * self.n_crops
        if not self.fast:
            if out is None:
                meta = self._sample_meta(samples[0])
                shape = tuple(int(x) for x in meta[4][0])
                oshape = shape if self.random_crop_size is None else tuple(min(r, s) for r, s in zip(self.random_crop_size, shape))
                out = self._alloc_out(B, oshape, meta[5][0].device)
            ch, params = self.chains([samples], [out], params, pre_dev)
            if ch is not None:
                dev = self._sample_meta(samples[0])[5][0].device
                engine.prepare_chain_steps(ch, [ch.shape[0]], dev, keep=list(out.values()) + [samples, pre_dev, params.get("win")]).run(0)
                return out
        plan = self.plan(samples, params, pre_dev)
        if out is None:
            out = self._alloc_out(B, tuple(int(x) for x in plan.shape[0]), plan.device)
        dst_ptr, dst_stride = self._dst(out, B)
        engine.execute_ptrs(plan, dst_ptr, dst_stride, keep=list(out.values()))
        return out

    # ---- native route: chains composed by adell_chain_prepare_steps (no numpy work per volume) ----
    def _template(self, batch: Sequence[dict], out: dict) -> np.ndarray:
        key = (id(batch[0]), id(batch[-1]), len(batch), id(out))
        hit = self._tmpl_cache.get(key)
        if hit is not None:   # ids can be recycled: the cached entry must hold these very objects, at these addresses
            ok = hit[2] is out and hit[3] == tuple(t.data_ptr() for t in out.values())
            for a, b in zip(hit[1], batch):
                ok = ok and a is b
            if not ok:
                hit = None
        if hit is None:
            metas = [self._sample_meta(s) for s in batch]
            nc = self.n_crops                     # volume order [sample, crop, key]: every crop reads the sample's volumes
            dst_ptr, dst_stride = self._dst(out, len(batch) * nc)
            cat = lambda i: np.concatenate([np.tile(m[i], (nc,) + (1,) * (m[i].ndim - 1)) for m in metas])
            ch = _chain_template(cat(1), cat(2), cat(3), cat(4), dst_ptr, dst_stride, self.modes * (len(batch) * nc), "reflection", self.strict)
            if self.random_crop_size is not None:
                shape = tuple(int(x) for x in metas[0][4][0])
                ch["crop0_size"] = [min(int(i * 1.10), s) for i, s in zip(self.random_crop_size, shape)]
                ch["crop1_size"] = self.random_crop_size
            if len(self._tmpl_cache) > 256:
                self._tmpl_cache.clear()
            hit = self._tmpl_cache[key] = (ch, list(batch), out, tuple(t.data_ptr() for t in out.values()))
        return hit[0]

    def chains(self, batches: Sequence[Sequence[dict]], outs: Sequence[dict], params=None, pre_dev: torch.Tensor | None = None):
        """``adell_chain`` descriptors of several consecutive steps (same draws, in the same RandomState order, as
        :meth:`plan`), or ``None`` when some sample needs more than one resample (both RandAffined fired: the
        reference resamples twice, which takes the multi-pass route of :class:`BatchPlan`)."""
        nk = len(self.keys)
        n_samples = sum(len(b) for b in batches) * self.n_crops      # (one "sample" per crop from here on)
        shape = tuple(int(x) for x in self._sample_meta(batches[0][0])[4][0])
        if params is None:
            picks = self.draw_picks([s for b in batches for s in b], shape) if self.posneg else None
            params = self.draw(n_samples, shape)
            params["picks"] = picks
        fired = params["fired"]
        if fired.shape[0] > 1 and (fired.sum(axis=0) > 1).any():
            return None, params
        # (templates are concatenated as raw bytes: numpy's concatenate of structured arrays re-derives the dtype per call)
        raw = [self._template(b, o).view(np.uint8) for b, o in zip(batches, outs)]
        ch = (np.concatenate(raw) if len(raw) > 1 else raw[0].copy()).view(CHAIN_DTYPE)
        if fired.shape[0]:
            any_fired = fired.any(axis=0)
            if any_fired.any():
                which = fired.argmax(axis=0)
                A = params["mats"][which, np.arange(n_samples), :3].reshape(n_samples, 12)
                ch["A"] = np.repeat(A, nk, axis=0)
                ch["flags"] |= np.repeat(np.where(any_fired, _lib.CHAIN_AFFINE, 0).astype(np.uint8), nk)
        ch["flip1"] = np.repeat(params["flips"].astype(np.uint8) @ _FLIP_BITS, nk)
        if params["starts"] is not None:
            ch["crop0_start"] = np.repeat(params["starts"], nk, axis=0)
        if self.posneg:
            # label-guided crops: ONE select launch for the crops of all these steps; K1 reads each window's start from
            # device memory when its step runs
            from .transforms import _posneg_starts

            dev = self._sample_meta(batches[0][0])[5][0].device
            pre = [min(int(i * 1.10), s) for i, s in zip(self.random_crop_size, shape)]
            win = params.get("win")
            if win is None:
                win = params["win"] = _posneg_starts(params["picks"], shape, pre, dev)       # [crops, 3] int32 on the device
            ch["win_dev"] = np.repeat(win.data_ptr() + 12 * np.arange(n_samples, dtype=np.uint64), nk)
        if pre_dev is not None:
            if pre_dev.shape != (ch.shape[0], 2) or pre_dev.dtype != torch.float32 or not pre_dev.is_contiguous():
                raise ValueError("pre_dev must be a contiguous [n, 2] float32 tensor")
            ch["pre_dev"] = pre_dev.data_ptr() + 8 * np.arange(ch.shape[0], dtype=np.uint64)
        return ch, params

    def prepare_steps(self, batches: Sequence[Sequence[dict]], outs: Sequence[dict], pre_dev: torch.Tensor | None = None) -> "engine.PreparedSteps":
        """Draw and compose several consecutive steps at once (same RandomState order as calling
        the augmenter step by step); ``outs[k]`` receives step ``k`` when ``prepared.run(k)`` is
        called.  Amortises host composition over the steps.  ``pre_dev``: ``[volumes of all steps, 2]`` device
        ``{scale, offset}`` rows the kernel reads when a step RUNS — the statistics kernels of step ``k`` may fill its
        rows right before ``run(k)``."""
        samples = [s for b in batches for s in b]
        nk = len(self.keys)
        ch, params = (None, None) if self.fast else self.chains(batches, outs, pre_dev=pre_dev)
        if ch is not None:
            dev = self._sample_meta(samples[0])[5][0].device
            return engine.prepare_chain_steps(ch, [len(b) * self.n_crops * nk for b in batches], dev,
                                              keep=[t for o in outs for t in o.values()] + [batches, pre_dev, params.get("win")])
        if pre_dev is not None:
            raise NotImplementedError("device-side intensity rows need the single-resample route (some sample fired both RandAffined)")
        plan = self.plan(samples, params)
        ptrs, strides = [], []
        for b, out in zip(batches, outs):
            p, st = self._dst(out, len(b) * self.n_crops)
            ptrs.append(p); strides.append(st)
        nk = len(self.keys)
        return engine.prepare_steps(plan, np.concatenate(ptrs), np.concatenate(strides), [len(b) * self.n_crops * nk for b in batches],
                                    keep=[t for o in outs for t in o.values()])


class _BatchBase:
    """Shared plumbing of the batch augmenters: per-sample metadata cache, destination pointers."""

    def __init__(self):
        self._meta = {}
        self._tmpl_cache = {}

    def _sample_meta(self, s: dict, keys):
        m = self._meta.get(id(s))
        if m is None or m[0] is not s:
            vols = [s[k][c] for k in keys for c in range(s[k].shape[0])]
            plan = BatchPlan(vols)
            m = (s, plan.parent_ptr, plan.parent_stride, plan.parent_dtype, plan.shape, vols)
            self._meta[id(s)] = m
        return m

    def _base_plan(self, samples, keys, repeat: int = 1):
        """One volume per (sample, [view,] key-channel); ``repeat`` > 1 lists each sample's volumes
        that many times (the SSL views share their source)."""
        metas = [self._sample_meta(s, keys) for s in samples]
        key = ("base", id(samples[0]), id(samples[-1]), len(samples), repeat)
        hit = self._tmpl_cache.get(key)
        if hit is not None:
            for a, b in zip(hit[1], samples):
                if a is not b:
                    hit = None
                    break
        if hit is None:
            cat = lambda i: np.concatenate([np.tile(m[i], (repeat,) + (1,) * (m[i].ndim - 1)) for m in metas])
            if len(self._tmpl_cache) > 256:
                self._tmpl_cache.clear()
            hit = self._tmpl_cache[key] = ((cat(1), cat(2), cat(3), cat(4)), list(samples))
        a1, a2, a3, a4 = hit[0]
        return BatchPlan.from_arrays(a1, a2, a3, a4, metas[0][5][0].device, [m[5] for m in metas],
                                     fast=self.fast, strict=self.strict), metas

    @staticmethod
    def _dst_of(out: torch.Tensor):
        """Pointers / strides of every [b, c] volume of a collated ``[B, C, H, W, D]`` tensor."""
        B, Cn = out.shape[:2]
        bi = np.arange(B, dtype=np.int64)[:, None]
        ci = np.arange(Cn, dtype=np.int64)[None, :]
        ptr = (out.data_ptr() + 4 * (bi * out.stride(0) + ci * out.stride(1))).astype(np.uint64)
        stride = np.broadcast_to(np.asarray(out.stride()[2:], np.int64), (B, Cn, 3))
        return ptr, stride


class ClassificationBatchAugmenter(_BatchBase):
    """``get_augmentations_class`` chain (+ the final ``CenterSpatialCropd`` / ``ConcatItemsd`` of
    ``ClassificationTransforms.post_transforms``) for a batch of cached samples
    (/root/reference/adell_mri/transform_factory/augmentations.py:181-320,
    /root/reference/adell_mri/transform_factory/transforms.py:492-509): ``OneOf`` over the flip
    combinations FIRST (a flip before the resample is a negative source stride), then
    ``RandAffined(translate, rotate x, scale; zeros)``, optionally the shear one, centre crop, and
    every key (mask last) concatenated into ``"image"``.  ``set_random_state(seed)`` seeds exactly
    like ``get_augmentations_class(...).set_random_state(seed)``."""

    def __init__(self, augment: Sequence[str], image_keys: Sequence[str], mask_key: str | None = None,
                 flip_axis: Sequence[int] = (0, 1), prob: float = 0.1, crop_size: Sequence[int] | None = None,
                 strict: bool = False, fast: bool = False):
        super().__init__()
        valid = ["intensity", "noise", "rbf", "affine", "shear", "flip", "blur", "lowres", "distort", "trivial"]
        _check_augment(augment, valid)
        if "trivial" in augment:
            raise NotImplementedError("'trivial' (SomeOf) is handled by the dict-transform surface, not the batch fast path")
        self.keys = list(image_keys) + ([mask_key] if mask_key is not None else [])
        self.modes = ["bilinear" if k != mask_key else "nearest" for k in self.keys]
        self.prob, self.crop_size = prob, None if crop_size is None else [int(x) for x in crop_size]
        self.strict, self.fast = strict, fast
        flip_axis = [flip_axis] if isinstance(flip_axis, int) else list(flip_axis)
        self.flip_combos = [c for i in range(len(flip_axis)) for c in itertools.combinations(flip_axis, i + 1)] if "flip" in augment else []
        self.oneof_R = np.random.RandomState()
        self.flip_R = [np.random.RandomState() for _ in self.flip_combos]
        self.samplers = []
        if "affine" in augment:
            self.samplers.append(RandAffineSampler(prob=prob, translate_range=[4, 4, 1], rotate_range=[np.pi / 16],
                                                   scale_range=[0.1, 0.1, 0.05]))
        if "shear" in augment:
            self.samplers.append(RandAffineSampler(prob=prob, shear_range=((0.9, 1.1), (0.9, 1.1), (0.9, 1.1))))
        self.set_random_state(None)

    def set_random_state(self, seed=None, nested: bool = False):
        """``nested=True``: ``seed`` is the seed of the pipeline's OUTER Compose (the reference's
        ``transforms_train.set_random_state(args.seed)``, entrypoints/classification/train.py:243), whose only
        Randomizable child is this augmentation Compose."""
        if nested and seed is not None:
            seed = child_seeds(seed, 1)[0]
        n = (1 if self.flip_combos else 0) + len(self.samplers)
        seeds = child_seeds(seed, n) if seed is not None else [None] * n
        i = 0
        if self.flip_combos:
            self.oneof_R = np.random.RandomState(seeds[i]); i += 1
            fs = [int(self.oneof_R.randint(2 ** 32, dtype="uint32")) for _ in self.flip_combos] if seed is not None else [None] * len(self.flip_combos)
            self.flip_R = [np.random.RandomState(f) for f in fs]
        for smp in self.samplers:
            smp.set_random_state(seeds[i]); i += 1
        return self

    def draw(self, batch: int):
        nk = len(self.keys)
        flips = np.zeros((batch, 3), bool)
        if self.flip_combos:
            w = [1.0 / len(self.flip_combos)] * len(self.flip_combos)
            idx = self.oneof_R.multinomial(1, w, size=batch).argmax(1)      # == one OneOf.__call__ per sample
            for j, combo in enumerate(self.flip_combos):
                use = np.nonzero(idx == j)[0]
                fire = self.flip_R[j].random_sample(use.size) < self.prob   # that member's own stream, in sample order
                for a in combo:
                    flips[use[fire], a] = True
        fired = np.zeros((len(self.samplers), batch), bool)
        mats = np.tile(np.eye(4, dtype=np.float32), (len(self.samplers), batch, 1, 1))
        for si, smp in enumerate(self.samplers):
            f, p = smp.draw_batch(batch, n_keys=nk)
            fired[si] = f
            if f.any():
                mats[si, f] = geometry.compose_affine(p["rotate"], p["shear"], p["translate"], p["scale"], batch=int(f.sum()))
        return dict(flips=flips, fired=fired, mats=mats)

    def plan(self, samples: Sequence[dict], params=None, pre_dev: torch.Tensor | None = None) -> BatchPlan:
        """``pre_dev``: optional per-volume ``{scale, offset}`` on the device (see SegmentationBatchAugmenter.plan)."""
        B = len(samples)
        if params is None:
            params = self.draw(B)
        plan, metas = self._base_plan(samples, self.keys)
        if pre_dev is not None:
            plan.intensity_from_device(pre_dev)
        per = plan.n // B                      # volumes per sample (keys x channels)
        modes = []
        for k, m in zip(self.keys, self.modes):
            modes += [m] * samples[0][k].shape[0]
        rep = lambda x: np.repeat(x, per, axis=0)
        plan.flip(rep(params["flips"]))
        for si in range(len(self.samplers)):
            plan.affine(rep(params["mats"][si]), modes * B, "zeros", where=rep(params["fired"][si]))
        if self.crop_size is not None:
            plan.center_crop(self.crop_size)
        return plan

    def chains(self, samples: Sequence[dict], out: dict, params=None, pre_dev: torch.Tensor | None = None):
        """``adell_chain`` descriptors of one batch for the native composer (``None`` when a sample fired both
        RandAffined: two resamples, the multi-pass route)."""
        B = len(samples)
        if params is None:
            params = self.draw(B)
        fired = params["fired"]
        if fired.shape[0] > 1 and (fired.sum(axis=0) > 1).any():
            return None, params
        key = (tuple(id(s) for s in samples), id(out), out["image"].data_ptr())
        hit = self._tmpl_cache.get(key)
        if hit is None or any(a is not b for a, b in zip(hit[1], samples)):
            metas = [self._sample_meta(s, self.keys) for s in samples]
            cat = lambda i: np.concatenate([m[i] for m in metas])
            modes = []
            for k, m in zip(self.keys, self.modes):
                modes += [m] * samples[0][k].shape[0]
            ptr, stride = self._dst_of(out["image"])
            ch = _chain_template(cat(1), cat(2), cat(3), cat(4), ptr.reshape(-1), stride.reshape(-1, 3), modes * B, "zeros", self.strict)
            if self.crop_size is not None:
                ch["crop1_size"] = self.crop_size
            if len(self._tmpl_cache) > 256:
                self._tmpl_cache.clear()
            hit = self._tmpl_cache[key] = (ch, list(samples))
        ch = hit[0].copy()
        per = ch.shape[0] // B
        if fired.shape[0]:
            any_fired = fired.any(axis=0)
            if any_fired.any():
                which = fired.argmax(axis=0)
                ch["A"] = np.repeat(params["mats"][which, np.arange(B), :3].reshape(B, 12), per, axis=0)
                ch["flags"] |= np.repeat(np.where(any_fired, _lib.CHAIN_AFFINE, 0).astype(np.uint8), per)
        ch["flip0"] = np.repeat(params["flips"].astype(np.uint8) @ _FLIP_BITS, per)
        if pre_dev is not None:
            if pre_dev.shape != (ch.shape[0], 2) or pre_dev.dtype != torch.float32 or not pre_dev.is_contiguous():
                raise ValueError("pre_dev must be a contiguous [n, 2] float32 tensor")
            ch["pre_dev"] = pre_dev.data_ptr() + 8 * np.arange(ch.shape[0], dtype=np.uint64)
        return ch, params

    def prepare_steps(self, batches: Sequence[Sequence[dict]], outs: Sequence[dict], pre_dev: torch.Tensor | None = None):
        """Draw and compose several consecutive steps at once (one draw over all their samples: every stream is consumed
        in sample order, so this equals step-by-step draws); ``outs[k]`` receives step ``k`` when ``prepared.run(k)`` is
        called.  ``pre_dev``: ``[volumes of all steps, 2]`` device ``{scale, offset}`` rows read when a step RUNS — the
        statistics kernels of step ``k`` may fill its rows right before ``run(k)``.  Returns ``None`` when some sample
        fired both RandAffined (two resamples: call the augmenter step by step)."""
        if self.fast:
            return None
        total = sum(len(b) for b in batches)
        params = self.draw(total)
        chs, sizes, s0, v0 = [], [], 0, 0
        for b, o in zip(batches, outs):
            B = len(b)
            sub = dict(flips=params["flips"][s0:s0 + B], fired=params["fired"][:, s0:s0 + B], mats=params["mats"][:, s0:s0 + B])
            per = sum(b[0][k].shape[0] for k in self.keys)
            ch, _ = self.chains(b, o, sub, None if pre_dev is None else pre_dev[v0:v0 + B * per])
            if ch is None:
                return None
            chs.append(ch.view(np.uint8))
            sizes.append(ch.shape[0])
            s0 += B
            v0 += B * per
        ch = np.concatenate(chs).view(CHAIN_DTYPE)
        return engine.prepare_chain_steps(ch, sizes, outs[0]["image"].device, keep=[o["image"] for o in outs] + [batches, pre_dev])

    def __call__(self, samples: Sequence[dict], params=None, out: dict | None = None, pre_dev: torch.Tensor | None = None) -> dict:
        B = len(samples)
        if not self.fast:
            if out is None:
                meta = self._sample_meta(samples[0], self.keys)
                shape = tuple(int(x) for x in meta[4][0])
                oshape = shape if self.crop_size is None else tuple(min(r, s) for r, s in zip(self.crop_size, shape))
                n_ch = sum(samples[0][k].shape[0] for k in self.keys)
                out = {"image": torch.empty((B, n_ch, *oshape), dtype=torch.float32, device=meta[5][0].device)}
            ch, params = self.chains(samples, out, params, pre_dev)
            if ch is not None:
                engine.prepare_chain_steps(ch, [ch.shape[0]], out["image"].device, keep=[out["image"], samples, pre_dev]).run(0)
                return out
        plan = self.plan(samples, params, pre_dev)
        per = plan.n // B
        if out is None:
            out = {"image": torch.empty((B, per, *(int(x) for x in plan.shape[0])), dtype=torch.float32, device=plan.device)}
        ptr, stride = self._dst_of(out["image"])
        engine.execute_ptrs(plan, ptr.reshape(-1), stride.reshape(-1, 3), keep=[out["image"]])
        return out


#: workhorse members the fused path implements, in the reference's list order
#: (/root/reference/adell_mri/modules/augmentations.py:10-37)
SSL_FUSED_MEMBERS = ["gaussian_noise", "shift_intensity", "scale_intensity",
                     "rotate_x", "rotate_y", "rotate_z", "translate_x", "translate_y", "translate_z",
                     "shear_x", "shear_y", "shear_z", "scale_x", "scale_y", "scale_z"]


def _ssl_member_ranges(name: str, max_mult: float):
    """Parameter range of one spatial member after ``max_mult`` and AUG_PARAM_CORRECTION
    (modules/augmentations.py:103-162): a per-axis tuple with the active axis a (lo, hi) pair."""
    kind, c = name.rsplit("_", 1)
    i = "xyz".index(c)
    rng = [0, 0, 0]
    if kind == "rotate":
        a = (np.pi / 6 if c != "z" else np.pi / 16) * max_mult
        rng[i] = (-a, a)
        return dict(rotate_range=tuple(rng))
    if kind == "translate":
        t = (30 if c != "z" else 5) * max_mult
        rng[i] = (-t, t)
        return dict(translate_range=tuple(rng))
    x = (0.5 if kind == "shear" else 0.3) * max_mult
    rng[i] = (1 - x, 1 + x)
    return dict(shear_range=tuple(rng)) if kind == "shear" else dict(scale_range=tuple(rng))


class _WorkhorseDraws:
    """Vectorised draws of one ``AugmentationWorkhorsed`` over a batch: which members each sample
    applies (in which order) and every member's parameters from ITS OWN streams, consumed in
    sample order — the same values the per-sample transform objects would draw."""

    def __init__(self, members, n_keys: int, N: int, max_mult: float = 0.5):
        self.members, self.n_keys, self.N = list(members), n_keys, N
        self.R = np.random.RandomState()
        self.samplers, self.R_outer, self.R_inner = {}, {}, {}
        for m in self.members:
            if m in ("gaussian_noise", "shift_intensity", "scale_intensity"):
                self.R_outer[m], self.R_inner[m] = np.random.RandomState(), np.random.RandomState()
            else:
                self.samplers[m] = RandAffineSampler(prob=1.0, **_ssl_member_ranges(m, max_mult))
        self.noise_std = 1 * max_mult
        self.shift, self.scale = 0.5 * max_mult, 0.5 * max_mult
        # per spatial member: which of the 12 parameter columns (rotate 0-2, shear 3-5, translate 6-8, scale 9-11) its
        # three uniforms fill, and the affine map u -> lo + (hi - lo) u (+ 1 for scale) of numpy's uniform()
        self._cols = {}
        for m, smp in self.samplers.items():
            base, rng, add = next((b, r, a) for b, r, a in ((0, smp.rotate_range, 0.0), (3, smp.shear_range, 0.0),
                                                            (6, smp.translate_range, 0.0), (9, smp.scale_range, 1.0)) if r is not None)
            lo = np.array([(f[0] if isinstance(f, (tuple, list)) else -f) for f in rng], np.float64)
            hi = np.array([(f[1] if isinstance(f, (tuple, list)) else f) for f in rng], np.float64)
            self._cols[m] = (base, lo, hi - lo, add)

    def set_random_state(self, seed):
        self.R = np.random.RandomState(seed)
        for m in self.members:  # AugmentationWorkhorsed.set_random_state fan-out (transform_factory.py)
            s = int(self.R.randint(2 ** 32, dtype="uint32")) if seed is not None else None
            if m in self.samplers:
                self.samplers[m].set_random_state(s)
            else:
                self.R_outer[m], self.R_inner[m] = np.random.RandomState(s), np.random.RandomState(s)
        return self

    def draw_members(self, choice: np.ndarray, shape, philox: bool):
        """``choice``: [B, N] member indices.  Returns per member the samples using it and its
        parameters (matrices / factors / offsets / noise)."""
        out = {}
        spatial = []   # (member, use, params): composed in ONE call below (absent factors = identity matrices: exact)
        for mi, m in enumerate(self.members):
            use = np.nonzero((choice == mi).any(axis=1))[0]
            if use.size == 0:
                continue
            if m in self.samplers:
                smp = self.samplers[m]
                if smp.R is not smp.R_inner:   # (separately seeded streams: one numpy call each)
                    spatial.append((m, use, draw_used_uniforms(smp, use.size, self.n_keys)))
                else:
                    _, p = smp.draw_batch(use.size, n_keys=self.n_keys)
                    spatial.append((m, use, p))
            elif m == "gaussian_noise":
                self.R_outer[m].random_sample(use.size)
                if philox:
                    # per use: gate random_sample(), uniform(0, noise_std), randint(2**32) = 2 + 2 + 1 words of the
                    # Mersenne twister, drawn raw in one call and assembled the way numpy does (53-bit doubles)
                    w = self.R_inner[m].randint(0, 2 ** 32, size=5 * use.size, dtype=np.uint32).reshape(-1, 5).astype(np.uint64)
                    u = ((w[:, 2] >> np.uint64(5)).astype(np.float64) * 67108864.0 + (w[:, 3] >> np.uint64(6)).astype(np.float64)) / 9007199254740992.0
                    std = (0.0 + (self.noise_std - 0.0) * u).astype(np.float32)
                    vals = list(zip(std, (int(x) for x in w[:, 4])))
                else:
                    vals = []
                    for _ in range(use.size):   # the noise volume itself is drawn per use (host float64 normal)
                        self.R_inner[m].random_sample()
                        std = self.R_inner[m].uniform(0, self.noise_std)
                        vals.append(self.R_inner[m].normal(0.0, std, size=shape).astype(np.float32))
                out[m] = (use, vals)
            else:
                self.R_outer[m].random_sample(use.size)
                u = self.R_inner[m].random_sample(2 * use.size).reshape(-1, 2)[:, 1]   # gate, then uniform(lo, hi)
                f = self.shift if m == "shift_intensity" else self.scale
                out[m] = (use, -f + (f - (-f)) * u)
        if spatial:
            n = sum(u.size for _, u, _ in spatial)
            par = np.zeros((n, 12))
            par[:, 9:] = 1.0
            o = 0
            for m, use, p in spatial:
                if isinstance(p, np.ndarray):        # raw uniforms of the used randomize
                    base, lo, span, add = self._cols[m]
                    par[o:o + use.size, base:base + 3] = lo + span * p + add
                else:
                    for b, name in ((0, "rotate"), (3, "shear"), (6, "translate"), (9, "scale")):
                        if p[name].shape[1]:
                            par[o:o + use.size, b:b + p[name].shape[1]] = p[name]
                o += use.size
            mats = geometry.compose_affine(par[:, 0:3], par[:, 3:6], par[:, 6:9], par[:, 9:12], batch=n)
            o = 0
            for m, use, _ in spatial:
                out[m] = (use, mats[o:o + use.size])
                o += use.size
        return out


class SSLBatchAugmenter(_BatchBase):
    """``get_augmentations_ssl`` two-view chain for a batch (augmentations.py:391-516;
    modules/augmentations.py:189-256): the crop window(s), then per view ``n_transforms`` of the
    fused workhorse members in the drawn order.  ``choice="global"`` draws the member subsets with
    ``np.random.choice`` on the global stream, sample by sample, exactly like the reference (whose
    global stream is unseeded, so nothing reproducible is lost by ``choice="vectorised"``, which
    draws all subsets at once from this object's own stream).  ``noise="philox"`` replaces the
    host-drawn noise volumes by the device generator (sigma still drawn on the host)."""

    def __init__(self, all_keys: Sequence[str], roi_size: Sequence[int], n_transforms: int = 3, different_crop: bool = False,
                 vicregl: bool = False, members: Sequence[str] | None = None, choice: str = "global", noise: str = "injected",
                 strict: bool = False, fast: bool = False):
        super().__init__()
        self.keys = list(all_keys)
        self.roi = [int(x) for x in roi_size]
        self.N, self.different_crop, self.vicregl = n_transforms, different_crop or vicregl, vicregl
        members = list(SSL_FUSED_MEMBERS if members is None else members)
        if vicregl:
            members = [m for m in members if m in ("gaussian_noise", "shift_intensity", "scale_intensity")]
        bad = [m for m in members if m not in SSL_FUSED_MEMBERS]
        if bad:
            raise NotImplementedError(f"workhorse members {bad} are outside the fused GPU hot path")
        self.members = members
        self.choice, self.noise, self.strict, self.fast = choice, noise, strict, fast
        self.crop_R = [np.random.RandomState(), np.random.RandomState()]
        self.views = None
        self.choice_R = np.random.RandomState()
        self.set_random_state(None)

    def _n_channels(self, sample):
        return sum(sample[k].shape[0] for k in self.keys)

    def set_random_state(self, seed=None):
        n = (2 if self.different_crop else 1) + 2
        seeds = child_seeds(seed, n) if seed is not None else [None] * n
        i = 0
        self.crop_R[0] = np.random.RandomState(seeds[i]); i += 1
        if self.different_crop:
            self.crop_R[1] = np.random.RandomState(seeds[i]); i += 1
        self._view_seeds = seeds[i:i + 2]
        self.views = None
        self.choice_R = np.random.RandomState(None if seed is None else seed + 1)
        return self

    def _ensure_views(self, n_keys):
        if self.views is None:
            self.views = [_WorkhorseDraws(self.members, n_keys, self.N).set_random_state(s) for s in self._view_seeds]

    def draw(self, batch: int, shape, n_keys: int):
        self._ensure_views(n_keys)
        nm = len(self.members)
        starts = np.zeros((2, batch, 3), np.int64)
        # RandSpatialCropd(random_size=False): a randint per axis that can move, sample by sample — drawn for the whole
        # batch by ONE call with an array of upper bounds (the legacy generator fills it element by element with the same
        # masked rejection as scalar calls: identical values, identical stream position; tests/test_host_logic.py)
        mov = [a for a, (d, r) in enumerate(zip(shape, self.roi)) if d > r]
        if mov and batch:
            high = np.tile(np.array([shape[a] - self.roi[a] + 1 for a in mov], np.int64), (batch, 1))
            starts[0][:, mov] = self.crop_R[0].randint(0, high)
            if self.different_crop:
                starts[1][:, mov] = self.crop_R[1].randint(0, high)
        if not self.different_crop:
            starts[1] = starts[0]
        choice = np.zeros((2, batch, self.N), np.int64)
        if self.choice == "global":
            for b in range(batch):
                for v in range(2):   # per sample: view 1's workhorse, then view 2's
                    choice[v, b] = np.random.choice(nm, self.N, replace=False)
        else:
            for v in range(2):
                choice[v] = np.argsort(self.choice_R.random_sample((batch, nm)), axis=1)[:, : self.N]
        roi_shape = tuple(min(r, d) for r, d in zip(self.roi, shape))
        draws = [self.views[v].draw_members(choice[v], (n_keys, *roi_shape), self.noise == "philox") for v in range(2)]
        return dict(starts=starts, choice=choice, draws=draws)

    def _slots(self, params, B: int, nc: int, vox: int, dev=None):
        """Slot ``s`` of every (sample, view) of a drawn batch, as batch-wide records: each volume applies exactly ONE
        member per slot, so the members of a slot touch disjoint volumes and collapse into at most one affine / one
        intensity / one noise record for the whole batch (the reference runs them one transform call per sample; the
        order between different volumes is immaterial).  Yields ``dict(A, has_aff, sc, of, has_int, std, seed, off_p,
        has_phx, noise)`` over the ``n = B * 2 * nc`` volumes (order ``[b, view, channel]``)."""
        n = B * 2 * nc
        ch = np.arange(nc)[None, :]
        eye = np.eye(4, dtype=np.float32)
        # Per view: where every member's per-use values live, so that a slot of the whole batch is gathered with a few
        # fancy-index operations instead of a Python loop over the members.
        nm = len(self.members)
        kind = np.array([0 if m in self.views[0].samplers else (1 if m == "scale_intensity" else (2 if m == "shift_intensity" else 3))
                         for m in self.members])
        tables = []
        for v in range(2):
            choice, draws = params["choice"][v], params["draws"][v]
            used = np.zeros((nm, B), bool)
            used[choice.reshape(-1), np.repeat(np.arange(B), self.N)] = True
            pos = np.cumsum(used, axis=1) - 1                       # index of sample b in member mi's use list
            off = np.zeros(nm, np.int64)                            # start of member mi inside the concatenated values of its kind
            cat = {0: [], 1: [], 2: [], 3: []}
            cnt = {0: 0, 1: 0, 2: 0, 3: 0}
            for mi, m in enumerate(self.members):
                if m not in draws:
                    continue
                k = int(kind[mi])
                off[mi] = cnt[k]
                vals = draws[m][1]
                cat[k].append(vals)
                cnt[k] += len(vals)
            mats = np.concatenate(cat[0]) if cat[0] else np.zeros((0, 4, 4), np.float32)
            scv = np.concatenate(cat[1]) if cat[1] else np.zeros(0)
            shv = np.concatenate(cat[2]) if cat[2] else np.zeros(0)
            noise_vals = [x for lst in cat[3] for x in lst]          # philox: (std, seed) pairs; injected: arrays
            if self.noise == "philox" and noise_vals:
                noise_vals = (np.array([q[0] for q in noise_vals], np.float32), np.array([q[1] for q in noise_vals], np.uint64))
            tables.append((pos, off, mats, scv, shv, noise_vals))
        barange = np.arange(B)
        for s in range(self.N):
            A = np.empty((n, 4, 4), np.float32)
            A[:] = eye
            has_aff = np.zeros(n, bool)
            sc, of, has_int = np.ones(n), np.zeros(n), np.zeros(n, bool)
            std, seed, off_p = np.zeros(n, np.float32), np.zeros(n, np.uint64), np.zeros(n, np.uint64)
            has_phx = np.zeros(n, bool)
            noise = None
            for v in range(2):
                pos, off, mats, scv, shv, noise_vals = tables[v]
                mi = params["choice"][v][:, s]                       # member of every sample in this slot
                k = kind[mi]
                where = off[mi] + pos[mi, barange]                   # row of the sample's value among its kind's values
                vol = (barange[:, None] * 2 + v) * nc + ch           # [B, nc] volume indices (order [b, view, channel])
                sel = k == 0
                if sel.any():
                    A[vol[sel]] = mats[where[sel]][:, None]
                    has_aff[vol[sel]] = True
                sel = k == 1
                if sel.any():
                    sc[vol[sel]] = (1 + scv[where[sel]]).astype(np.float32)[:, None]
                    has_int[vol[sel]] = True
                sel = k == 2
                if sel.any():
                    of[vol[sel]] = shv[where[sel]].astype(np.float32)[:, None]
                    has_int[vol[sel]] = True
                sel = np.nonzero(k == 3)[0]
                if sel.size:
                    if self.noise == "philox":
                        std[vol[sel]] = noise_vals[0][where[sel]][:, None]
                        seed[vol[sel]] = noise_vals[1][where[sel]][:, None]
                        off_p[vol[sel]] = (np.arange(nc, dtype=np.uint64) * np.uint64(vox))[None, :]
                        has_phx[vol[sel]] = True
                    else:
                        if noise is None:
                            noise = [None] * n
                        for b, q in zip(sel, where[sel]):
                            t = torch.from_numpy(noise_vals[q])
                            t = t.pin_memory().to(dev, non_blocking=True) if dev is not None and dev.type == "cuda" else t
                            for c in range(nc):
                                noise[int(vol[b, c])] = t[c]
            yield dict(A=A, has_aff=has_aff, sc=sc, of=of, has_int=has_int, std=std, seed=seed, off_p=off_p, has_phx=has_phx, noise=noise)

    def plan(self, samples: Sequence[dict], params=None):
        B = len(samples)
        nc = self._n_channels(samples[0])
        plan, metas = self._base_plan(samples, self.keys, repeat=2)      # volume order: [b, view, channel]
        shape = tuple(int(x) for x in metas[0][4][0])
        if params is None:
            params = self.draw(B, shape, nc)
        vol_b = np.repeat(np.arange(B), 2 * nc)
        vol_v = np.tile(np.repeat(np.arange(2), nc), B)
        plan.crop(params["starts"][vol_v, vol_b], self.roi)
        vox = int(np.prod(plan.shape[0]))
        for r in self._slots(params, B, nc, vox, plan.device):
            if r["has_aff"].any():
                plan.affine(r["A"], "bilinear", "zeros", where=r["has_aff"])
            if r["has_int"].any():
                plan.intensity(scale=r["sc"], offset=r["of"], where=r["has_int"])
            if r["has_phx"].any():
                plan.add_philox_noise(r["std"], r["seed"], r["off_p"], where=r["has_phx"])
            if r["noise"] is not None:
                plan.add_noise(r["noise"])
        return plan, params

    # ------------------------------------------------------------------ native route (adell_seq)
    def native_ok(self) -> bool:
        """The native sequence composer covers every member except host-drawn (injected) noise volumes."""
        return (self.noise == "philox" or "gaussian_noise" not in self.members) and self.N <= _lib.SEQ_MAX_OPS

    def seqs(self, batches: Sequence[Sequence[dict]], outs: Sequence[dict], params=None) -> np.ndarray:
        """``adell_seq`` descriptors of several consecutive steps (``batches[k]`` written to ``outs[k]``): ONE draw over
        all their samples (the streams are consumed in sample order, so this equals step-by-step draws) and a handful
        of numpy assignments per slot; composition itself happens in ``adell_seq_prepare_steps``."""
        samples = [s for b in batches for s in b]
        Bt = len(samples)
        nc = self._n_channels(samples[0])
        metas = [self._sample_meta(s, self.keys) for s in samples]
        shape = tuple(int(x) for x in metas[0][4][0])
        if params is None:
            params = self.draw(Bt, shape, nc)
        key = ("seq", id(samples[0]), id(samples[-1]), Bt, tuple(id(o["augmented_image_1"]) for o in outs))
        hit = self._tmpl_cache.get(key)
        if hit is not None and any(a is not b for a, b in zip(hit[1], samples)):
            hit = None
        if hit is None:
            n = Bt * 2 * nc
            sq = np.zeros(n, engine.SEQ_DTYPE)
            rep = lambda i: np.concatenate([np.tile(m[i], (2,) + (1,) * (m[i].ndim - 1)) for m in metas])
            sq["src"], sq["src_stride"], sq["src_dtype"], sq["src_shape"] = rep(1), rep(2), rep(3), rep(4)
            ptrs, strides = [], []
            for o in outs:
                p1, s1 = self._dst_of(o["augmented_image_1"])
                p2, s2 = self._dst_of(o["augmented_image_2"])
                ptrs.append(np.stack([p1, p2], axis=1).reshape(-1))          # [b, view, channel]
                strides.append(np.stack([s1, s2], axis=1).reshape(-1, 3))
            sq["dst"], sq["dst_stride"] = np.concatenate(ptrs), np.concatenate(strides)
            sq["crop0_size"] = self.roi
            sq["n_ops"] = self.N
            sq["flags"] = (_lib.SEQ_FAST if self.fast else 0) | (_lib.SEQ_STRICT if self.strict else 0)
            if len(self._tmpl_cache) > 256:
                self._tmpl_cache.clear()
            hit = self._tmpl_cache[key] = (sq, list(samples))
        sq = hit[0].copy()
        vol_b = np.repeat(np.arange(Bt), 2 * nc)
        vol_v = np.tile(np.repeat(np.arange(2), nc), Bt)
        sq["crop0_start"] = params["starts"][vol_v, vol_b]
        vox = int(np.prod([min(r, d) for r, d in zip(self.roi, shape)]))
        ops = sq["ops"]
        zeros_pad, bilinear = _lib.PADDING_MODES["zeros"], _lib.INTERP_MODES["bilinear"]
        for s, r in enumerate(self._slots(params, Bt, nc, vox)):
            o = ops[:, s]
            o["kind"] = np.where(r["has_aff"], _lib.OP_AFFINE, np.where(r["has_int"], _lib.OP_INTENSITY,
                                                                         np.where(r["has_phx"], _lib.OP_PHILOX, _lib.OP_NONE)))
            o["A"] = r["A"][:, :3].reshape(-1, 12)
            o["interp"], o["padding"] = bilinear, zeros_pad
            o["scale"], o["offset"] = r["sc"], r["of"]
            o["philox_std"], o["philox_seed"], o["philox_offset"] = r["std"], r["seed"], r["off_p"]
            ops[:, s] = o
        return sq, params

    def prepare_steps(self, batches: Sequence[Sequence[dict]], outs: Sequence[dict], params=None) -> "engine.PreparedSeqSteps":
        """Draw, compose and upload several consecutive steps at once (like a prefetching loader working ahead);
        ``run(k)`` then only enqueues step ``k``'s launches."""
        if not self.native_ok():
            raise NotImplementedError("injected (host-drawn) noise volumes are composed by BatchPlan: use __call__")
        self._ensure_views(self._n_channels(batches[0][0]))
        sq, _ = self.seqs(batches, outs, params)
        nc = self._n_channels(batches[0][0])
        dev = batches[0][0][self.keys[0]].device
        keep = [o[k] for o in outs for k in ("augmented_image_1", "augmented_image_2")] + [m[5] for b in batches for m in
                                                                                             (self._sample_meta(s, self.keys) for s in b)]
        return engine.prepare_seq_steps(sq, [len(b) * 2 * nc for b in batches], dev, keep=keep)

    def boxes(self, params, shape):
        """``box_1`` / ``box_2`` of the VICRegL variant: ``flatten_box(extra_info.cropped)``
        (augmentations.py:402-406)."""
        out = []
        for v in range(2):
            st = params["starts"][v]
            size = np.minimum(np.asarray(self.roi), np.asarray(shape))
            end_gap = np.asarray(shape)[None, :] - (st + size[None, :])
            out.append(np.concatenate([st, np.asarray(self.roi)[None, :] - end_gap], axis=1).astype(np.float32))
        return out

    def __call__(self, samples: Sequence[dict], params=None, out: dict | None = None, native: bool | None = None) -> dict:
        B, nc = len(samples), self._n_channels(samples[0])
        if native is None:
            native = self.native_ok()
        if native:
            shape = tuple(int(x) for x in self._sample_meta(samples[0], self.keys)[4][0])
            oshape = tuple(min(r, d) for r, d in zip(self.roi, shape))
            dev = samples[0][self.keys[0]].device
            if out is None:
                out = {k: torch.empty((B, nc, *oshape), dtype=torch.float32, device=dev)
                       for k in ("augmented_image_1", "augmented_image_2")}
            if params is None:
                params = self.draw(B, shape, nc)
            self.prepare_steps([samples], [out], params).run(0)
        else:
            plan, params = self.plan(samples, params)
            oshape = tuple(int(x) for x in plan.shape[0])
            if out is None:
                out = {k: torch.empty((B, nc, *oshape), dtype=torch.float32, device=plan.device)
                       for k in ("augmented_image_1", "augmented_image_2")}
            p1, s1 = self._dst_of(out["augmented_image_1"])
            p2, s2 = self._dst_of(out["augmented_image_2"])
            ptr = np.stack([p1, p2], axis=1).reshape(-1)          # [b, view, channel]
            stride = np.stack([s1, s2], axis=1).reshape(-1, 3)
            engine.execute_ptrs(plan, ptr, stride, keep=[out["augmented_image_1"], out["augmented_image_2"]])
        if self.vicregl:
            shape = tuple(int(x) for x in self._meta[id(samples[0])][4][0])
            b1, b2 = self.boxes(params, shape)
            out["box_1"], out["box_2"] = torch.from_numpy(b1), torch.from_numpy(b2)
        return out
