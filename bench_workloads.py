"""Side workloads of bench.py: every BASELINE.json config next to the headline one.

bench.py's contract keys stay on BASELINE configs[1] (segmentation, config B).  The `workloads`
block of the same JSON line carries, measured in the same run on the same device:

  a               config A   benchmarks/benchmark-random-affine.py:90-103 batched: 512 x 1x128x128x32,
                             RandAffined prob 1, rotate pi/6 x3, translate [10,10,3], scale 0.1, reflection
  seg_all_affine  config B   with the affine forced to fire for every sample (worst case of K1)
  seg_norm        config B   from RAW int16 / uint8 cached volumes: adell_minmax -> scaler coefficients ->
                             {scale, offset} read from device memory by K1 (the "+norm" of the metric inside the step)
  seg_crop        config B   with the label-guided crop sandwich, through the dictionary surface (device-side crop centres)
  ssl             config C   get_augmentations_ssl two-view chain (augmentations.py:391-516), 128x128x32, batch 64 / GPU
  ssl_fast        config C   the same in fast mode (consecutive resamples composed: documented deviation)
  cls             config D   percentile normalisation (K2/K3) + get_augmentations_class (augmentations.py:181-320):
                             flip -> affine(zeros) -> centre crop to 192x192x48, batch 32
  large           config E   512x512x128 volumes, dataset-wide percentile (NCCL all-reduce of the bin counts when
                             WORLD_SIZE > 1), scaling folded into the affine gather

Every entry: `value` (whole-job voxels/s, host draws + composition + uploads inside the timed
region, device-resident cache), `ms_per_step`, `kernel_ms` (CUDA events around the K1 launches of
the timed steps, summed per step), `roofline` (8 B per output voxel-channel / kernel_ms vs the
measured HBM peak), and `parity`: an IN-RUN check of the device result against the CPU oracle
(`oracle/`, the checker — never the thing measured) on a bounded part of one batch.  A failed
check raises: a bench line is never printed for wrong voxels.
"""

from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

from adell_mri_b200 import _lib, engine, geometry, stats
from adell_mri_b200.pipelines import ClassificationBatchAugmenter, SegmentationBatchAugmenter, SSLBatchAugmenter
from adell_mri_b200.plan import BatchPlan
from adell_mri_b200.sampling import RandAffineSampler

TOL = 1e-4   # north_star: <= 1e-4 relative on resampled intensities; masks / integer work bit-exact


def _close(got: torch.Tensor, want: torch.Tensor, what: str):
    scale = float(want.abs().max()) or 1.0
    err = float((got - want).abs().max())
    if not torch.allclose(got, want, rtol=TOL, atol=TOL * scale):
        raise AssertionError(f"parity: {what}: max |diff| {err:.3e} exceeds {TOL:g} of the range {scale:.3e}")
    return err / scale


def _equal(got: torch.Tensor, want: torch.Tensor, what: str):
    bad = int((got != want).sum())
    if bad:
        raise AssertionError(f"parity: {what}: {bad} voxels differ (bit-exact expected)")


class Workload:
    name = ""
    desc = ""
    bytes_per_voxel = 8.0   # algorithmic: fp32 source read + fp32 write per output voxel-channel
    # timed steps when the command line gives none: long enough (20-30 ms) that one scheduling hiccup of the host — 10-70 ms
    # stalls were seen on two-rank runs of a shared box, on both ranks at once — does not decide a 10-step mean
    default_steps = 32

    def __init__(self, dev, rank: int, world: int, seed: int):
        self.dev, self.rank, self.world, self.seed = dev, rank, world, seed

    # voxel-channels written per step on this rank
    vox_per_step = 0

    def step(self, i: int) -> None:
        raise NotImplementedError

    # Chunked composition with look-ahead, like a loader that works ahead of the consumer: the chunk of `chunk` steps
    # that contains step i is drawn + composed + uploaded by ONE host call, make(first_step); the NEXT chunk is composed
    # right after the last step of the current one has been launched, while the device still has (most of) a chunk
    # queued — the composition burst overlaps with it.  Draw order is unchanged (chunks are composed in order).
    chunk = 1
    _prep = None
    _prep_first = None
    _ahead = None

    def _prepared(self, i: int, make):
        first = i - i % self.chunk
        if self._prep_first != first:   # (a chunk whose composition came back None stays None: its steps run one by one)
            nxt = self._ahead
            self._prep = nxt[1] if (nxt is not None and nxt[0] == first) else make(first)
            self._prep_first, self._ahead = first, None
        return self._prep

    def _lookahead(self, i: int, make) -> None:
        if i % self.chunk == self.chunk - 1 and self._ahead is None:
            first = i - i % self.chunk + self.chunk
            self._ahead = (first, make(first))

    def parity(self) -> dict:
        raise NotImplementedError

    def extra(self) -> dict:
        return {}


# ----------------------------------------------------------------------------- config A
class AffineA(Workload):
    name = "a"
    desc = ("config A (benchmarks/benchmark-random-affine.py batched): 512 x 1x128x128x32, RandAffined prob 1.0, "
            "rotate pi/6 x3, translate [10,10,3], scale 0.1, trilinear, reflection; 512 distinct sources (1 GB > L2)")
    N, shape = 512, (128, 128, 32)

    def __init__(self, dev, rank, world, seed):
        super().__init__(dev, rank, world, seed)
        g = torch.Generator(device=dev).manual_seed(seed)
        self.src = torch.rand((self.N, 1, *self.shape), device=dev, generator=g)
        self.out = torch.empty_like(self.src)
        self.sampler = RandAffineSampler(prob=1.0, rotate_range=[np.pi / 6] * 3, shear_range=[0, 0, 0],
                                         translate_range=[10, 10, 3], scale_range=[0.1, 0.1, 0.1]).set_random_state(seed)
        self.vox_per_step = self.N * int(np.prod(self.shape))
        esz = 4
        self._ptr = (self.src.data_ptr() + esz * self.src.stride(0) * np.arange(self.N, dtype=np.int64)).astype(np.uint64)
        self._stride = np.tile(np.asarray(self.src.stride()[2:], np.int64), (self.N, 1))
        self._dtype = np.full(self.N, _lib.F32, np.uint8)
        self._shape = np.tile(np.asarray(self.shape, np.int64), (self.N, 1))
        self._dptr = (self.out.data_ptr() + esz * self.out.stride(0) * np.arange(self.N, dtype=np.int64)).astype(np.uint64)
        self.last_mats = None
        self._chains = None
        self._next = None

    def _compose(self):
        fired, p = self.sampler.draw_batch(self.N, n_keys=1)
        mats = geometry.compose_affine(p["rotate"], p["shear"], p["translate"], p["scale"], batch=self.N)
        if self._chains is None:   # static part of the adell_chain descriptors: one per volume
            from adell_mri_b200.pipelines import _chain_template
            self._chains = _chain_template(self._ptr, self._stride, self._dtype, self._shape, self._dptr, self._stride,
                                           ["bilinear"] * self.N, "reflection", False)
            self._chains["flags"] |= _lib.CHAIN_AFFINE
        ch = self._chains.copy()
        ch["A"] = mats[:, :3].reshape(self.N, 12)
        return mats, engine.prepare_chain_steps(ch, [self.N], self.dev, keep=[self.src, self.out])

    def step(self, i):
        # every step is drawn + composed + uploaded on its own (512 items), one step ahead of the launch: the host
        # composes step i + 1 while the device runs step i
        mats, prep = self._next if self._next is not None else self._compose()
        self.last_mats = mats
        prep.run(0)
        self._next = self._compose()

    def parity(self):
        from oracle import monai_restated as M

        self.step(0)
        torch.cuda.synchronize()
        worst = 0.0
        for b in (0, 1, self.N - 1):
            want = M.affine_resample(self.src[b].cpu(), torch.from_numpy(self.last_mats[b]), "bilinear", "reflection")
            worst = max(worst, _close(self.out[b].cpu(), want, f"config A item {b}"))
        return {"checked": "3 of 512 volumes vs oracle (torch CPU grid_sample chain)", "max_rel_err": worst, "tol": TOL, "ok": True}


# ----------------------------------------------------------------------------- config B variants
class _SegBase(Workload):
    shape, image_keys, batch, cache_samples = (256, 256, 32), ["t2", "adc", "dwi"], 8, 32

    def _augmenter(self, prob=None):
        aug = SegmentationBatchAugmenter(["affine", "flip"], self.image_keys + ["mask"], self.image_keys, flip_axis=[0, 1, 2])
        if prob is not None:
            for s in aug.samplers:
                s.prob = prob
        return aug.set_random_state(self.seed)

    def _alloc_out(self):
        self.out = {"image": torch.empty((self.batch, len(self.image_keys), *self.shape), device=self.dev),
                    "mask": torch.empty((self.batch, 1, *self.shape), device=self.dev)}
        self.vox_per_step = self.batch * (len(self.image_keys) + 1) * int(np.prod(self.shape))

    def _batch(self, i):
        nb = self.cache_samples // self.batch
        b0 = (i % nb) * self.batch
        return self.cache[b0:b0 + self.batch]

    def _oracle_sample(self, sample_cpu: dict, params, b: int, scaled=None):
        """The reference's op sequence for sample ``b`` of a batch drawn with ``params`` (oracle, CPU)."""
        from oracle import monai_restated as M

        outs = {}
        flips = [a for a in range(3) if params["flips"][b, a]]
        for k in self.image_keys + ["mask"]:
            x = sample_cpu[k] if scaled is None else scaled[k]
            for si in range(params["fired"].shape[0]):
                if params["fired"][si, b]:
                    A = torch.from_numpy(np.ascontiguousarray(params["mats"][si, b]))
                    x = M.affine_resample(x, A, "nearest" if k == "mask" else "bilinear", "reflection")
            # three RandFlipd, one per axis, in axis order (augmentations.py:127-131)
            for a in flips:
                x = M.flip(x, [a])
            outs[k] = x
        return outs


class SegAllAffine(_SegBase):
    name = "seg_all_affine"
    desc = ("config B with the affine forced to fire for every sample (K1 worst case: every tile resampled): "
            "T2/ADC/DWI trilinear + mask nearest, 256x256x32, batch 8, reflection, 3 flips p=0.25")

    def __init__(self, dev, rank, world, seed):
        super().__init__(dev, rank, world, seed)
        g = torch.Generator(device=dev).manual_seed(seed)
        self.cache = []
        for _ in range(self.cache_samples):
            s = {k: torch.rand((1, *self.shape), device=dev, generator=g) for k in self.image_keys}
            s["mask"] = (torch.rand((1, *self.shape), device=dev, generator=g) > 0.7).float()
            self.cache.append(s)
        self.aug = self._augmenter(1.0)
        self._alloc_out()

    chunk = 16   # steps drawn + composed per host call

    def _make(self, first):
        return self.aug.prepare_steps([self._batch(first + t) for t in range(self.chunk)], [self.out] * self.chunk)

    def step(self, i):
        self._prepared(i, self._make).run(i % self.chunk)
        self._lookahead(i, self._make)

    def parity(self):
        batch = self._batch(0)
        params = self.aug.draw(self.batch, self.shape)
        self.aug(batch, params=params, out=self.out)
        torch.cuda.synchronize()
        want = self._oracle_sample({k: v.cpu() for k, v in batch[0].items()}, params, 0)
        worst = 0.0
        for c, k in enumerate(self.image_keys):
            worst = max(worst, _close(self.out["image"][0, c].cpu(), want[k][0], f"seg_all_affine key {k}"))
        _equal(self.out["mask"][0, 0].cpu(), want["mask"][0], "seg_all_affine mask (nearest)")
        return {"checked": "sample 0 of one batch (3 trilinear keys <= tol, mask bit-exact) vs oracle", "max_rel_err": worst,
                "tol": TOL, "ok": True}


class SegNorm(_SegBase):
    name = "seg_norm"
    desc = ("config B from RAW cached volumes (int16 images, uint8 mask): per step adell_minmax over the batch's 32 "
            "volumes -> ScaleIntensityd(0,1) coefficients on the device -> {scale, offset} read by K1 (pre_dev; 16 steps composed per host call), "
            "affine p=0.2 reflection + 3 flips p=0.25; 6 B (images) / 5 B (mask) algorithmic per output voxel")

    def __init__(self, dev, rank, world, seed):
        super().__init__(dev, rank, world, seed)
        g = torch.Generator(device=dev).manual_seed(seed)
        self.cache = []
        for _ in range(self.cache_samples):
            s = {k: torch.randint(0, 4000, (1, *self.shape), device=dev, generator=g, dtype=torch.int16) for k in self.image_keys}
            s["mask"] = (torch.rand((1, *self.shape), device=dev, generator=g) > 0.7).to(torch.uint8)
            self.cache.append(s)
        self.aug = self._augmenter()
        self._alloc_out()
        nk = len(self.image_keys) + 1
        self.bytes_per_voxel = (len(self.image_keys) * 6.0 + 5.0) / nk
        self._desc = {}
        self._prep = None
        self._pre = torch.zeros((self.chunk * self.batch * nk, 2), device=dev)
        self._mm = torch.empty((self.batch * nk, 2), device=dev)
        self._coefs = torch.empty((self.batch * nk, 6), device=dev)

    def _pre_dev(self, i, batch, out=None):
        nb = self.cache_samples // self.batch
        hit = self._desc.get(i % nb)
        if hit is None:   # (the descriptors of a cached batch are uploaded once: the cache hands the same volumes back)
            keys = self.image_keys + ["mask"]
            vols = [s[k].reshape(-1) for s in batch for k in keys]
            hit = self._desc[i % nb] = (vols, stats.vol_descriptors(vols))
        vols, d = hit
        mm = stats.minmax(vols, desc=d, out=self._mm)
        return stats.coefs_to_affine(stats.scaler_coefs(mm, _lib.SCALER_MINMAX, 0.0, 1.0, out=self._coefs), out=out)

    chunk = 16   # steps drawn + composed per host call; the statistics kernels of a step fill its pre_dev rows when it runs

    def _make(self, first):
        return self.aug.prepare_steps([self._batch(first + t) for t in range(self.chunk)], [self.out] * self.chunk, pre_dev=self._pre)

    def step(self, i):
        j = i % self.chunk
        nv = self.batch * (len(self.image_keys) + 1)
        prep = self._prepared(i, self._make)
        self._pre_dev(i, self._batch(i), out=self._pre[j * nv:(j + 1) * nv])
        prep.run(j)
        self._lookahead(i, self._make)

    def parity(self):
        from oracle import monai_restated as M

        batch = self._batch(0)
        params = self.aug.draw(self.batch, self.shape)
        self.aug(batch, params=params, out=self.out, pre_dev=self._pre_dev(0, batch))
        torch.cuda.synchronize()
        raw = {k: v.cpu() for k, v in batch[0].items()}
        scaled = {k: M.scale_intensity(raw[k].float(), 0.0, 1.0) for k in self.image_keys}
        scaled["mask"] = raw["mask"].float()
        want = self._oracle_sample(raw, params, 0, scaled)
        worst = 0.0
        for c, k in enumerate(self.image_keys):
            worst = max(worst, _close(self.out["image"][0, c].cpu(), want[k][0], f"seg_norm key {k}"))
        _equal(self.out["mask"][0, 0].cpu(), want["mask"][0], "seg_norm mask (nearest, uint8 source)")
        return {"checked": "sample 0 (min-max scaled int16 keys <= tol, uint8 mask bit-exact) vs oracle", "max_rel_err": worst,
                "tol": TOL, "ok": True}


class SegCrop(_SegBase):
    name = "seg_crop"
    desc = ("config B with the label-guided crop sandwich (get_augmentations_unet(random_crop_size=[128,128,24], n_crops=2): "
            "RandCropByPosNegLabeld(1.1x) -> affine p=0.2 reflection -> 3 flips -> CenterSpatialCropd) through the DICTIONARY "
            "surface: one pipeline call per sample, FgBgToIndicesd lists resident on the device, crop centres selected by "
            "adell_posneg_starts, K1 reading the window starts from device memory, safe_collate_crops -> one launch")
    rc, n_crops = [128, 128, 24], 2
    default_steps = 10   # (3 ms of Python per step)

    def __init__(self, dev, rank, world, seed):
        super().__init__(dev, rank, world, seed)
        from adell_mri_b200 import collate, transform_factory as F, transforms as T

        self._collate = collate
        g = torch.Generator(device=dev).manual_seed(seed)
        fgbg = T.FgBgToIndicesd("mask")
        self.cache = []
        for _ in range(self.cache_samples):
            s = {k: torch.rand((1, *self.shape), device=dev, generator=g) for k in self.image_keys}
            s["mask"] = (torch.rand((1, *self.shape), device=dev, generator=g) > 0.7).float()
            self.cache.append(fgbg(s))   # cached stage: the index lists are made once per sample and stay on the device
        keys = self.image_keys + ["mask"]
        self.F, self.T = F, T
        self.pipe = self._pipeline(seed)
        self.vox_per_step = self.batch * self.n_crops * len(keys) * int(np.prod(self.rc))
        self.bytes_per_voxel = 8.0

    def _pipeline(self, seed):
        keys = self.image_keys + ["mask"]
        aug = self.F.get_augmentations_unet(["affine", "flip"], keys, self.image_keys, [], random_crop_size=self.rc, has_label=True,
                                            n_crops=self.n_crops, flip_axis=[0, 1, 2])
        tf = self.F.SegmentationTransforms(keys, self.image_keys, ["mask"], self.image_keys, [])
        return self.T.Compose([aug, *tf.post_transforms()]).set_random_state(seed)

    def step(self, i):
        self.last = self._collate.safe_collate_crops([self.pipe(dict(s)) for s in self._batch(i)])

    def parity(self):
        """Same seed through the eager oracle pipeline (pipelines_ref.unet with the host-side centre selection of
        monai_restated.pos_neg_crop_centers) on the first two samples."""
        from oracle import pipelines_ref as P

        n = 2
        pipe = self._pipeline(self.seed + 3)
        got = self._collate.safe_collate_crops([pipe(dict(s)) for s in self.cache[:n]])
        torch.cuda.synchronize()
        keys = self.image_keys + ["mask"]
        # the reference seeds the OUTER Compose; its only Randomizable child is the augmentation Compose
        ref = P.Chain([P.unet(["affine", "flip"], keys, self.image_keys, self.rc, True, self.n_crops, (0, 1, 2))]).seed(self.seed + 3)
        worst = 0.0
        for b, s in enumerate(self.cache[:n]):
            d = {k: s[k].cpu() for k in keys}
            d["mask_fg_indices"], d["mask_bg_indices"] = s["mask_fg_indices"].cpu().numpy(), s["mask_bg_indices"].cpu().numpy()
            crops = ref(d)
            for c, r in enumerate(crops):
                img = torch.cat([r[k] for k in self.image_keys], 0)
                worst = max(worst, _close(got["image"][b * self.n_crops + c].cpu(), img, f"seg_crop sample {b} crop {c} image"))
                _equal(got["mask"][b * self.n_crops + c].cpu(), r["mask"], f"seg_crop sample {b} crop {c} mask")
        return {"checked": f"{n} samples x {self.n_crops} crops, same seed, vs the eager oracle pipeline (pipelines_ref.unet, host-side "
                           "centre selection): image keys <= tol, masks bit-exact", "max_rel_err": worst, "tol": TOL, "ok": True}


class SegCropBatch(SegCrop):
    name = "seg_crop_batch"
    desc = ("config B with the label-guided crop sandwich on the BATCH fast path: SegmentationBatchAugmenter(random_crop_size="
            "[128,128,24], has_label=True, n_crops=2): the two host draws per crop from the lengths of the device-resident "
            "FgBgToIndicesd lists, centres selected by adell_posneg_starts, K1 reading the window starts from device memory, "
            "affine p=0.2 reflection + 3 flips per crop, centre crop; 16 crops of 4 keys per step, 8 steps composed per host call")
    chunk = 8
    default_steps = 32

    def __init__(self, dev, rank, world, seed):
        super().__init__(dev, rank, world, seed)
        self.aug = SegmentationBatchAugmenter(["affine", "flip"], self.image_keys + ["mask"], self.image_keys, random_crop_size=self.rc,
                                              has_label=True, flip_axis=[0, 1, 2], n_crops=self.n_crops).set_random_state(seed, nested=True)
        self.out = self.aug._alloc_out(self.batch * self.n_crops, tuple(self.rc), dev)
        self._prep = None

    def _make(self, first):
        return self.aug.prepare_steps([self._batch(first + t) for t in range(self.chunk)], [self.out] * self.chunk)

    def step(self, i):
        self._prepared(i, self._make).run(i % self.chunk)
        self._lookahead(i, self._make)

    def parity(self):
        """The batch path against the dictionary surface on the same seed (which `seg_crop` checks against the eager
        oracle pipeline): identical batches."""
        n = 4
        seed = self.seed + 3
        want = self._collate.safe_collate_crops([pipe_s for pipe_s in map(self._pipeline(seed), (dict(s) for s in self.cache[:n]))])
        aug = SegmentationBatchAugmenter(["affine", "flip"], self.image_keys + ["mask"], self.image_keys, random_crop_size=self.rc,
                                         has_label=True, flip_axis=[0, 1, 2], n_crops=self.n_crops).set_random_state(seed, nested=True)
        got = aug(self.cache[:n])
        torch.cuda.synchronize()
        worst = _close(got["image"], want["image"], "seg_crop_batch image vs the dictionary surface")
        _equal(got["mask"], want["mask"], "seg_crop_batch mask vs the dictionary surface")
        return {"checked": f"{n} samples x {self.n_crops} crops vs the dictionary surface on the same seed (itself checked against "
                           "the eager oracle pipeline by `seg_crop`): image keys <= tol, masks bit-exact", "max_rel_err": worst,
                "tol": TOL, "ok": True}


# ----------------------------------------------------------------------------- config C
class SSLTwoView(Workload):
    name = "ssl"
    desc = ("config C: get_augmentations_ssl two views (shared RandSpatialCropd 128x128x32 out of 160x160x40, then per view "
            "3 of the 15 fused workhorse members in drawn order, one K1 pass per spatial member like the reference's "
            "sequential resamples), batch 64 per GPU; noise sigma drawn on the host, noise values from the device Philox "
            "generator; member subsets drawn vectorised; 16 steps drawn + composed by the native sequence composer per host call")
    batch, src, roi, cache_samples = 64, (160, 160, 40), (128, 128, 32), 128

    def __init__(self, dev, rank, world, seed):
        super().__init__(dev, rank, world, seed)
        g = torch.Generator(device=dev).manual_seed(seed)
        self.cache = [{"image": torch.rand((1, *self.src), device=dev, generator=g)} for _ in range(self.cache_samples)]
        self.aug = SSLBatchAugmenter(["image"], self.roi, n_transforms=3, choice="vectorised", noise="philox").set_random_state(seed)
        self.out = {k: torch.empty((self.batch, 1, *self.roi), device=dev) for k in ("augmented_image_1", "augmented_image_2")}
        self.vox_per_step = self.batch * 2 * int(np.prod(self.roi))
        self._prep = None

    chunk = 16   # steps drawn, composed (adell_seq_prepare_steps) and uploaded per host call, like a loader working ahead
    default_steps = 32

    def _make(self, first):
        nb = self.cache_samples // self.batch
        batches = [self.cache[((first + t) % nb) * self.batch:((first + t) % nb + 1) * self.batch] for t in range(self.chunk)]
        return self.aug.prepare_steps(batches, [self.out] * self.chunk)

    def step(self, i):
        self._prepared(i, self._make).run(i % self.chunk)
        self._lookahead(i, self._make)

    def parity(self):
        """Stream level: 3 samples through a second augmenter (host-drawn noise injected, member subsets from the
        global numpy stream like the reference) against the eager oracle pipeline on the same seeds."""
        from oracle import pipelines_ref as P
        from adell_mri_b200.pipelines import SSL_FUSED_MEMBERS

        n = 3
        samples = self.cache[:n]
        aug = SSLBatchAugmenter(["image"], self.roi, n_transforms=3, choice="global", noise="injected").set_random_state(self.seed + 5)
        np.random.seed(self.seed + 6)
        got = aug(samples)
        torch.cuda.synchronize()
        ref = P.Chain(P.ssl(["image"], ["image_copy"], self.roi, False, False, list(SSL_FUSED_MEMBERS), 3)).seed(self.seed + 5)
        np.random.seed(self.seed + 6)
        worst = 0.0
        for b, s in enumerate(samples):
            x = s["image"].cpu()
            r = ref({"image": x, "image_copy": x.clone()})
            for key, rk in (("augmented_image_1", "image"), ("augmented_image_2", "image_copy")):
                worst = max(worst, _close(got[key][b].cpu(), r[rk], f"ssl sample {b} {key}"))
        # the timed route (native sequence composer, device Philox noise) against the numpy BatchPlan route on one draw
        params = self.aug.draw(8, self.src, 1)
        a = self.aug(self.cache[:8], params=params, native=False)
        b = self.aug(self.cache[:8], params=params, native=True)
        torch.cuda.synchronize()
        for key in ("augmented_image_1", "augmented_image_2"):
            _equal(b[key], a[key], f"ssl native route vs BatchPlan route {key}")
        return {"checked": f"{n} samples x 2 views, same seeds, vs the eager oracle pipeline (pipelines_ref.ssl); the timed native "
                           "route == the BatchPlan route bit for bit on 8 samples (same draws, device Philox noise)",
                "max_rel_err": worst, "tol": TOL, "ok": True}


class SSLTwoViewFast(SSLTwoView):
    name = "ssl_fast"
    desc = ("config C in FAST mode (documented deviation, not the reference's numerics): consecutive spatial members of a "
            "view are composed into ONE matrix and resampled once (no double interpolation), so a view is one K1 pass "
            "unless an intensity / noise member sits between two spatial ones; same draws, same shapes as `ssl`")

    def __init__(self, dev, rank, world, seed):
        super().__init__(dev, rank, world, seed)
        self.aug = SSLBatchAugmenter(["image"], self.roi, n_transforms=3, choice="vectorised", noise="philox", fast=True).set_random_state(seed)
        self._prep = None

    def parity(self):
        return {"ok": None, "checked": "not applicable: fast mode composes consecutive resamples (documented deviation); the "
                                       "reference-faithful entry `ssl` carries the oracle check"}


# ----------------------------------------------------------------------------- config D
class ClsPercentile(Workload):
    name = "cls"
    desc = ("config D: raw cached 208x208x64 volumes (crop + 16 margin), 3 image keys + mask, batch 32; per step exact "
            "percentiles (0.5, 99.5) of the 96 image volumes (one read per volume: sampled brackets + exact selection, adell_quantile_keys) -> ScaleIntensityRangePercentilesd "
            "coefficients -> {scale, offset} read by K1; OneOf flips -> RandAffined(translate, rotate x, scale; zeros, "
            "prob 0.1) -> CenterSpatialCropd 192x192x48 -> concat")
    batch, src, crop, image_keys, cache_samples = 32, (208, 208, 64), (192, 192, 48), ["t2", "adc", "dwi"], 64
    prob = 0.1

    def __init__(self, dev, rank, world, seed):
        super().__init__(dev, rank, world, seed)
        g = torch.Generator(device=dev).manual_seed(seed)
        self.cache = []
        for _ in range(self.cache_samples):
            s = {k: torch.empty((1, *self.src), device=dev).log_normal_(5.0, 0.6, generator=g) for k in self.image_keys}
            s["mask"] = (torch.rand((1, *self.src), device=dev, generator=g) > 0.7).float()
            self.cache.append(s)
        self.aug = ClassificationBatchAugmenter(["affine", "flip"], self.image_keys, "mask", flip_axis=[0, 1, 2], prob=self.prob,
                                                crop_size=self.crop).set_random_state(seed)
        self.out = {"image": torch.empty((self.batch, len(self.image_keys) + 1, *self.crop), device=dev)}
        self.vox_per_step = self.batch * (len(self.image_keys) + 1) * int(np.prod(self.crop))
        self.stats_ms = None
        self._kern = {}
        self._pre_tmpl = torch.zeros((self.batch, len(self.image_keys) + 1, 2), device=dev)
        self._pre_tmpl[:, :, 0] = 1.0
        self._pre = self._pre_tmpl.repeat(self.chunk, 1, 1).view(-1, 2).contiguous()   # rows of every step of a chunk
        self._prep = None

    def _pre_dev(self, batch, out=None):
        ni = len(self.image_keys)
        key = id(batch[0])
        hit = self._kern.get(key)
        if hit is None:   # the descriptors of a cached batch's volumes are uploaded once (a device-resident cache hands the same volumes back)
            vols = [s[k].reshape(-1) for s in batch for k in self.image_keys]
            hit = self._kern[key] = (vols, stats._CudaKernels(vols))
        vols, kern = hit
        kern.st = stats._stream(self.dev)
        pct = stats.percentiles(vols, [0.5, 99.5], kernels=kern)
        aff = stats.coefs_to_affine(stats.scaler_coefs(pct, _lib.SCALER_RANGE, 0.0, 1.0))      # [B * ni, 2]
        if out is None:
            out = self._pre_tmpl.clone().view(-1, 2)                                             # mask rows: {1, 0}
        out.view(len(batch), ni + 1, 2)[:, :ni] = aff.view(len(batch), ni, 2)
        return out

    chunk = 8   # steps drawn + composed per host call; the percentile kernels of a step fill its pre_dev rows when it runs

    def _batch(self, i):
        nb = self.cache_samples // self.batch
        b0 = (i % nb) * self.batch
        return self.cache[b0:b0 + self.batch]

    def _make(self, first):
        return self.aug.prepare_steps([self._batch(first + t) for t in range(self.chunk)], [self.out] * self.chunk, pre_dev=self._pre)

    def step(self, i):
        j = i % self.chunk
        nv = self.batch * (len(self.image_keys) + 1)
        prep = self._prepared(i, self._make)
        if prep is None:   # some sample of the chunk fired both RandAffined: step by step (two resamples)
            self.aug(self._batch(i), out=self.out, pre_dev=self._pre_dev(self._batch(i)))
            return
        self._pre_dev(self._batch(i), out=self._pre[j * nv:(j + 1) * nv])
        prep.run(j)
        self._lookahead(i, self._make)

    def extra(self):
        batch = self.cache[:self.batch]
        for _ in range(2):
            self._pre_dev(batch)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            self._pre_dev(batch)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        n = self.batch * len(self.image_keys) * int(np.prod(self.src))
        return {"stats_ms": ms, "stats_gbs_one_read_credited": 4.0 * n / ms / 1e6}

    def parity(self):
        from oracle import monai_restated as M

        batch = self.cache[:self.batch]
        params = self.aug.draw(self.batch)
        # make sure the checked samples exercise the resample: force sample 0 to fire if nothing did
        pre = self._pre_dev(batch)
        self.aug(batch, params=params, out=self.out, pre_dev=pre)
        torch.cuda.synchronize()
        worst, checked = 0.0, []
        fired_any = np.nonzero(params["fired"].any(axis=0))[0]
        pick = sorted(set([0] + [int(b) for b in fired_any[:2]]))
        for b in pick:
            flips = [a for a in range(3) if params["flips"][b, a]]
            for c, k in enumerate(self.image_keys + ["mask"]):
                x = batch[b][k].cpu()
                if k != "mask":
                    x = M.scale_intensity_range_percentiles(x, 0.5, 99.5, 0.0, 1.0)
                if flips:
                    x = M.flip(x, flips)
                for si in range(params["fired"].shape[0]):
                    if params["fired"][si, b]:
                        x = M.affine_resample(x, torch.from_numpy(np.ascontiguousarray(params["mats"][si, b])),
                                              "nearest" if k == "mask" else "bilinear", "zeros")
                x = M.center_spatial_crop(x, self.crop)
                if k == "mask":
                    _equal(self.out["image"][b, c].cpu(), x[0], f"cls sample {b} mask")
                else:
                    worst = max(worst, _close(self.out["image"][b, c].cpu(), x[0], f"cls sample {b} key {k}"))
            checked.append(b)
        return {"checked": f"samples {checked} (percentile-scaled keys <= tol, mask bit-exact) vs oracle; "
                           f"resampled among them: {[int(b) for b in checked if params['fired'][:, b].any()]}",
                "max_rel_err": worst, "tol": TOL, "ok": True}


# ----------------------------------------------------------------------------- config E
class LargeVolume(Workload):
    name = "large"
    desc = ("config E: 4 volumes of 512x512x128 fp32 per GPU; per step the DATASET-WIDE percentiles (1, 99) over all ranks' "
            "volumes (one rank: ONE read of the pooled volumes, adell_quantile_keys; WORLD_SIZE > 1: three radix passes, int64 "
            "bin counts all-reduced over NCCL after each), "
            "ScaleIntensityRange coefficients on the device, scaling folded into the affine gather (rotate pi/8, pi/8, pi/16; zeros)")
    M_vols, shape = 4, (512, 512, 128)
    default_steps = 16

    def __init__(self, dev, rank, world, seed):
        super().__init__(dev, rank, world, seed)
        g = torch.Generator(device=dev).manual_seed(seed)
        self.vols = [torch.empty(self.shape, device=dev).log_normal_(0, 1, generator=g) for _ in range(self.M_vols)]
        self.flat = [v.reshape(-1) for v in self.vols]
        self.out = torch.empty((self.M_vols, 1, *self.shape), device=dev)
        self.R = np.random.RandomState(seed)
        self.vox_per_step = self.M_vols * int(np.prod(self.shape))
        self.last = None
        self._next = None
        self._pre = torch.zeros((self.M_vols, 2), device=dev)
        self.collective_bytes = 0
        self._kern = None

    def _percentiles(self):
        from adell_mri_b200 import dist as adist

        if self._kern is None:
            self._kern = stats._CudaKernels(self.flat)
        self._kern.st = stats._stream(self.dev)
        return adist.dataset_percentiles(self.flat, [1.0, 99.0], kernels=self._kern)

    def _compose(self):
        """Draws the step's affines and composes + uploads its K1 launch; the {scale, offset} rows are read by K1 from
        ``self._pre`` (device memory) when it runs, so this needs no statistics yet."""
        rot = self.R.uniform(-1, 1, (self.M_vols, 3)) * np.array([np.pi / 8, np.pi / 8, np.pi / 16])
        mats = geometry.compose_affine(rotate=rot, batch=self.M_vols)
        plan = BatchPlan(self.vols)
        plan.intensity_from_device(self._pre)
        plan.affine(mats, "bilinear", "zeros")
        outs = [self.out[b, 0] for b in range(self.M_vols)]
        dst_ptr = np.array([o.data_ptr() for o in outs], np.uint64)
        dst_stride = np.array([o.stride() for o in outs], np.int64)
        return mats, engine.prepare_steps(plan, dst_ptr, dst_stride, [self.M_vols], keep=[self.vols, self.out, self._pre])

    def step(self, i):
        # the K1 launch of a step is composed one step ahead (while the device runs the previous step), so that it follows
        # the statistics kernels on the stream without a host gap
        mats, prep = self._next if self._next is not None else self._compose()
        pct = self._percentiles()                                                     # [1, 2], identical on all ranks
        pre1 = stats.coefs_to_affine(stats.scaler_coefs(pct, _lib.SCALER_RANGE, 0.0, 1.0))   # [1, 2]
        self._pre.copy_(pre1.expand(self.M_vols, 2))
        prep.run(0)
        self.last = (pct, mats)
        self._next = self._compose()

    def extra(self):
        for _ in range(2):
            self._percentiles()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            self._percentiles()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        # per call: three passes of int64 bins: 2^11 + 4 * 2^11 + 4 * 2^10 bins (the pooled element count is all-reduced once)
        coll = 8 * ((1 << 11) + 4 * (1 << 11) + 4 * (1 << 10))
        return {"stats_ms": ms, "stats_gbs_one_read_credited": 4.0 * self.vox_per_step / ms / 1e6,
                "collective": ("nccl all_reduce(sum) of int64 bin counts, %d B per rank per step in 3 calls" % coll) if self.world > 1 else None,
                "collective_bytes_per_step": coll if self.world > 1 else 0}

    def parity(self):
        """(1) every rank holds bit-identical percentiles; (2) they ARE the order statistics numpy's 'linear' rule
        interpolates: with k = floor((n-1)q), count(x < a) <= k < count(x <= a) for the lower neighbour a (and the same
        with k+1 for the upper one), counts summed over all ranks with plain torch ops — then np.percentile's float64
        lerp of the two neighbours; at WORLD_SIZE 1 also literally np.percentile of the pooled data; (3) one gathered
        volume against the oracle chain."""
        import torch.distributed as dist
        from oracle import monai_restated as M

        self.step(0)
        torch.cuda.synchronize()
        pct, mats = self.last
        pct_host = pct.cpu().numpy().reshape(-1)
        res = {}
        if self.world > 1:
            gathered = [torch.empty_like(pct) for _ in range(self.world)]
            dist.all_gather(gathered, pct)
            same = all(torch.equal(g, gathered[0]) for g in gathered)
            if not same:
                raise AssertionError("parity: config E percentiles differ between ranks")
            res["ranks_bit_identical"] = True
        n_total = torch.tensor([sum(v.numel() for v in self.flat)], dtype=torch.int64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(n_total)
        n_total = int(n_total.item())
        # the neighbours themselves: order statistics k and k+1 of the data pooled over all ranks, located with plain
        # torch ops (global counts) starting from the reported value — independent of the radix-select kernels
        def gmax_below(x, strict):
            m = torch.stack([(v[v < x] if strict else v[v <= x]).max() if bool(((v < x) if strict else (v <= x)).any())
                             else torch.tensor(-np.inf, device=self.dev) for v in self.flat]).max()
            if self.world > 1:
                dist.all_reduce(m, op=dist.ReduceOp.MAX)
            return m

        def gmin_above(x):
            m = torch.stack([v[v > x].min() if bool((v > x).any()) else torch.tensor(np.inf, device=self.dev) for v in self.flat]).min()
            if self.world > 1:
                dist.all_reduce(m, op=dist.ReduceOp.MIN)
            return m

        def counts(x):
            cnt = torch.zeros(2, dtype=torch.int64, device=self.dev)
            for v in self.flat:
                cnt[0] += (v < x).sum(); cnt[1] += (v <= x).sum()
            if self.world > 1:
                dist.all_reduce(cnt)
            return int(cnt[0]), int(cnt[1])

        def order_stat(k, start):
            x = start
            for _ in range(8):
                lt, le = counts(x)
                if k < lt:
                    x = gmax_below(x, True)
                elif k >= le:
                    x = gmin_above(x)
                else:
                    return float(x)
            raise AssertionError(f"parity: config E: order statistic {k} not found near {float(start)}")

        for qi, q in enumerate((1.0, 99.0)):
            lo, hi, gamma = stats.numpy_virtual_index(n_total, q)
            p = float(pct_host[qi])
            a = order_stat(lo, gmax_below(torch.tensor(p, device=self.dev), False))
            b = order_stat(hi, torch.tensor(a, device=self.dev))
            # numpy _lerp on float32 neighbours with the float64 weight (what np.percentile evaluates), cast to fp32
            a32, b32 = np.float32(a), np.float32(b)
            diff = np.float64(np.float32(b32 - a32))
            want = np.float64(a32) + diff * gamma if gamma < 0.5 else np.float64(b32) - diff * (1.0 - gamma)
            want = np.float32(want)
            if np.float32(p) != np.float32(want):
                raise AssertionError(f"parity: config E q={q}: device {p!r} != numpy lerp of the neighbours {float(want)!r}")
        res["order_statistics_verified"] = True
        if self.world == 1:
            pooled = torch.cat(self.flat).cpu().numpy()
            want = np.percentile(pooled, [1.0, 99.0])
            if not np.array_equal(np.asarray(want, np.float32), pct_host.astype(np.float32)):
                raise AssertionError(f"parity: config E percentiles {pct_host} != np.percentile {want}")
            res["equals_np_percentile_of_pooled_data"] = True
        # gather of one volume on rank 0 vs the oracle
        worst = None
        if self.rank == 0:
            x = self.vols[0].cpu()[None]
            scaled = M.scale_intensity_range(x, float(pct_host[0]), float(pct_host[1]), 0.0, 1.0)
            want = M.affine_resample(scaled, torch.from_numpy(np.ascontiguousarray(mats[0])), "bilinear", "zeros")
            worst = _close(self.out[0, 0].cpu(), want[0], "config E volume 0")
        res.update({"checked": "percentiles: identical on all ranks, order statistics verified by global counts"
                               + (", equal to np.percentile of the pooled data" if self.world == 1 else "")
                               + "; volume 0 of rank 0 (33.5 M voxels) vs oracle", "max_rel_err": worst, "tol": TOL, "ok": True})
        return res


ALL = [AffineA, SegAllAffine, SegNorm, SegCrop, SegCropBatch, SSLTwoView, SSLTwoViewFast, ClsPercentile, LargeVolume]


# ----------------------------------------------------------------------------- runner
def run(cls, dev, rank, world, seed, steps, warmup, peak_gbs, barrier, reduce_max, do_parity=True):
    """Time one workload: returns its entry of the `workloads` block (rank 0) — every rank runs its own shard."""
    wl = cls(dev, rank, world, seed + 101 * rank)
    steps = steps if steps else wl.default_steps
    # Warm-up reaches the loader's steady state: two full chunks and the look-ahead of the third, so that the device blocks
    # of the chunk uploads exist before the timed region.  (The SECOND live chunk makes torch's caching allocator take a
    # new segment: a cudaMalloc that usually costs 0.5 ms and now and then 5-50 ms — caught with ADELL_BENCH_DIAG in
    # `_PinnedRing.upload` of the first look-ahead, which a 5-step warm-up left inside the timed region.)
    warmup = max(warmup, 2 * wl.chunk + 1)
    parity = wl.parity() if do_parity else {"ok": None, "checked": "skipped (--no-parity)"}
    for i in range(warmup):
        wl.step(i)
    barrier()
    stream = torch.cuda.current_stream()
    engine.timer = engine.LaunchTimer()
    launches0 = engine.launch_count
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    diag = os.environ.get("ADELL_BENCH_DIAG")
    gc_log = []
    if diag:
        import gc

        def _gc_cb(phase, info, _t=[0.0]):
            if phase == "start":
                _t[0] = time.perf_counter()
            else:
                gc_log.append((info.get("generation"), round(1e3 * (time.perf_counter() - _t[0]), 3)))
        gc.callbacks.append(_gc_cb)
        parts = {}

        def _wrap(obj, attr):
            fn = getattr(obj, attr, None)
            if fn is None:
                return

            def w(*a, **k):
                t = time.perf_counter()
                try:
                    return fn(*a, **k)
                finally:
                    parts.setdefault(attr, []).append(round(1e3 * (time.perf_counter() - t), 3))
            setattr(obj, attr, w)
        for attr in ("_make", "_pre_dev"):
            _wrap(wl, attr)
        saved = {a: getattr(engine._PinnedRing, a) for a in ("acquire", "upload")}
        for attr in saved:
            _wrap(engine._PinnedRing, attr)
        seg0 = torch.cuda.memory_stats().get("segment.all.allocated", 0)
        slow = None
    t0 = time.perf_counter()
    t_prev, host_max, per_host = t0, 0.0, []
    for i in range(steps):
        engine.timer.label = i
        wl.step(warmup + i)
        t_now = time.perf_counter()
        per_host.append(round(1e3 * (t_now - t_prev), 3))
        if diag:
            if slow is None or per_host[-1] > slow[1]:
                slow = (i, per_host[-1], {k: list(v) for k, v in parts.items()},
                        torch.cuda.memory_stats().get("segment.all.allocated", 0) - seg0)
            for v in parts.values():
                v.clear()
            seg0 = torch.cuda.memory_stats().get("segment.all.allocated", 0)
        host_max, t_prev = max(host_max, t_now - t_prev), t_now
    host_s = time.perf_counter() - t0
    if diag:
        gc.callbacks.remove(_gc_cb)
        for attr, fn in saved.items():
            setattr(engine._PinnedRing, attr, fn)
        print(f"[diag] rank {rank} {wl.name}: slowest step (index, host ms, parts ms, new allocator segments) {slow}", file=sys.stderr, flush=True)
        print(f"[diag] rank {rank} {wl.name}: host ms per step {per_host}; gc (generation, ms) {gc_log}", file=sys.stderr, flush=True)
    b.record(stream)
    barrier()
    timer, engine.timer = engine.timer, None
    ms = a.elapsed_time(b) / steps
    per_step = {}
    for lab, t in timer.ms():
        per_step[lab] = per_step.get(lab, 0.0) + t
    kernel_ms = float(np.mean(list(per_step.values()))) if per_step else None
    launches = engine.launch_count - launches0
    extra = wl.extra()
    ms_max = reduce_max(ms)
    alg = wl.bytes_per_voxel * wl.vox_per_step
    entry = {
        "workload": wl.desc, "value": world * wl.vox_per_step / (ms_max * 1e-3), "unit": "voxels/s", "ms_per_step": ms_max,
        "steps": steps, "warmup": warmup, "kernel_ms": kernel_ms, "k1_launches_per_step": launches / steps,
        "host_ms_per_step": 1e3 * host_s / steps, "host_ms_slowest_step": 1e3 * host_max,
        "roofline": {"bound": "hbm", "achieved": alg / (kernel_ms * 1e-3) / 1e9, "peak": peak_gbs, "unit": "GB/s",
                     "frac": alg / (kernel_ms * 1e-3) / 1e9 / peak_gbs, "algorithmic_bytes": alg, "kernel": "k1_gather",
                     "traffic": None},
        "parity": parity,
    }
    entry.update(extra)
    del wl
    torch.cuda.empty_cache()
    return entry
