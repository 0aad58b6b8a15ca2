"""CPU: the native chain composer (csrc/compose.cu: adell_chain -> adell_item) against the numpy composer
(plan.BatchPlan) — byte for byte on the composed + planned items — for the batch augmenters that route their
single-resample chains through it.  (No driver needed: both sides run the host policy without encoding.)"""

import ctypes as C

import numpy as np
import pytest
import torch

from adell_mri_b200 import _lib, engine
from adell_mri_b200.pipelines import ClassificationBatchAugmenter, SegmentationBatchAugmenter
from adell_mri_b200.plan import ITEM_DTYPE


def _plan_items(plan, dst_ptr, dst_stride, sizes):
    """BatchPlan route: items of the last (only) launch, then adell_aug_plan per step."""
    items = plan.build_launches(dst_ptr, dst_stride, None)[-1].copy()
    out, start = [], 0
    for n in sizes:
        it = items[start:start + n].copy()
        tiles = np.zeros(n + 5, np.int32)
        info = _lib.LaunchInfo()
        _lib.check(_lib.load().adell_aug_plan(it.ctypes.data, n, tiles.ctypes.data, C.byref(info)), "plan")
        out.append((it, tiles, info))
        start += n
    return out


def _native_items(ch, sizes):
    buf, offs, infos = engine.compose_chains_host(ch, sizes)
    out = []
    for k, n in enumerate(sizes):
        o = int(offs[k])
        it = buf[o:o + n * engine.ISZ].view(ITEM_DTYPE)
        tiles = buf[o + n * engine.ISZ:o + n * engine.ISZ + 4 * (n + 5)].view(np.int32)
        out.append((it, tiles, infos[k]))
    return out


def _same(a, b):
    for (ia, ta, fa), (ib, tb, fb) in zip(a, b):
        assert ia.tobytes() == ib.tobytes()
        assert np.array_equal(ta, tb)
        assert (fa.total_tiles, fa.smem_bytes, fa.n_staged, fa.first_copy_tile) == (fb.total_tiles, fb.smem_bytes, fb.n_staged, fb.first_copy_tile)


def _samples(R, n, keys, shape, dtype=np.float32):
    out = []
    for _ in range(n):
        s = {k: torch.from_numpy((R.rand(1, *shape) * 100).astype(dtype)) for k in keys}
        s["mask"] = torch.from_numpy((R.rand(1, *shape) > 0.7).astype(np.float32))
        out.append(s)
    return out


@pytest.mark.parametrize("augment,crop,strict", [
    (["affine", "flip"], None, False), (["affine", "flip"], None, True), (["flip"], None, False),
    (["affine", "shear", "flip"], None, False), (["affine", "flip"], [20, 24, 8], False), (["affine"], [40, 40, 40], False),
])
def test_segmentation_chains_equal_batchplan_items(augment, crop, strict):
    R = np.random.RandomState(1)
    keys = ["t2", "adc"]
    shape = (36, 40, 12)
    has_label = crop is None
    samples = _samples(R, 12, keys, shape)
    aug = SegmentationBatchAugmenter(augment, keys + ["mask"], keys, random_crop_size=crop, has_label=has_label,
                                     flip_axis=[0, 1, 2], strict=strict).set_random_state(3)
    for s in aug.samplers:
        s.prob = 0.5 if len(aug.samplers) == 1 else 0.08
    oshape = shape if crop is None else tuple(min(c, s) for c, s in zip(crop, shape))
    outs = [aug._alloc_out(4, oshape, torch.device("cpu")) for _ in range(3)]
    batches = [samples[0:4], samples[4:8], samples[8:12]]
    hit = 0
    for trial in range(6):
        params = aug.draw(12, shape)
        ch, _ = aug.chains(batches, outs, params)
        if ch is None:   # some sample fired both RandAffined: multi-pass, not on the native route
            assert len(aug.samplers) == 2 and (params["fired"].sum(0) > 1).any()
            continue
        hit += 1
        plan = aug.plan([s for b in batches for s in b], params)
        assert not plan.passes
        ptrs, strides = zip(*[aug._dst(o, 4) for o in outs])
        nk = len(aug.keys)
        sizes = [4 * nk] * 3
        _same(_native_items(ch, sizes), _plan_items(plan, np.concatenate(ptrs), np.concatenate(strides), sizes))
    assert hit >= 1


def test_segmentation_chains_carry_device_intensity_pointers_and_integer_sources():
    R = np.random.RandomState(2)
    keys = ["t2"]
    samples = _samples(R, 4, keys, (32, 32, 16), np.int16)
    aug = SegmentationBatchAugmenter(["affine", "flip"], keys + ["mask"], keys, flip_axis=[0, 1, 2]).set_random_state(5)
    out = aug._alloc_out(4, (32, 32, 16), torch.device("cpu"))
    pre = torch.rand(8, 2)
    params = aug.draw(4, (32, 32, 16))
    ch, _ = aug.chains([samples], [out], params, pre_dev=pre)
    plan = aug.plan(samples, params, pre_dev=pre)
    p, s = aug._dst(out, 4)
    _same(_native_items(ch, [8]), _plan_items(plan, p, s, [8]))


@pytest.mark.parametrize("augment", [["flip", "affine"], ["affine"], ["flip"]])
def test_classification_chains_equal_batchplan_items(augment):
    R = np.random.RandomState(4)
    keys = ["t2", "adc"]
    shape, crop = (40, 36, 24), [24, 20, 8]
    samples = _samples(R, 6, keys, shape)
    aug = ClassificationBatchAugmenter(augment, keys, "mask", flip_axis=[0, 1, 2], prob=0.6, crop_size=crop).set_random_state(9)
    out = {"image": torch.empty(6, 3, *crop)}
    for trial in range(4):
        params = aug.draw(6)
        ch, _ = aug.chains(samples, out, params)
        assert ch is not None
        plan = aug.plan(samples, params)
        ptr, stride = aug._dst_of(out["image"])
        _same(_native_items(ch, [18]), _plan_items(plan, ptr.reshape(-1), stride.reshape(-1, 3), [18]))
