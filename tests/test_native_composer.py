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
        it = engine.aligned_items(n, items[start:start + n])
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


# ----------------------------------------------------------------------------- multi-pass sequences (adell_seq)
def _ssl_setup(seed, B, fast=False, strict=False, members=None, nk=1, shape=(40, 36, 20), roi=(32, 32, 16), choice="vectorised"):
    from adell_mri_b200.pipelines import SSLBatchAugmenter

    R = np.random.RandomState(seed)
    keys = ["image"] if nk == 1 else ["image", "aux"]
    samples = [{k: torch.from_numpy(R.rand(1, *shape).astype(np.float32)) for k in keys} for _ in range(B)]
    aug = SSLBatchAugmenter(keys, roi, n_transforms=3, choice=choice, noise="philox", fast=fast, strict=strict,
                            members=members).set_random_state(seed + 1)
    return aug, samples, shape, roi


@pytest.mark.parametrize("fast,strict,nk", [(False, False, 1), (False, True, 1), (True, False, 1), (False, False, 2)])
def test_ssl_sequences_equal_batchplan_launches(fast, strict, nk):
    """adell_seq_prepare_steps == BatchPlan.build_launches for the two-view SSL stream: the same launches in the same
    order, every item byte for byte (scratch addresses included), over several drawn batches."""
    B = 10
    aug, samples, shape, roi = _ssl_setup(11, B, fast, strict, nk=nk)
    out = {k: torch.empty(B, nk, *roi) for k in ("augmented_image_1", "augmented_image_2")}
    depth = set()
    for trial in range(5):
        params = aug.draw(B, shape, nk)
        plan, _ = aug.plan(samples, params)
        holder = {}

        def alloc(n):
            holder["buf"] = torch.empty(max(n, 1))
            return holder["buf"]

        p1, s1 = aug._dst_of(out["augmented_image_1"])
        p2, s2 = aug._dst_of(out["augmented_image_2"])
        ptr = np.stack([p1, p2], axis=1).reshape(-1)
        stride = np.stack([s1, s2], axis=1).reshape(-1, 3)
        want = plan.build_launches(ptr, stride, alloc)
        sq, _ = aug.seqs([samples], [out], params)
        got, scratch_used = engine.compose_seqs_host(sq, [len(sq)], mode=2,
                                                     scratch_ptr=holder["buf"].data_ptr() if "buf" in holder else 0)
        assert len(got) == len(want)
        depth.add(len(want))
        if "buf" in holder:
            assert scratch_used <= holder["buf"].numel()
        for (step, it), w in zip(got, want):
            assert step == 0 and it.shape == w.shape
            if fast:   # the composed matrices come from a float64 product: numpy's BLAS may sum in another order
                assert np.allclose(it["A"], w["A"], rtol=0, atol=1e-6)
                it = it.copy(); it["A"] = w["A"]
            assert it.tobytes() == w.tobytes()
    assert max(depth) >= 2   # the draws did exercise closed passes


def test_ssl_sequences_several_steps_equal_step_by_step():
    """One native call over three steps == three single-step calls, with every step's scratch volumes starting at the
    scratch base again; and with the reference's member choice (global numpy stream, sample by sample) one draw over the
    samples of three steps equals three consecutive draws (every stream is consumed in sample order)."""
    B = 6
    aug, samples, shape, roi = _ssl_setup(21, 3 * B, choice="global")
    np.random.seed(77)
    outs = [{k: torch.empty(B, 1, *roi) for k in ("augmented_image_1", "augmented_image_2")} for _ in range(3)]
    batches = [samples[0:B], samples[B:2 * B], samples[2 * B:]]
    sq, params = aug.seqs(batches, outs)
    n = 2 * B
    got, _ = engine.compose_seqs_host(sq, [n] * 3, mode=1, scratch_ptr=4096)
    for k in range(3):
        one, _ = engine.compose_seqs_host(np.ascontiguousarray(sq[k * n:(k + 1) * n]), [n], mode=1, scratch_ptr=4096)
        mine = [it for step, it in got if step == k]
        assert len(mine) == len(one)
        for a, (_, b) in zip(mine, one):
            assert a.tobytes() == b.tobytes()
    # and the draws of the big batch are those of three consecutive small ones
    aug2, _, _, _ = _ssl_setup(21, 3 * B, choice="global")
    np.random.seed(77)
    for k in range(3):
        p = aug2.draw(B, shape, 1)
        assert np.array_equal(p["starts"], params["starts"][:, k * B:(k + 1) * B])
        assert np.array_equal(p["choice"], params["choice"][:, k * B:(k + 1) * B])
        for v in range(2):
            for m, (use, vals) in p["draws"][v].items():
                big_use, big_vals = params["draws"][v][m]
                sel = (big_use >= k * B) & (big_use < (k + 1) * B)
                assert np.array_equal(big_use[sel] - k * B, use)
                if isinstance(vals, np.ndarray):
                    assert np.array_equal(np.asarray(big_vals)[sel], vals)
                else:
                    assert [big_vals[i] for i in np.nonzero(sel)[0]] == list(vals)


def test_seq_prepare_reports_missing_space():
    aug, samples, shape, roi = _ssl_setup(31, 8)
    out = {k: torch.empty(8, 1, *roi) for k in ("augmented_image_1", "augmented_image_2")}
    sq, _ = aug.seqs([samples], [out])
    launches = (_lib.SeqLaunch * 16)()
    buf = engine.aligned_bytes(1 << 20)
    st, nl, used, sused = engine._seq_call(sq, (len(sq),), 4096, 0, buf.ctypes.data, buf.size, launches, 2)
    if sused > 0:
        assert st == _lib.ERR_NO_SPACE
    st, nl, used, sused2 = engine._seq_call(sq, (len(sq),), 4096, sused, buf.ctypes.data, 256, launches, 2)
    assert st == _lib.ERR_NO_SPACE and sused2 == sused and used > 256
    st, nl, used, _ = engine._seq_call(sq, (len(sq),), 4096, sused, buf.ctypes.data, buf.size, launches, 2)
    assert st == 0 and nl >= 1


def test_label_guided_sandwich_chains_equal_batchplan_items(monkeypatch):
    """The crop sandwich with device-side windows (RandCropByPosNegLabeld on the batch path): adell_chain.win_dev ==
    BatchPlan.crop_from_device, byte for byte (the select kernel itself is replaced by a host tensor here: only the
    addresses of the window rows enter the items)."""
    from adell_mri_b200 import transforms as T

    monkeypatch.setattr(T, "_posneg_starts", lambda picks, shape, size, dev: torch.zeros((len(picks), 3), dtype=torch.int32))
    R = np.random.RandomState(6)
    keys, shape, rc, nc = ["t2", "adc"], (36, 40, 12), [20, 24, 8], 2
    samples = _samples(R, 6, keys, shape)
    for s in samples:
        flat = (s["mask"] > 0).reshape(-1)
        s["mask_fg_indices"], s["mask_bg_indices"] = torch.nonzero(flat).reshape(-1), torch.nonzero(~flat).reshape(-1)
    aug = SegmentationBatchAugmenter(["affine", "flip"], keys + ["mask"], keys, random_crop_size=rc, has_label=True, flip_axis=[0, 1, 2],
                                     n_crops=nc).set_random_state(2)
    for smp in aug.samplers:
        smp.prob = 0.5
    outs = [aug._alloc_out(2 * nc, tuple(rc), torch.device("cpu")) for _ in range(3)]
    batches = [samples[0:2], samples[2:4], samples[4:6]]
    ch, params = aug.chains(batches, outs)
    assert ch is not None and (ch["win_dev"] != 0).all()
    plan = aug.plan(samples, params)
    assert not plan.passes
    ptrs, strides = zip(*[aug._dst(o, 2 * nc) for o in outs])
    nk = len(aug.keys)
    sizes = [2 * nc * nk] * 3
    _same(_native_items(ch, sizes), _plan_items(plan, np.concatenate(ptrs), np.concatenate(strides), sizes))
