"""Batch fast paths (adell_mri_b200.pipelines.ClassificationBatchAugmenter / SSLBatchAugmenter)
against the dictionary-transform surface on identical seeds: the vectorised draws must consume the
very same RandomState streams as one transform call per sample does.  (The dictionary surface is
itself pinned against the eager oracle pipelines in tests/test_lazy_pipelines.py.)  CPU: plans run
through the C restatement; GPU (marked): through CUDA."""

import numpy as np
import pytest
import torch

from adell_mri_b200 import collate, engine, transform_factory as F, transforms as T
from adell_mri_b200.pipelines import SSL_FUSED_MEMBERS, ClassificationBatchAugmenter, SSLBatchAugmenter
from oracle import cref


def _cref_execute_ptrs(plan, dst_ptr, dst_stride, keep=None):
    launches = plan.build_launches(dst_ptr, dst_stride, lambda n: torch.empty(max(n, 1), dtype=torch.float32))
    for items in launches:
        cref.gather(items)


@pytest.fixture(params=["cpu", pytest.param("cuda:0", marks=pytest.mark.gpu)])
def dev(request, monkeypatch):
    if request.param == "cpu":
        from tests.helpers import cref_execute, cref_prepare_chain_steps, cref_prepare_seq_steps
        monkeypatch.setattr(engine, "prepare_seq_steps", cref_prepare_seq_steps)
        monkeypatch.setattr(engine, "execute", cref_execute)
        monkeypatch.setattr(engine, "execute_ptrs", _cref_execute_ptrs)
        monkeypatch.setattr(engine, "prepare_chain_steps", cref_prepare_chain_steps)
    T.set_mode(strict=True, fast=False, noise="injected")
    yield request.param
    T.set_mode(strict=False)


def _apply(chain, d):
    for t in chain:
        d = t(d)
    return d


def _samples(R, n, keys, shape, dev, mask=True):
    out = []
    for _ in range(n):
        s = {k: torch.from_numpy(R.rand(1, *shape).astype(np.float32)).to(dev) for k in keys}
        if mask:
            s["mask"] = torch.from_numpy((R.rand(1, *shape) > 0.7).astype(np.float32)).to(dev)
        out.append(s)
    return out


@pytest.mark.parametrize("augment", [["flip", "affine"], ["flip", "affine", "shear"], ["affine"]])
def test_classification_batch_equals_dictionary_surface(dev, augment):
    R = np.random.RandomState(3)
    keys, shape, crop = ["t2", "adc"], (40, 36, 24), [24, 20, 8]
    samples = _samples(R, 7, keys, shape, dev)
    # (an outer Compose would re-seed its children on construction, like MONAI's: chain by hand)
    chain = [F.get_augmentations_class(augment, keys, "mask", [], flip_axis=[0, 1, 2], prob=0.6).set_random_state(17),
             T.CenterSpatialCropd(keys + ["mask"], crop), T.ConcatItemsd(keys + ["mask"], "image")]
    want = collate.safe_collate([_apply(chain, dict(s)) for s in samples])["image"]
    aug = ClassificationBatchAugmenter(augment, keys, "mask", flip_axis=[0, 1, 2], prob=0.6, crop_size=crop, strict=True)
    got = aug.set_random_state(17)(samples)["image"]
    assert got.shape == (7, 3, *crop)
    assert torch.equal(got.cpu(), want.cpu())


@pytest.mark.parametrize("different_crop,vicregl", [(False, False), (True, False), (False, True)])
def test_ssl_batch_equals_dictionary_surface(dev, different_crop, vicregl):
    R = np.random.RandomState(5)
    keys, copied, shape, roi = ["image"], ["image_copy"], (36, 32, 16), [24, 24, 12]
    samples = _samples(R, 6, keys, shape, dev, mask=False)
    tf = F.SSLTransforms(keys, copied, adc_keys=[], non_adc_keys=[])
    chain = [tf.pre_transforms()[-1],   # CopyEntryd
             # the batch fast path holds the single-launch members (the power law / Rician members need a pass of their own)
             T.Compose(F.get_augmentations_ssl(keys, copied, None, roi, vicregl, different_crop, n_transforms=3,
                                               aug_list=list(SSL_FUSED_MEMBERS))).set_random_state(23),
             *tf.post_transforms()]
    np.random.seed(77)
    want = collate.safe_collate([_apply(chain, dict(s)) for s in samples])
    aug = SSLBatchAugmenter(keys, roi, n_transforms=3, different_crop=different_crop, vicregl=vicregl, strict=True)
    np.random.seed(77)
    got = aug.set_random_state(23)(samples)
    for k in ("augmented_image_1", "augmented_image_2"):
        assert got[k].shape == (6, 1, *roi)
        assert torch.allclose(got[k].cpu(), want[k].cpu(), rtol=2e-6, atol=2e-6), (k, float((got[k].cpu() - want[k].cpu()).abs().max()))
    if vicregl:
        for k in ("box_1", "box_2"):
            assert np.array_equal(np.asarray(got[k]), np.asarray(want[k]))


@pytest.mark.parametrize("fast,vicregl", [(False, False), (True, False), (False, True)])
def test_ssl_native_sequences_equal_batchplan_route(dev, fast, vicregl):
    """The native sequence route (adell_seq: several steps composed by one call) writes the very voxels of the numpy
    BatchPlan route on the same draws — device Philox noise included."""
    R = np.random.RandomState(8)
    shape, roi = (36, 32, 16), [24, 24, 12]
    samples = _samples(R, 8, ["image"], shape, dev, mask=False)
    # (the C restatement of the CPU leg has no Philox generator: the noise member is exercised on the GPU leg)
    members = None if dev != "cpu" else [m for m in SSL_FUSED_MEMBERS if m != "gaussian_noise"]
    aug = SSLBatchAugmenter(["image"], roi, n_transforms=(2 if vicregl and dev == "cpu" else 3), vicregl=vicregl, choice="vectorised",
                            noise="philox", fast=fast, members=members).set_random_state(4)
    assert aug.native_ok()
    params = aug.draw(8, shape, 1)
    want = aug(samples, params=params, native=False)
    got = aug(samples, params=params, native=True)
    for k in ("augmented_image_1", "augmented_image_2"):
        assert torch.equal(got[k].cpu(), want[k].cpu()), k
    # two steps prepared at once == the same samples as one batch
    outs = [{k: torch.empty(4, 1, *roi, device=dev) for k in ("augmented_image_1", "augmented_image_2")} for _ in range(2)]
    steps = aug.prepare_steps([samples[:4], samples[4:]], outs, params)
    assert len(steps) == 2
    for k in range(2):
        steps.run(k)
    for k in ("augmented_image_1", "augmented_image_2"):
        assert torch.equal(torch.cat([outs[0][k], outs[1][k]]).cpu(), want[k].cpu()), k
