"""GPU: RandCropByPosNegLabeld with the crop centres selected on the device (reference
transform_factory/augmentations.py:147-158, transforms.py:196-203).  The foreground / background index lists stay on
the device, the host makes MONAI's two draws per crop from their lengths, `adell_posneg_starts` looks the entry up,
corrects the centre and writes the crop start where K1 reads it (ADELL_F_WIN_DEV)."""

import numpy as np
import pytest
import torch

from adell_mri_b200 import collate, transform_factory as F, transforms as T
from adell_mri_b200.plan import BatchPlan
from oracle import monai_restated as M
from tests.helpers import mismatch, rand_affine_matrix, run_plan_cuda

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("shape,size", [((40, 36, 20), (16, 16, 8)), ((33, 17, 9), (32, 17, 4)), ((24, 24, 24), (30, 8, 24)), ((64, 64, 16), (35, 35, 17))])
def test_select_kernel_equals_the_restated_centre_correction(shape, size):
    R = np.random.RandomState(sum(shape))
    lab = torch.from_numpy((R.rand(*shape) > 0.8).astype(np.float32))
    flat = (lab > 0).reshape(-1)
    fg, bg = torch.nonzero(flat).reshape(-1), torch.nonzero(~flat).reshape(-1)
    fg_d, bg_d = fg.to(DEV), bg.to(DEV)
    picks, want = [], []
    for i in range(64):
        lst_h, lst_d = (fg, fg_d) if i % 2 else (bg, bg_d)
        p = int(R.randint(len(lst_h)))
        picks.append((lst_d, p))
        # the oracle's centre correction for exactly this entry (one crop, the draw replaced by the known entry)
        class OneDraw:
            def rand(self_inner):
                return 0.0 if i % 2 else 1.0
            def randint(self_inner, n):
                return p
        centers, sz = M.pos_neg_crop_centers(OneDraw(), size, 1, 0.5, shape, fg.numpy(), bg.numpy())
        want.append([max(int(c) - int(s) // 2, 0) for c, s in zip(centers[0], sz)])
    got = T._posneg_starts(picks, shape, [min(s, d) for s, d in zip(size, shape)], torch.device(DEV)).cpu().numpy()
    assert np.array_equal(got, np.asarray(want, np.int32))


@pytest.mark.parametrize("padding", ["border", "reflection"])
@pytest.mark.parametrize("dtype", ["f32", "i16"])
def test_device_windows_through_every_k1_path(padding, dtype):
    """Items whose window start is read from device memory: resampled (staged), identity copies with flips (box
    copy, unaligned starts along the contiguous axis), strict / default — against the same chains with a static crop."""
    R = np.random.RandomState(4)
    shape, size = (96, 80, 48), (64, 48, 32)
    if dtype == "f32":
        img = torch.from_numpy(R.rand(1, *shape).astype(np.float32))
    else:
        img = torch.from_numpy(R.randint(-100, 3000, size=(1, *shape)).astype(np.int16))
    dev = img[0].to(DEV)
    for trial in range(10):
        start = [int(R.randint(s - z + 1)) for s, z in zip(shape, size)]
        win = torch.tensor([start], dtype=torch.int32, device=DEV)
        A = rand_affine_matrix(R, rotate=(0.35, 0.35, 0.2), translate=(3, 3, 1), scale=(0.08, 0.08, 0.04))
        flips = np.array([R.rand() < 0.5 for _ in range(3)])
        for mode in ("bilinear", "nearest"):
            for strict in (True, False):
                dyn = BatchPlan([dev], strict=strict).crop_from_device(win, size).affine(A.numpy(), mode, padding).flip(flips).center_crop((56, 40, 28))
                sta = BatchPlan([dev], strict=strict).crop(start, size).affine(A.numpy(), mode, padding).flip(flips).center_crop((56, 40, 28))
                a, b = run_plan_cuda(dyn)[0], run_plan_cuda(sta)[0]
                if strict or mode == "nearest":
                    assert torch.equal(a, b), (trial, mode, strict)
                else:
                    assert torch.allclose(a, b, rtol=1e-5, atol=1e-5 * float(b.abs().max()))
            ref = M.center_spatial_crop(M.flip(M.affine_resample(M.crop(img, start, size), A, mode, padding),
                                               [i for i in range(3) if flips[i]]) if flips.any() else
                                        M.affine_resample(M.crop(img, start, size), A, mode, padding), (56, 40, 28))[0]
            assert mismatch(run_plan_cuda(BatchPlan([dev], strict=True).crop_from_device(win, size).affine(A.numpy(), mode, padding)
                                          .flip(flips).center_crop((56, 40, 28)))[0].cpu(), ref) == 0
        ident = BatchPlan([dev]).crop_from_device(win, size).flip(flips).center_crop((56, 40, 28))
        ref = M.crop(img, start, size).to(torch.float32)
        ref = M.center_spatial_crop(M.flip(ref, [i for i in range(3) if flips[i]]) if flips.any() else ref, (56, 40, 28))[0]
        assert mismatch(run_plan_cuda(ident)[0].cpu(), ref) == 0, (trial, start, flips)


def test_zeros_padding_on_a_device_window_is_refused():
    img = torch.rand(32, 32, 16, device=DEV)
    win = torch.zeros((1, 3), dtype=torch.int32, device=DEV)
    plan = BatchPlan([img]).crop_from_device(win, (16, 16, 8)).affine(np.eye(4, dtype=np.float32), "bilinear", "zeros")
    with pytest.raises(RuntimeError):
        run_plan_cuda(plan)


@pytest.mark.parametrize("n_crops", [1, 3])
def test_unet_crop_sandwich_device_selection_equals_host_selection(n_crops):
    """get_augmentations_unet(random_crop_size, has_label=True): FgBgToIndicesd with device lists + device-side
    centres against the host lists, same seeds: identical batches (strict mode: bit for bit)."""
    R = np.random.RandomState(2)
    keys, shape, rc = ["t2", "adc"], (72, 64, 40), [40, 36, 24]
    samples = []
    for _ in range(4):
        s = {k: torch.from_numpy(R.rand(1, *shape).astype(np.float32)).to(DEV) for k in keys}
        s["mask"] = torch.from_numpy((R.rand(1, *shape) > 0.85).astype(np.float32)).to(DEV)
        samples.append(s)
    T.set_mode(strict=True)
    try:
        outs = []
        for on_device in (True, False):
            fgbg = T.FgBgToIndicesd("mask", on_device=on_device)
            aug = F.get_augmentations_unet(["affine", "flip"], keys + ["mask"], keys, [], random_crop_size=rc, has_label=True,
                                           n_crops=n_crops, flip_axis=[0, 1, 2]).set_random_state(11)
            tf = F.SegmentationTransforms(keys + ["mask"], keys, ["mask"], keys, [])
            post = T.Compose(tf.post_transforms())
            batch = []
            for s in samples:
                crops = aug(fgbg(dict(s)))
                batch.append([post(c) for c in crops])
            outs.append(collate.safe_collate_crops(batch))
        for k in ("image", "mask"):
            assert outs[0][k].shape == outs[1][k].shape and outs[0][k].shape[0] == 4 * n_crops
            assert torch.equal(outs[0][k], outs[1][k]), k
    finally:
        T.set_mode(strict=False)


def test_batch_fast_path_of_the_label_guided_sandwich_equals_the_dictionary_surface():
    """SegmentationBatchAugmenter(random_crop_size, has_label=True, n_crops): the vectorised draws consume the crop
    transform's / samplers' / flips' streams like one pipeline call per sample does, the centres are selected on the
    device, and the collated [B * n_crops] batch equals safe_collate_crops of the dictionary surface bit for bit; several
    steps prepared at once equal step-by-step calls."""
    from adell_mri_b200 import collate, transform_factory as F, transforms as T
    from adell_mri_b200.pipelines import SegmentationBatchAugmenter

    R = np.random.RandomState(12)
    shape, rc, n_crops = (56, 48, 20), [24, 24, 10], 2
    image_keys = ["t2", "adc"]
    keys = image_keys + ["mask"]
    fgbg = T.FgBgToIndicesd("mask")
    cache = []
    for _ in range(6):
        s = {k: torch.from_numpy(R.rand(1, *shape).astype(np.float32)).to(DEV) for k in image_keys}
        s["mask"] = torch.from_numpy((R.rand(1, *shape) > 0.8).astype(np.float32)).to(DEV)
        cache.append(fgbg(s))
    T.set_mode(strict=True, fast=False, noise="injected")
    try:
        aug = F.get_augmentations_unet(["affine", "flip"], keys, image_keys, [], random_crop_size=rc, has_label=True, n_crops=n_crops,
                                       flip_axis=[0, 1, 2])
        tf = F.SegmentationTransforms(keys, image_keys, ["mask"], image_keys, [])
        for t in aug.transforms[1].transforms:          # make the affine fire often enough to matter in a small batch
            if isinstance(t, T.RandAffined):
                t.prob = t.sampler.prob = 0.6
        pipe = T.Compose([aug, *tf.post_transforms()]).set_random_state(41)
        want = collate.safe_collate_crops([pipe(dict(s)) for s in cache])
    finally:
        T.set_mode(strict=False)
    bat = SegmentationBatchAugmenter(["affine", "flip"], keys, image_keys, random_crop_size=rc, has_label=True, flip_axis=[0, 1, 2],
                                     n_crops=n_crops, strict=True)
    for smp in bat.samplers:
        smp.prob = 0.6
    got = bat.set_random_state(41, nested=True)(cache)
    assert got["image"].shape == (6 * n_crops, 2, *rc) and got["mask"].shape == (6 * n_crops, 1, *rc)
    assert torch.equal(got["image"], want["image"]) and torch.equal(got["mask"], want["mask"])
    # three steps of two samples prepared at once == the same six samples in one call (same seed)
    outs = [bat._alloc_out(2 * n_crops, tuple(rc), torch.device(DEV)) for _ in range(3)]
    steps = bat.set_random_state(41, nested=True).prepare_steps([cache[0:2], cache[2:4], cache[4:6]], outs)
    for k in range(3):
        steps.run(k)
    torch.cuda.synchronize()
    assert torch.equal(torch.cat([o["image"] for o in outs]), want["image"])
    assert torch.equal(torch.cat([o["mask"] for o in outs]), want["mask"])
