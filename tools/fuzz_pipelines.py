"""Randomised sweep of the lazy dictionary-transform surface (adell_mri_b200.transforms / transform_factory /
collate) against the eager oracle restatement of the reference's pipelines (oracle/pipelines_ref.py): random
seeds, shapes, key sets, flip axes, probabilities, crop sandwiches and augment lists for the unet, classification
and SSL builders.  Runs on the CPU: plans execute through the C restatement (oracle/gather_ref.c), so this sweeps
the HOST logic — draw orders, seed fan-out, chain composition, collation — not the CUDA kernels.  Test
infrastructure (imports oracle/).

    python tools/fuzz_pipelines.py [n_rounds] [seed] [cuda:0]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from adell_mri_b200 import collate, engine, transform_factory as F, transforms as T
from oracle import pipelines_ref as P
from adell_mri_b200.pipelines import SSL_FUSED_MEMBERS, ClassificationBatchAugmenter, SegmentationBatchAugmenter, SSLBatchAugmenter
from oracle import cref
from tests.helpers import cref_execute, cref_prepare_chain_steps


def cref_execute_ptrs(plan, dst_ptr, dst_stride, keep=None):
    launches = plan.build_launches(dst_ptr, dst_stride, lambda n: torch.empty(max(n, 1), dtype=torch.float32))
    for items in launches:
        cref.gather(items)



DEVICE = "cpu"


def samples_of(R, n, keys, shape, mask=True):
    """CPU samples (what the eager oracle consumes); `dev()` moves a copy to the device under test."""
    out = []
    for _ in range(n):
        s = {k: torch.from_numpy(R.rand(1, *shape).astype(np.float32)) for k in keys}
        if mask:
            s["mask"] = torch.from_numpy((R.rand(1, *shape) > 0.7).astype(np.float32))
        out.append(s)
    return out


def dev(s):
    return {k: (v.to(DEVICE) if isinstance(v, torch.Tensor) else v) for k, v in s.items()}


def host(d):
    return {k: (v.cpu() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}


def shape_of(R):
    return (int(R.choice([16, 20, 24, 28, 33])), int(R.choice([16, 18, 24, 30])), int(R.choice([8, 12, 16])))


def unet_round(R):
    keys = list(R.choice(["t2", "adc", "dwi", "hbv"], size=int(R.randint(1, 4)), replace=False))
    shape = shape_of(R)
    augment = [a for a in ["affine", "shear", "flip"] if R.rand() < 0.7] or ["flip"]
    if R.rand() < 0.3:
        augment = ["trivial"] + augment
    flip_axis = sorted(int(a) for a in R.choice([0, 1, 2], size=int(R.randint(1, 4)), replace=False))
    seed = int(R.randint(1 << 30))
    has_label = bool(R.rand() < 0.7)
    rc = None
    if R.rand() < 0.4 and not has_label:
        rc = [int(s * R.uniform(0.5, 0.8)) for s in shape]
    all_keys = keys + (["mask"] if has_label else [])
    samples = samples_of(R, 3, keys, shape, mask=has_label)
    lazy = T.Compose([F.get_augmentations_unet(augment, all_keys, keys, [], random_crop_size=rc, has_label=has_label, flip_axis=flip_axis),
                      T.ConcatItemsd(keys, "image")]).set_random_state(seed)
    ref = P.Chain([P.unet(augment, all_keys, keys, random_crop_size=rc, has_label=has_label, flip_axis=tuple(flip_axis)),
                   P.ConcatD(keys, "image")]).seed(seed)
    got = host(collate.safe_collate([lazy(dev(s)) for s in samples]))
    desc = f"unet augment={augment} keys={keys} shape={shape} flip_axis={flip_axis} rc={rc} label={has_label} seed={seed}"
    bad = 0
    for b, s in enumerate(samples):
        w = ref(s)
        bad += int(not torch.equal(got["image"][b], w["image"]))
        if has_label:
            bad += int(not torch.equal(got["mask"][b], w["mask"].to(torch.float32)))
    return bad, desc


def class_round(R):
    keys = list(R.choice(["t2", "adc", "dwi"], size=int(R.randint(1, 4)), replace=False))
    shape = (int(R.choice([32, 36, 40])), int(R.choice([32, 40])), int(R.choice([20, 24])))
    crop = [shape[0] - 16 - int(R.choice([0, 2])), shape[1] - 16, shape[2] - 16]
    augment = [a for a in ["flip", "affine", "shear"] if R.rand() < 0.7] or ["affine"]
    if R.rand() < 0.3:
        augment = ["trivial"] + augment
    flip_axis = sorted(int(a) for a in R.choice([0, 1, 2], size=int(R.randint(1, 4)), replace=False))
    prob = float(R.choice([0.1, 0.5, 0.9]))
    seed = int(R.randint(1 << 30))
    samples = samples_of(R, 3, keys, shape)
    tf = F.ClassificationTransforms(keys, adc_keys=[], crop_size=crop, mask_key="mask")
    lazy = T.Compose([*tf.pre_transforms()[-2:], F.get_augmentations_class(augment, keys, "mask", [], flip_axis=flip_axis, prob=prob),
                      *tf.post_transforms()]).set_random_state(seed)
    m = [c + 16 for c in crop]
    ref = P.Chain([P.CenterCropD(keys + ["mask"], m), P.classification(augment, keys, "mask", flip_axis=tuple(flip_axis), prob=prob),
                   P.CenterCropD(keys + ["mask"], crop), P.ConcatD(keys + ["mask"], "image")]).seed(seed)
    got = host(collate.safe_collate([lazy(dev(s)) for s in samples]))
    desc = f"class augment={augment} keys={keys} shape={shape} crop={crop} flip_axis={flip_axis} prob={prob} seed={seed}"
    bad = sum(int(not torch.equal(got["image"][b], ref(s)["image"])) for b, s in enumerate(samples))
    return bad, desc


def ssl_round(R):
    shape = (int(R.choice([28, 32, 36])), int(R.choice([28, 32])), int(R.choice([12, 16])))
    roi = [shape[0] - int(R.choice([4, 8])), shape[1] - int(R.choice([4, 8])), shape[2] - int(R.choice([0, 4]))]
    vicregl, different = bool(R.rand() < 0.3), bool(R.rand() < 0.4)
    n_t = int(R.randint(1, 4))
    names = [m for m in F.FUSED_AUGMENTS if m not in ("contrast", "rician_noise")]
    names = [m for m in names if R.rand() < 0.8] or ["rotate_z", "shift_intensity", "gaussian_noise"]
    if vicregl:   # the spatial members are dropped by the builder: keep enough of the others
        names = ["gaussian_noise", "shift_intensity", "scale_intensity"] + [m for m in names if m not in ("gaussian_noise", "shift_intensity", "scale_intensity")]
    while len(names) < n_t:
        names.append([m for m in F.FUSED_AUGMENTS if m not in names and m not in ("contrast", "rician_noise")][0])
    seed, gseed = int(R.randint(1 << 30)), int(R.randint(1 << 30))
    samples = samples_of(R, 3, ["image"], shape, mask=False)
    tf = F.SSLTransforms(["image"], ["image_copy"], adc_keys=[], non_adc_keys=[])
    lazy = tf.transforms(F.get_augmentations_ssl(["image"], ["image_copy"], None, roi, vicregl, different, n_transforms=n_t,
                                                 aug_list=list(names))).set_random_state(seed)
    ref = P.Chain(P.ssl(["image"], ["image_copy"], roi, vicregl, different, names, n_t)).seed(seed)
    np.random.seed(gseed)
    got = host(collate.safe_collate([lazy(dev(s)) for s in samples]))
    np.random.seed(gseed)
    desc = f"ssl names={names} shape={shape} roi={roi} vicregl={vicregl} different={different} N={n_t} seed={seed}"
    bad = 0
    for b, s in enumerate(samples):
        d = dict(s); d["image_copy"] = s["image"].clone()
        w = ref(d)
        for gk, wk in (("augmented_image_1", "image"), ("augmented_image_2", "image_copy")):
            bad += int(not torch.allclose(got[gk][b], w[wk].to(torch.float32), rtol=2e-6, atol=2e-6))
    return bad, desc


def seg_batch_round(R):
    """The batch fast path (what bench.py times) against the dictionary surface on the same seed: identical
    streams drawn batch-wise, one plan for the whole batch."""
    keys = list(R.choice(["t2", "adc", "dwi", "hbv"], size=int(R.randint(1, 4)), replace=False))
    shape = shape_of(R)
    augment = [a for a in ["affine", "shear", "flip"] if R.rand() < 0.7] or ["flip"]
    flip_axis = sorted(int(a) for a in R.choice([0, 1, 2], size=int(R.randint(1, 4)), replace=False))
    has_label = bool(R.rand() < 0.7)
    rc = [int(s * R.uniform(0.5, 0.8)) for s in shape] if (R.rand() < 0.4 and not has_label) else None
    seed = int(R.randint(1 << 30))
    all_keys = keys + (["mask"] if has_label else [])
    samples = samples_of(R, 5, keys, shape, mask=has_label)
    # the augmenter's seed is the seed of the augmentation Compose itself (the reference nests that Compose inside
    # the pipeline's outer Compose, which hands it its own child seed: `nested=True` below)
    nested = bool(R.rand() < 0.5)
    aug_t = F.get_augmentations_unet(augment, all_keys, keys, [], random_crop_size=rc, has_label=has_label, flip_axis=flip_axis)
    concat = T.ConcatItemsd(keys, "image")
    if nested:
        lazy = T.Compose([aug_t, concat]).set_random_state(seed)
    else:   # chained by hand: constructing an outer Compose would re-seed its children (MONAI behaviour)
        aug_t.set_random_state(seed)
        lazy = lambda d: concat(aug_t(d))   # noqa: E731
    want = host(collate.safe_collate([lazy(dev(s)) for s in samples]))
    aug = SegmentationBatchAugmenter(augment, all_keys, keys, random_crop_size=rc, has_label=has_label, flip_axis=flip_axis, strict=True)
    got = host(aug.set_random_state(seed, nested=nested)([dev(s) for s in samples]))
    desc = f"seg_batch nested={nested} augment={augment} keys={keys} shape={shape} flip_axis={flip_axis} rc={rc} label={has_label} seed={seed}"
    bad = int(not torch.equal(got["image"], want["image"]))
    if has_label:
        bad += int(not torch.equal(got["mask"], want["mask"]))
    return bad, desc


def class_batch_round(R):
    keys = list(R.choice(["t2", "adc", "dwi"], size=int(R.randint(1, 4)), replace=False))
    shape = (int(R.choice([32, 36, 40])), int(R.choice([32, 40])), int(R.choice([20, 24])))
    crop = [shape[0] - 16, shape[1] - 16 - int(R.choice([0, 4])), shape[2] - 16]
    augment = [a for a in ["flip", "affine", "shear"] if R.rand() < 0.7] or ["affine"]
    flip_axis = sorted(int(a) for a in R.choice([0, 1, 2], size=int(R.randint(1, 4)), replace=False))
    prob, seed, nested = float(R.choice([0.1, 0.5, 0.9])), int(R.randint(1 << 30)), bool(R.rand() < 0.5)
    mask_key = "mask" if R.rand() < 0.7 else None
    samples = samples_of(R, 5, keys, shape, mask=mask_key is not None)
    all_keys = keys + ([mask_key] if mask_key else [])
    aug_t = F.get_augmentations_class(augment, keys, mask_key, [], flip_axis=flip_axis, prob=prob)
    tail = [T.CenterSpatialCropd(all_keys, crop), T.ConcatItemsd(all_keys, "image")]
    if nested:
        lazy = T.Compose([aug_t, *tail]).set_random_state(seed)
    else:
        aug_t.set_random_state(seed)
        lazy = lambda d: tail[1](tail[0](aug_t(d)))   # noqa: E731
    want = host(collate.safe_collate([lazy(dev(s)) for s in samples]))["image"]
    aug = ClassificationBatchAugmenter(augment, keys, mask_key, flip_axis=flip_axis, prob=prob, crop_size=crop, strict=True)
    got = aug.set_random_state(seed, nested=nested)([dev(s) for s in samples])["image"].cpu()
    desc = f"class_batch nested={nested} augment={augment} keys={keys} mask={mask_key} shape={shape} crop={crop} flip_axis={flip_axis} prob={prob} seed={seed}"
    return int(not torch.equal(got, want)), desc


def ssl_batch_round(R):
    shape = (int(R.choice([28, 32, 36])), int(R.choice([28, 32])), int(R.choice([12, 16])))
    roi = [shape[0] - int(R.choice([4, 8])), shape[1] - int(R.choice([4, 8])), shape[2] - int(R.choice([0, 4]))]
    vicregl, different = bool(R.rand() < 0.3), bool(R.rand() < 0.4)
    n_t = int(R.randint(1, 4))
    members = [m for m in SSL_FUSED_MEMBERS if R.rand() < 0.8]
    for m in ("gaussian_noise", "shift_intensity", "scale_intensity"):
        if m not in members:
            members.insert(0, m)
    members = [m for m in SSL_FUSED_MEMBERS if m in members]    # the reference's list order
    seed, gseed = int(R.randint(1 << 30)), int(R.randint(1 << 30))
    samples = samples_of(R, 4, ["image"], shape, mask=False)
    tf = F.SSLTransforms(["image"], ["image_copy"], adc_keys=[], non_adc_keys=[])
    chain = [tf.pre_transforms()[-1],
             T.Compose(F.get_augmentations_ssl(["image"], ["image_copy"], None, roi, vicregl, different, n_transforms=n_t,
                                               aug_list=list(members))).set_random_state(seed),
             *tf.post_transforms()]

    def run(d):
        for t in chain:
            d = t(d)
        return d

    np.random.seed(gseed)
    want = host(collate.safe_collate([run(dev(s)) for s in samples]))
    aug = SSLBatchAugmenter(["image"], roi, n_transforms=n_t, different_crop=different, vicregl=vicregl, members=members, strict=True)
    np.random.seed(gseed)
    got = host(aug.set_random_state(seed)([dev(s) for s in samples]))
    desc = f"ssl_batch members={members} shape={shape} roi={roi} vicregl={vicregl} different={different} N={n_t} seed={seed}"
    bad = sum(int(not torch.allclose(got[k], want[k], rtol=2e-6, atol=2e-6)) for k in ("augmented_image_1", "augmented_image_2"))
    if vicregl:
        bad += sum(int(not np.array_equal(np.asarray(got[k]), np.asarray(want[k]))) for k in ("box_1", "box_2"))
    return bad, desc


def posneg_round(R):
    """RandCropByPosNegLabeld sandwich (has_label, random_crop_size, n_crops): list outputs, fg / bg index lists
    from the cached stage, masks with few or no foreground voxels."""
    keys = list(R.choice(["t2", "adc", "dwi"], size=int(R.randint(1, 3)), replace=False))
    shape = (int(R.choice([24, 28, 32])), int(R.choice([24, 30])), int(R.choice([8, 12])))
    rc = [int(s * R.uniform(0.4, 0.8)) for s in shape]
    n_crops = int(R.randint(1, 4))
    augment = [a for a in ["affine", "flip"] if R.rand() < 0.7] or ["flip"]
    seed = int(R.randint(1 << 30))
    thr = float(R.choice([0.7, 0.97, 0.999, 0.02]))
    samples = samples_of(R, 3, keys, shape)
    for s in samples:
        s["mask"] = (torch.from_numpy(R.rand(1, *shape).astype(np.float32)) > thr).to(torch.float32)
        if not (s["mask"] > 0).any() or (s["mask"] > 0).all():   # MONAI needs both classes (it warns and degenerates otherwise)
            s["mask"][0, 1, 1, 1] = 1.0; s["mask"][0, 2, 2, 2] = 0.0
        flat = (s["mask"] > 0).reshape(-1).numpy()
        s["mask_fg_indices"], s["mask_bg_indices"] = np.nonzero(flat)[0], np.nonzero(~flat)[0]
    lazy = T.Compose([F.get_augmentations_unet(augment, keys + ["mask"], keys, [], random_crop_size=rc, n_crops=n_crops, flip_axis=[0, 1, 2]),
                      T.ConcatItemsd(keys, "image"), T.SelectItemsd(["image", "mask"])]).set_random_state(seed)
    ref = P.Chain([P.unet(augment, keys + ["mask"], keys, random_crop_size=rc, n_crops=n_crops, flip_axis=(0, 1, 2)), P.ConcatD(keys, "image")]).seed(seed)
    got = host(collate.safe_collate_crops([lazy(dev(s)) for s in samples]))
    want = [c for s in samples for c in ref(s)]
    desc = f"posneg augment={augment} keys={keys} shape={shape} rc={rc} n_crops={n_crops} thr={thr} seed={seed}"
    pre = [int(i * 1.10) for i in rc]
    if tuple(got["image"].shape) != (3 * n_crops, len(keys), *rc):
        return 1, desc + f" shape {tuple(got['image'].shape)}"
    bad = 0
    for b, w in enumerate(want):
        bad += int(not torch.equal(got["image"][b], w["image"])) + int(not torch.equal(got["mask"][b], w["mask"]))
    return bad, desc


def sweep(rounds, seed, verbose=True, device="cpu"):
    """device="cpu": plans execute through the C restatement (host logic only); "cuda:0": through the CUDA path."""
    global DEVICE
    DEVICE = device
    R = np.random.RandomState(seed)
    T.set_mode(strict=True, fast=False, noise="injected")
    saved, saved_ptrs, saved_chain = engine.execute, engine.execute_ptrs, engine.prepare_chain_steps
    if device == "cpu":     # CPU stand-ins for the launcher (test infrastructure)
        engine.execute, engine.execute_ptrs = cref_execute, cref_execute_ptrs
        engine.prepare_chain_steps = cref_prepare_chain_steps
    bad = 0
    try:
        for r in range(rounds):
            fn = [unet_round, class_round, ssl_round, seg_batch_round, class_batch_round, ssl_batch_round, posneg_round][r % 7]
            b, desc = fn(R)
            if b:
                bad += 1
                if verbose:
                    print(f"MISMATCH round {r}: {desc} ({b} tensors differ)")
    finally:
        T.set_mode(strict=False)
        engine.execute = saved
        engine.execute_ptrs = saved_ptrs
        engine.prepare_chain_steps = saved_chain
    return rounds, bad


if __name__ == "__main__":
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    n, bad = sweep(rounds, seed, device=sys.argv[3] if len(sys.argv) > 3 else "cpu")
    print(f"fuzz_pipelines: {n} rounds, {bad} with mismatches (seed {seed})")
    sys.exit(1 if bad else 0)
