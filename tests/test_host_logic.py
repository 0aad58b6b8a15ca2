"""CPU: host-side logic of the product — RNG stream replicas, batched geometry, the composer's
pass bookkeeping and the batch augmenter — checked against the oracle / sequential semantics."""

import numpy as np
import pytest
import torch

from adell_mri_b200 import geometry as G
from adell_mri_b200.pipelines import SegmentationBatchAugmenter
from adell_mri_b200.plan import BatchPlan
from adell_mri_b200.sampling import RandAffineSampler, child_seeds
from oracle import monai_restated as M
from tests.helpers import mismatch, rand_affine_matrix, run_plan_cref


def test_compose_seed_fanout_matches_oracle():
    assert child_seeds(42, 5) == M.compose_set_random_state(42, 5)


@pytest.mark.parametrize("idle", [False, True])
@pytest.mark.parametrize("prob", [0.2, 1.0])
def test_rand_affine_sampler_follows_the_restated_stream(prob, idle):
    kw = dict(rotate_range=[np.pi / 8, np.pi / 8, np.pi / 16], scale_range=[0.1, 0.1, 0.05])
    a = RandAffineSampler(prob=prob, per_key_draws_when_idle=idle, **kw).set_random_state(7)
    b = M.RandAffinedDraws(prob, n_keys=4, idle_per_key_draws=idle, **kw).set_random_state(7)
    for _ in range(20):
        fa, pa = a.draw(n_keys=4)
        fb, pb = b.draw()
        assert fa == fb and pa == pb


def test_idle_calls_consume_no_per_key_draws_by_default():
    """† MONAI >= 1.0: `self.rand_affine(d[key], ..., randomize=True, grid)` sits under `if do_resampling:`; an idle
    call advances the gate streams by one draw each and the grid stream by ONE randomize."""
    kw = dict(rotate_range=[0.3, 0.3, 0.1])
    s = RandAffineSampler(prob=0.0, **kw).set_random_state(5)
    for _ in range(6):
        assert s.draw(n_keys=4) == (False, None)
    r = np.random.RandomState(5)
    r.random_sample(6)
    assert s.R.rand() == r.rand() and s.R_inner.rand() == np.random.RandomState(5).random_sample(7)[-1]
    g = np.random.RandomState(5)
    g.random_sample(6 * 3)
    assert s.R_grid.rand() == g.rand()
    old = RandAffineSampler(prob=0.0, per_key_draws_when_idle=True, **kw).set_random_state(5)
    for _ in range(6):
        old.draw(n_keys=4)
    g = np.random.RandomState(5)
    g.random_sample(6 * 3 * 5)
    assert old.R_grid.rand() == g.rand()


def test_vectorised_draws_equal_sequential_draws():
    for trial in range(12):
        kw = dict(prob=[0.2, 0.5, 1.0][trial % 3], rotate_range=[0.3, 0.3, 0.1],
                  shear_range=((0.9, 1.1),) * 3 if trial % 2 else None,
                  translate_range=[4, 4, 1] if trial % 4 == 0 else None)
        kw["per_key_draws_when_idle"] = bool(trial % 2) if trial < 6 else None
        a = RandAffineSampler(**kw).set_random_state(trial)
        b = RandAffineSampler(**kw).set_random_state(trial)
        for _ in range(3):
            seq = [a.draw(4) for _ in range(9)]
            fired, p = b.draw_batch(9, 4)
            assert list(fired) == [s[0] for s in seq]
            j = 0
            for f, ps in seq:
                if f:
                    for k in p:
                        assert np.array_equal(np.asarray(ps[k], np.float64), p[k][j])
                    j += 1


def test_batched_affine_composition_is_bit_identical_to_monai_order():
    R = np.random.RandomState(0)
    for trial in range(40):
        B = 4
        rot, sh = R.uniform(-1, 1, (B, 3)), R.uniform(0.9, 1.1, (B, 3))
        tr, sc = R.uniform(-5, 5, (B, 3)), R.uniform(0.9, 1.1, (B, 3))
        args = dict(rotate=rot[:, : 1 + trial % 3], shear=sh if trial % 2 else None,
                    translate=tr if trial % 3 else None, scale=sc if trial % 5 else None)
        Ab = G.compose_affine(**args)
        for b in range(B):
            ref = M.compose_affine(*(None if v is None else list(v[b]) for v in
                                     (args["rotate"], args["shear"], args["translate"], args["scale"]))).numpy()
            assert np.array_equal(ref, Ab[b])


def test_fast_mode_composes_consecutive_affines():
    R = np.random.RandomState(3)
    img = torch.from_numpy(R.rand(16, 16, 8).astype(np.float32))
    A1, A2 = rand_affine_matrix(R).numpy(), rand_affine_matrix(R).numpy()
    plan = BatchPlan([img], fast=True).affine(A1, "bilinear", "zeros").affine(A2, "bilinear", "zeros")
    assert len(plan.passes) == 0
    assert np.allclose(plan.st.A[0], (A1.astype(np.float64) @ A2.astype(np.float64))[:3], atol=1e-6)
    plan = BatchPlan([img]).affine(A1, "bilinear", "zeros").affine(A2, "bilinear", "zeros")
    assert len(plan.passes) == 1


def test_pad_after_offset_closes_a_pass_and_stays_exact():
    R = np.random.RandomState(4)
    img = torch.from_numpy(R.rand(1, 10, 9, 8).astype(np.float32))
    ref = M.spatial_pad(img + torch.tensor(0.25), (14, 12, 10))[0]
    plan = BatchPlan([img[0]], strict=True).intensity(offset=0.25).spatial_pad((14, 12, 10))
    assert mismatch(run_plan_cref(plan)[0], ref) == 0


def test_segmentation_augmenter_matches_oracle_chain():
    keys = ["t2", "adc", "dwi", "mask"]
    R = np.random.RandomState(0)
    samples = [{k: torch.from_numpy(R.rand(1, 32, 28, 12).astype(np.float32)) for k in keys} for _ in range(6)]
    aug = SegmentationBatchAugmenter(["affine", "shear", "flip"], keys, keys[:3], flip_axis=[0, 1, 2], strict=True)
    for s in aug.samplers:
        s.prob = 0.6
    aug.set_random_state(42)
    params = aug.draw(6, (32, 28, 12))
    outs = run_plan_cref(aug.plan(samples, params))
    for b in range(6):
        fl = [a for a in range(3) if params["flips"][b, a]]
        for ki, k in enumerate(keys):
            ref = samples[b][k]
            for si in range(2):  # affine then shear: two sequential RandAffined firings
                if params["fired"][si, b]:
                    ref = M.affine_resample(ref, torch.from_numpy(params["mats"][si, b]), aug.modes[ki], "reflection")
            if fl:
                ref = M.flip(ref, fl)
            assert mismatch(outs[b * 4 + ki], ref[0].contiguous()) == 0


def test_unknown_or_out_of_scope_augment_raises():
    with pytest.raises(NotImplementedError):
        SegmentationBatchAugmenter(["sharpen"], ["a", "mask"], ["a"])
    with pytest.raises(NotImplementedError):
        SegmentationBatchAugmenter(["blur"], ["a", "mask"], ["a"])


def test_cpu_plans_are_refused_by_the_engine():
    from adell_mri_b200 import engine

    img = torch.zeros(4, 4, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        engine.execute(BatchPlan([img]), [torch.empty(4, 4, 4)])


def test_percentile_positions_follow_numpys_linear_method_to_the_bit():
    """`stats.numpy_virtual_index` against numpy itself: np.percentile of arange(n) (float64) IS the lerp of the
    two order statistics lo, lo + 1 with numpy's own gamma, so its bits pin (lo, hi, gamma).  numpy evaluates
    (n - 1) * q for 'linear'; the equivalent alpha = beta = 1 formula rounds differently (tools/fuzz_stats.py
    found a percentile one ulp off at a near-tie of the final float32 rounding)."""
    from adell_mri_b200 import stats

    R = np.random.RandomState(0)
    ns = [1, 2, 3, 7, 1000, 65536, 66538, 2_097_152, 33_554_432, 268_435_456 + 5]
    qs = [0.0, 0.5, 1.0, 2.0, 25.0, 50.0, 75.0, 98.0, 99.0, 99.5, 100.0] + [float(x) for x in R.uniform(0, 100, 40)]
    for n in ns:
        a = np.arange(min(n, 70_000), dtype=np.float64) if n <= 70_000 else None
        for q in qs:
            lo, hi, g = stats.numpy_virtual_index(n, q)
            vi = (n - 1) * (np.float64(q) / 100.0)     # numpy's expression
            assert 0 <= lo <= hi <= n - 1 and hi - lo in (0, 1)
            if hi > lo:
                assert lo == int(np.floor(vi)) and g == vi - np.floor(vi)
            if a is not None:
                want = np.percentile(a, np.float64(q))
                d = np.float64(hi - lo)
                mine = np.float64(lo) + d * g if g < 0.5 else np.float64(hi) - d * (1 - g)
                assert mine == want, (n, q, mine, want)


def test_pooled_percentile_near_tie_matches_numpy():
    """The literal case of the sweep: three pooled volumes (2 + 65536 + 1000 values), q = 98."""
    from adell_mri_b200 import stats
    from oracle.radix_select import NumpyKernels

    R = np.random.RandomState(0)
    for _ in range(200):
        sizes = [int(R.choice([2, 17, 1000, 4097, 65536])) for _ in range(3)]
        vols = [torch.from_numpy(R.normal(0, 100, size=s).astype(np.float32)) for s in sizes]
        qs = [1.0, 98.0, float(R.uniform(0, 100))]
        got = stats.percentiles(vols, qs, dataset_wide=True, kernels=NumpyKernels(vols)).numpy()[0]
        ref = np.percentile(np.concatenate([v.numpy() for v in vols]), np.asarray(qs, np.float64)).astype(np.float32)
        assert np.array_equal(got, ref), (sizes, qs, got, ref)


def test_clip_then_intensity_then_affine_keeps_the_clip_before_the_map():
    """the kernel applies pre -> clip -> post per tap; folding a pending post map into pre would compute
    clip(2v + 0.5) instead of 2 clip(v) + 0.5.  The clipped pass is closed first."""
    R = np.random.RandomState(9)
    img = torch.from_numpy((R.rand(1, 12, 10, 8) * 2 - 0.5).astype(np.float32))
    A = rand_affine_matrix(R)
    plan = BatchPlan([img[0]], strict=True).clip(0.0, 1.0).intensity(2.0, 0.5).affine(A.numpy(), "bilinear", "border")
    assert len(plan.passes) == 1
    clipped = torch.clamp(img, 0.0, 1.0) * torch.tensor(2.0) + torch.tensor(0.5)
    ref = M.affine_resample(clipped, A, "bilinear", "border")[0]
    assert mismatch(run_plan_cref(plan)[0], ref) == 0


def test_fast_mode_does_not_compose_across_a_crop_that_starts_at_zero():
    """a crop at start 0 only shrinks the grid; the second affine acts about the centre of the CROPPED
    volume, so the pass is closed exactly as for any other crop start."""
    R = np.random.RandomState(10)
    img = torch.from_numpy(R.rand(16, 16, 16).astype(np.float32))
    A1, A2 = rand_affine_matrix(R).numpy(), rand_affine_matrix(R).numpy()
    for start in (0, 1):
        plan = BatchPlan([img], fast=True).affine(A1, "bilinear", "zeros").crop([start] * 3, [8, 8, 8]).affine(A2, "bilinear", "zeros")
        assert len(plan.passes) == 1, start


def test_batch_posneg_picks_consume_the_crop_stream_like_the_dictionary_transform():
    """SegmentationBatchAugmenter.draw_picks (label-guided crop sandwich on the batch path) makes, sample by sample, the
    two draws per crop RandCropByPosNegLabeld.randomize makes on device-resident index lists — same stream, same
    values, same list chosen — including samples with an empty foreground."""
    import torch

    from adell_mri_b200 import transforms as T
    from adell_mri_b200.pipelines import SegmentationBatchAugmenter

    R = np.random.RandomState(4)
    shape, n_crops = (20, 18, 10), 3
    samples = []
    for i in range(5):
        m = (R.rand(*shape) > (1.1 if i == 2 else 0.8))          # sample 2: no foreground at all
        flat = torch.from_numpy(m.reshape(-1))
        samples.append({"mask_fg_indices": torch.nonzero(flat).reshape(-1), "mask_bg_indices": torch.nonzero(~flat).reshape(-1)})
    aug = SegmentationBatchAugmenter(["affine", "flip"], ["t2", "mask"], ["t2"], random_crop_size=[8, 8, 4], has_label=True, n_crops=n_crops)
    aug.crop_R = np.random.RandomState(77)
    picks = aug.draw_picks(samples, shape)
    t = T.RandCropByPosNegLabeld(["t2", "mask"], "mask", [8, 8, 4], num_samples=n_crops, allow_smaller=True)
    t.R = np.random.RandomState(77)
    want = []
    for s in samples:
        t.randomize(shape, s["mask_fg_indices"], s["mask_bg_indices"])
        want += t._picks
    assert len(picks) == len(want) == 5 * n_crops
    for (la, pa), (lb, pb) in zip(picks, want):
        assert la is lb and pa == pb
    assert aug.crop_R.randint(1 << 30) == t.R.randint(1 << 30)
