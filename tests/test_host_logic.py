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


@pytest.mark.parametrize("prob", [0.2, 1.0])
def test_rand_affine_sampler_follows_the_restated_stream(prob):
    kw = dict(rotate_range=[np.pi / 8, np.pi / 8, np.pi / 16], scale_range=[0.1, 0.1, 0.05])
    a = RandAffineSampler(prob=prob, **kw).set_random_state(7)
    b = M.RandAffinedDraws(prob, n_keys=4, **kw).set_random_state(7)
    for _ in range(20):
        fa, pa = a.draw(n_keys=4)
        fb, pb = b.draw()
        assert fa == fb and pa == pb


def test_vectorised_draws_equal_sequential_draws():
    for trial in range(12):
        kw = dict(prob=[0.2, 0.5, 1.0][trial % 3], rotate_range=[0.3, 0.3, 0.1],
                  shear_range=((0.9, 1.1),) * 3 if trial % 2 else None,
                  translate_range=[4, 4, 1] if trial % 4 == 0 else None)
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
