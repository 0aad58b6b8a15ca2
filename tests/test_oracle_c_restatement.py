"""Pins the C restatement (oracle/gather_ref.c) and the host composer (adell_mri_b200/plan.py)
against the literal torch oracle (oracle/monai_restated.py: MONAI op order + ATen grid_sample).
Runs on CPU; the CUDA path is compared with both in tests/test_gpu_parity.py."""

import numpy as np
import pytest
import torch

from adell_mri_b200.plan import BatchPlan
from oracle import monai_restated as M
from tests.helpers import mismatch, rand_affine_matrix, run_plan_cref

SHAPES = [(24, 20, 16), (33, 17, 9), (16, 16, 32)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("padding", ["zeros", "border", "reflection"])
def test_nearest_bit_exact(shape, padding):
    R = np.random.RandomState(hash((shape, padding)) % 2**31)
    img = torch.from_numpy(R.rand(1, *shape).astype(np.float32))
    for _ in range(4):
        A = rand_affine_matrix(R)
        ref = M.affine_resample(img, A, "nearest", padding)[0]
        plan = BatchPlan([img[0]], strict=True).affine(A.numpy(), "nearest", padding)
        out = run_plan_cref(plan)[0]
        assert mismatch(out, ref) == 0


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("padding", ["zeros", "border", "reflection"])
def test_trilinear_strict_bit_exact(shape, padding):
    R = np.random.RandomState(hash((shape, padding, 1)) % 2**31)
    img = torch.from_numpy(R.rand(1, *shape).astype(np.float32))
    for _ in range(4):
        A = rand_affine_matrix(R)
        ref = M.affine_resample(img, A, "bilinear", padding)[0]
        plan = BatchPlan([img[0]], strict=True).affine(A.numpy(), "bilinear", padding)
        out = run_plan_cref(plan)[0]
        assert mismatch(out, ref) == 0
        # default (fma) accumulation: within the stated tolerance
        plan = BatchPlan([img[0]]).affine(A.numpy(), "bilinear", padding)
        out = run_plan_cref(plan)[0]
        assert torch.allclose(out, ref, rtol=1e-4, atol=1e-6)


def test_far_out_of_range_coordinates():
    """Large translations / scales drive reflection through many periods."""
    R = np.random.RandomState(7)
    img = torch.from_numpy(R.rand(1, 12, 10, 8).astype(np.float32))
    for _ in range(6):
        A = rand_affine_matrix(R, rotate=(1.5, 1.5, 1.5), translate=(60, 50, 40), scale=(0.8, 0.8, 0.8))
        for padding in ["zeros", "border", "reflection"]:
            for mode in ["nearest", "bilinear"]:
                ref = M.affine_resample(img, A, mode, padding)[0]
                out = run_plan_cref(BatchPlan([img[0]], strict=True).affine(A.numpy(), mode, padding))[0]
                assert mismatch(out, ref) == 0, (padding, mode)
