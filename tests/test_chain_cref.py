"""CPU: host composer + C restatement vs the op-by-op torch oracle, for whole chains."""

import pytest
import torch

from tests.cases import CHAIN_CASES, chain_case, noise_then_resample_case, two_resample_case
from tests.helpers import mismatch, run_plan_cref


@pytest.mark.parametrize("name,kw", CHAIN_CASES, ids=[c[0] for c in CHAIN_CASES])
def test_chain_bit_exact(name, kw):
    for seed in range(3):
        plan, ref = chain_case(seed, **kw)
        out = run_plan_cref(plan)[0]
        assert out.shape == ref.shape
        assert mismatch(out, ref) == 0, name


@pytest.mark.parametrize("name,kw", CHAIN_CASES[:8], ids=[c[0] for c in CHAIN_CASES[:8]])
def test_chain_default_mode_within_tolerance(name, kw):
    plan, ref = chain_case(11, strict=False, **kw)
    out = run_plan_cref(plan)[0]
    tol = 1e-4 * float(ref.abs().max())
    assert torch.allclose(out, ref, rtol=1e-4, atol=tol)


def test_two_resamples_are_two_passes():
    plan, ref = two_resample_case(3)
    assert len(plan.passes) == 1
    assert mismatch(run_plan_cref(plan)[0], ref) == 0


def test_noise_then_resample_closes_a_pass():
    plan, ref = noise_then_resample_case(5)
    assert len(plan.passes) == 1
    assert mismatch(run_plan_cref(plan)[0], ref) == 0
