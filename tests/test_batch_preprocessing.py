"""Batch-level mixing (label smoothing, mixup, partial mixup): the golden vectors were produced by the
reference module itself (tests/golden/make_golden_batch.py runs
/root/reference/adell_mri/utils/batch_preprocessing.py in the build container), so parity here is
PINNED against the reference, bit for bit.  CPU: the oracle restatement against the vectors.
GPU (marked): the CUDA path (`adell_mixup` behind `adell_mri_b200.batch_preprocessing`) against the
vectors and against the oracle at a collated-batch size."""

import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import batch_preprocessing_ref as O

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden_batch", os.path.join(HERE, "golden", "make_golden_batch.py"))
G = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(G)
GOLD = np.load(os.path.join(HERE, "golden", "batch_preprocessing.npz"))


@pytest.mark.parametrize("case", G.CASES, ids=[c[0] for c in G.CASES])
def test_oracle_restatement_equals_reference_vectors(case):
    name, shape, ydtype, kw, calls = case
    bp = O.BatchPreprocessing(**kw)
    for c in range(calls):
        x, y = G.inputs(name, shape, ydtype, c)
        X, Y = bp(torch.from_numpy(x.copy()), torch.from_numpy(y.copy()))
        assert np.array_equal(X.numpy(), GOLD[f"{name}/{c}/x"])
        assert np.array_equal(Y.numpy(), GOLD[f"{name}/{c}/y"]) and Y.numpy().dtype == GOLD[f"{name}/{c}/y"].dtype


def test_cpu_tensors_are_refused():
    from adell_mri_b200 import batch_preprocessing as B

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        B.mixup(torch.zeros(2, 1, 4, 4, 4), torch.zeros(2), 0.4, np.random.default_rng(0))


@pytest.mark.gpu
@pytest.mark.parametrize("case", G.CASES, ids=[c[0] for c in G.CASES])
def test_cuda_path_equals_reference_vectors(case):
    from adell_mri_b200 import batch_preprocessing as B

    name, shape, ydtype, kw, calls = case
    bp = B.BatchPreprocessing(**kw)
    for c in range(calls):
        x, y = G.inputs(name, shape, ydtype, c)
        X, Y = bp(torch.from_numpy(x.copy()).cuda(), torch.from_numpy(y.copy()).cuda())
        assert np.array_equal(X.cpu().numpy(), GOLD[f"{name}/{c}/x"])
        assert np.array_equal(Y.cpu().numpy(), GOLD[f"{name}/{c}/y"]) and Y.cpu().numpy().dtype == GOLD[f"{name}/{c}/y"].dtype


@pytest.mark.gpu
@pytest.mark.parametrize("partial", [None, 0.5])
def test_cuda_path_at_collated_batch_size(partial):
    """Config B's collated image batch [8, 3, 256, 256, 32] (odd per-sample offsets of the scalar path
    are covered by the small golden shapes): bit-exact against the oracle on the same draws."""
    from adell_mri_b200 import batch_preprocessing as B

    g = torch.Generator().manual_seed(5)
    x = torch.rand((8, 3, 256, 256, 32), generator=g)
    y = (torch.rand(8, generator=g) > 0.5).float()
    want_x, want_y = O.BatchPreprocessing(mixup_alpha=0.4, partial_mixup=partial, seed=9)(x.clone(), y.clone())
    got_x, got_y = B.BatchPreprocessing(mixup_alpha=0.4, partial_mixup=partial, seed=9)(x.cuda(), y.clone().cuda())
    assert torch.equal(got_x.cpu(), want_x)
    assert torch.equal(got_y.cpu(), want_y)
