"""RNG plumbing of the entrypoints: `get_generator_and_rng` / `get_global_rank`
against the reference's own definitions (/root/reference/adell_mri/utils/torch_utils.py:304-354).
The expected draws below were produced by executing those definitions (taken from the source with
`ast`, the technique of tests/golden/make_golden_collate.py) for seed 42 at ranks 0 and 3:
    g.initial_seed(), torch.randperm(8, generator=g), rng.integers(0, 1000, 4), np.random.rand(), torch.rand(1)
"""

import numpy as np
import pytest
import torch

from adell_mri_b200 import dist as D

EXPECTED = {  # rank -> (generator seed, randperm(8), rng.integers(0,1000,4), np.random.rand(), random.random())
    0: (42, [6, 3, 0, 7, 2, 1, 4, 5], [89, 773, 654, 438], 0.3745401188473625, 0.6394267984578837),
    3: (45, [3, 2, 0, 6, 4, 1, 7, 5], [89, 773, 654, 438], 0.3745401188473625, 0.6394267984578837),
}


@pytest.mark.parametrize("rank", [0, 3])
def test_generator_and_rng(monkeypatch, rank):
    import random

    for var in ("RANK", "LOCAL_RANK", "SLURM_PROCID", "SLURM_LOCALID"):
        monkeypatch.delenv(var, raising=False)
    if rank:
        monkeypatch.setenv("RANK", str(rank))
    assert D.get_global_rank() == rank
    g, rng = D.get_generator_and_rng(42)
    seed, perm, ints, npr, pyr = EXPECTED[rank]
    assert g.initial_seed() == seed
    assert torch.randperm(8, generator=g).tolist() == perm
    assert rng.integers(0, 1000, 4).tolist() == ints
    assert np.random.rand() == npr
    assert random.random() == pyr


def test_rank_precedence_follows_reference(monkeypatch):
    for var in ("RANK", "LOCAL_RANK", "SLURM_PROCID", "SLURM_LOCALID"):
        monkeypatch.delenv(var, raising=False)
    monkeypatch.setenv("SLURM_LOCALID", "5")
    assert D.get_global_rank() == 5
    monkeypatch.setenv("LOCAL_RANK", "2")
    assert D.get_global_rank() == 2
    monkeypatch.setenv("RANK", "7")
    assert D.get_global_rank() == 7
