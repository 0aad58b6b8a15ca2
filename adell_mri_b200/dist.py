"""Multi-GPU plumbing of the hot path: one process per GPU (``torch.distributed``, NCCL).

Batches shard by sample, so the transforms need no collective — each rank augments the samples
it owns from its own device-resident cache, exactly as the reference shards by seeding every
rank's sampler with ``seed + rank`` (/root/reference/adell_mri/utils/torch_utils.py:326-354).
The single exchange on the path is the dataset-wide percentile: per-rank histogram counts are
summed over NCCL after every radix pass, then every rank runs the identical selection.
"""

from __future__ import annotations

import os
import random
from typing import Sequence

import numpy as np
import torch
import torch.distributed as dist

from . import stats


def get_global_rank() -> int:
    """/root/reference/adell_mri/utils/torch_utils.py:304-323: launcher environment first
    (``RANK``, ``LOCAL_RANK``, ``SLURM_PROCID``, ``SLURM_LOCALID``), then ``torch.distributed``."""
    for var in ("RANK", "LOCAL_RANK", "SLURM_PROCID", "SLURM_LOCALID"):
        value = os.environ.get(var)
        if value is not None:
            return int(value)
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank()
    return 0


def get_generator_and_rng(seed: int) -> tuple[torch.Generator, np.random.Generator]:
    """/root/reference/adell_mri/utils/torch_utils.py:326-354: the global torch /
    random / numpy streams and the numpy Generator are seeded identically on every rank (dataset
    construction stays consistent), the sampler's torch Generator with ``seed + global rank`` (each
    rank draws different samples: sharding by sample without a collective)."""
    torch.manual_seed(seed)
    random.seed(seed)
    np.random.seed(seed)
    g = torch.Generator()
    g.manual_seed(seed + get_global_rank())
    rng = np.random.default_rng(seed)
    return g, rng


def rank_world(group=None) -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_indices(n: int, rank: int | None = None, world: int | None = None) -> list[int]:
    """Samples owned by ``rank``: ``i ≡ rank (mod world)``."""
    if rank is None or world is None:
        rank, world = rank_world()
    return list(range(rank, n, world))


def dataset_percentiles(vols: Sequence[torch.Tensor], qs: Sequence[float], group=None, kernels=None) -> torch.Tensor:
    """Percentiles over the union of all ranks' volumes (numpy 'linear' semantics on the pooled
    data): ``[1, len(qs)]`` fp32, identical on every rank."""
    rank, world = rank_world(group)
    local_n = sum(v.numel() for v in vols)
    if world == 1:
        return stats.percentiles(vols, qs, dataset_wide=True, kernels=kernels)
    # the pooled element count: all-reduced once per kernel object (= per set of cached volumes), not per call — it costs a
    # collective AND a host synchronisation
    total = getattr(kernels, "pooled_total", None) if kernels is not None else None
    if total is None or total[0] != (local_n, world):
        n = torch.tensor([local_n], dtype=torch.int64, device=vols[0].device)
        dist.all_reduce(n, group=group)
        total = ((local_n, world), int(n.item()))
        if kernels is not None:
            try:
                kernels.pooled_total = total
            except AttributeError:
                pass
    return stats.percentiles(
        vols, qs, dataset_wide=True, total_n=total[1],
        all_reduce=lambda bins: dist.all_reduce(bins, group=group), kernels=kernels,
    )
