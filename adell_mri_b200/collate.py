"""Collation of the reference's loaders with the fused launch behind it.

``safe_collate`` / ``safe_collate_crops`` keep the reference's contract
(/root/reference/adell_mri/utils/utils.py:308-377: per key ``torch.stack`` of the samples, the
plain list when stacking is impossible; the crops variant flattens ``list[list[dict]]`` first).
Entries that are still :class:`~adell_mri_b200.transforms.Pending` — i.e. everything the lazy
dictionary transforms touched — are executed here: the recorded chains of every sample and every
key are concatenated and run as ONE K1 launch per resample pass, each volume written straight
into its slot of the collated ``[B, C, H, W, D]`` tensor (no per-sample tensors, no stack copy).
"""

from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from . import engine
from .transforms import Pending, replay_parts


def _cat(x):
    try:
        x = [torch.as_tensor(y) for y in x]
    except Exception:
        return x
    try:
        return torch.stack(x)
    except Exception:
        return x


def execute_pending(groups: Sequence[Sequence[Pending]]) -> list[torch.Tensor]:
    """``groups[g]`` = the entries of one key over the batch (same output shape): returns one
    ``[B, C, H, W, D]`` float32 tensor per group, all produced by the same fused launch(es).  The recorded chains
    of every sample and key are composed here, grouped by chain signature: each op is one vectorised call for all
    the volumes that recorded it (``transforms.replay_parts``), not one per sample."""
    parts, ptrs, strides, outs = [], [], [], []
    for entries in groups:
        first = entries[0]
        nc = first.n_channels
        out = torch.empty((len(entries), *first.shape), dtype=torch.float32, device=first.device)
        outs.append(out)
        for e in entries:
            if e.n_channels != nc:
                raise ValueError("entries of one key must have the same number of channels")
            parts.extend(e.parts)
        # destination of volume [b, c]: computed, not sliced (a tensor view per volume costs more than its chain)
        b, c = np.arange(len(entries), dtype=np.int64)[:, None], np.arange(nc, dtype=np.int64)[None, :]
        ptrs.append((out.data_ptr() + 4 * (b * out.stride(0) + c * out.stride(1))).reshape(-1).astype(np.uint64))
        strides.append(np.broadcast_to(np.asarray(out.stride()[2:], np.int64), (len(entries) * nc, 3)))
    if parts:
        engine.execute_ptrs(replay_parts(parts), np.concatenate(ptrs), np.concatenate(strides), keep=outs)
    return outs


def safe_collate(X):
    """Mirror of ``adell_mri.utils.safe_collate`` with fused execution of pending entries."""
    example = X[0]
    if isinstance(example, list):
        return [_cat(e) for e in zip(*X)]
    keys = list(example.keys())
    out, groups, group_keys = {}, [], []
    for k in keys:
        vals = [x[k] if k in x else None for x in X]
        if all(isinstance(v, Pending) for v in vals):
            if len({(v.shape, str(v.device)) for v in vals}) == 1:
                groups.append(vals)
                group_keys.append(k)
                continue
            # ragged shapes: the reference returns the list of per-sample tensors
            out[k] = [v.tensor() for v in vals]
            continue
        vals = [v.tensor() if isinstance(v, Pending) else v for v in vals]
        out[k] = _cat(vals)
    for k, t in zip(group_keys, execute_pending(groups)):
        out[k] = t
    return {k: out[k] for k in keys}


def safe_collate_crops(X):
    """Mirror of ``adell_mri.utils.safe_collate_crops`` (lists of crops are flattened first)."""
    flat = []
    for x in X:
        flat.extend(x)
    return safe_collate(flat)
