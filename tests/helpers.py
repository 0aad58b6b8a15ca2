"""Shared test helpers: run a composed BatchPlan through the C restatement (CPU) or the
CUDA path, and build reference chains with the literal torch oracle."""

from __future__ import annotations

import numpy as np
import torch

from adell_mri_b200.plan import BatchPlan
from oracle import cref
from oracle import monai_restated as M


def run_plan_cref(plan: BatchPlan) -> list[torch.Tensor]:
    """Execute every pass of a CPU-resident plan with oracle/gather_ref.c."""
    shape = plan.shape
    outs = [torch.empty(tuple(int(x) for x in s), dtype=torch.float32) for s in shape]
    dst_ptr = np.array([o.data_ptr() for o in outs], np.uint64)
    dst_stride = np.array([o.stride() for o in outs], np.int64)
    launches = plan.build_launches(dst_ptr, dst_stride, lambda n: torch.empty(max(n, 1), dtype=torch.float32))
    for items in launches:
        cref.gather(items)
    return outs


def run_plan_cuda(plan: BatchPlan) -> list[torch.Tensor]:
    from adell_mri_b200 import engine

    shape = plan.shape
    outs = [torch.empty(tuple(int(x) for x in s), dtype=torch.float32, device=plan.device) for s in shape]
    engine.execute(plan, outs)
    torch.cuda.synchronize()
    return outs


def rand_affine_matrix(R: np.random.RandomState, rotate=(0.4, 0.4, 0.2), shear=None, translate=(6, 6, 2), scale=(0.1, 0.1, 0.1)):
    p = M.rand_affine_grid_params(R, rotate_range=rotate, shear_range=shear, translate_range=translate, scale_range=scale)
    return M.compose_affine(p["rotate"], p["shear"], p["translate"], p["scale"])


def mismatch(a: torch.Tensor, b: torch.Tensor) -> int:
    return int((a != b).sum())


def cref_execute(plan: BatchPlan, dsts) -> None:
    """Stand-in for ``engine.execute`` on CPU-resident plans: every pass runs through the C
    restatement (oracle/gather_ref.c).  Lets the lazy dictionary transforms / collation be
    tested without a GPU; the GPU tests run the same pipelines through the CUDA path."""
    dst_ptr = np.array([d.data_ptr() for d in dsts], np.uint64)
    dst_stride = np.array([d.stride() for d in dsts], np.int64)
    launches = plan.build_launches(dst_ptr, dst_stride, lambda n: torch.empty(max(n, 1), dtype=torch.float32))
    for items in launches:
        cref.gather(items)


class _CrefSteps:
    def __init__(self, steps):
        self.steps = steps

    def __len__(self):
        return len(self.steps)

    def run(self, k):
        cref.gather(self.steps[k])


def cref_prepare_chain_steps(chains, step_sizes, device, keep=None):
    """Stand-in for ``engine.prepare_chain_steps`` on CPU-resident chains: the NATIVE composer builds the items
    (host policy only, nothing encoded) and each step then runs through the C restatement."""
    from adell_mri_b200 import engine
    from adell_mri_b200.plan import ITEM_DTYPE

    sizes = [int(x) for x in step_sizes]
    buf, offs, _ = engine.compose_chains_host(chains, sizes)
    steps = [buf[int(o): int(o) + n * engine.ISZ].view(ITEM_DTYPE) for n, o in zip(sizes, offs)]
    return _CrefSteps(steps)


class _CrefSeqSteps:
    def __init__(self, steps, keep):
        self.steps, self.keep = steps, keep

    def __len__(self):
        return len(self.steps)

    def launches(self, k):
        return len(self.steps[k])

    def run(self, k):
        for items in self.steps[k]:
            cref.gather(items)


def cref_prepare_seq_steps(seqs, step_sizes, device, keep=None):
    """Stand-in for ``engine.prepare_seq_steps`` on CPU-resident sequences: the NATIVE sequence composer builds every
    launch (scratch volumes in a host buffer), each then runs through the C restatement."""
    from adell_mri_b200 import engine

    sizes = [int(x) for x in step_sizes]
    _, need = engine.compose_seqs_host(seqs, sizes, mode=2, scratch_ptr=4096)
    scratch = torch.empty(max(int(need), 1), dtype=torch.float32)
    launches, _ = engine.compose_seqs_host(seqs, sizes, mode=1, scratch_ptr=scratch.data_ptr(), scratch_elems=scratch.numel())
    steps = [[] for _ in sizes]
    for k, items in launches:
        steps[k].append(items)
    return _CrefSeqSteps(steps, [scratch, keep])


def cref_execute_ptrs(plan: BatchPlan, dst_ptr, dst_stride, keep=None) -> None:
    """Stand-in for ``engine.execute_ptrs`` on CPU-resident plans (see :func:`cref_execute`)."""
    launches = plan.build_launches(dst_ptr, dst_stride, lambda n: torch.empty(max(n, 1), dtype=torch.float32))
    for items in launches:
        cref.gather(items)


def patch_engine_for_cpu(monkeypatch) -> None:
    """Route every execution entry of the engine through the C restatement (CPU tests of the host logic)."""
    from adell_mri_b200 import engine

    monkeypatch.setattr(engine, "execute", cref_execute)
    monkeypatch.setattr(engine, "execute_ptrs", cref_execute_ptrs)
    monkeypatch.setattr(engine, "prepare_chain_steps", cref_prepare_chain_steps)
    monkeypatch.setattr(engine, "prepare_seq_steps", cref_prepare_seq_steps)
