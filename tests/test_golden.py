"""Committed golden vectors (tests/golden/k1_golden.npz, made by tests/golden/make_golden.py with
the torch oracle): the oracle itself, the C restatement and — on the GPU — the CUDA path must
all reproduce them bit for bit (strict mode)."""

import os

import numpy as np
import pytest
import torch

from adell_mri_b200.plan import BatchPlan
from oracle import monai_restated as M
from tests.helpers import mismatch, run_plan_cref, run_plan_cuda

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "k1_golden.npz"), allow_pickle=False)
N = int(G["n"])


def case(i):
    return {k.split(":", 1)[1]: G[k] for k in G.files if k.startswith(f"{i}:")}


def build_plan(c, device):
    plan = BatchPlan([torch.from_numpy(c["img"]).to(device)], strict=True)
    if len(c["pre_flip"]):
        plan.flip(np.array([a in c["pre_flip"] for a in range(3)]))
    plan.affine(c["A"], str(c["mode"]), str(c["padding"]))
    if len(c["post_flip"]):
        plan.flip(np.array([a in c["post_flip"] for a in range(3)]))
    if len(c["crop"]):
        plan.center_crop(c["crop"])
    return plan


@pytest.mark.parametrize("i", range(N))
def test_oracle_reproduces_golden(i):
    c = case(i)
    pre = [("flip", list(c["pre_flip"]))] if len(c["pre_flip"]) else []
    post = ([("flip", list(c["post_flip"]))] if len(c["post_flip"]) else []) + (
        [("center_crop", list(c["crop"]))] if len(c["crop"]) else [])
    out = M.canonical_item(torch.from_numpy(c["img"])[None], pre_ops=pre, affine=torch.from_numpy(c["A"]),
                           mode=str(c["mode"]), padding_mode=str(c["padding"]), post_ops=post)[0]
    assert np.array_equal(out.numpy(), c["out"])


@pytest.mark.parametrize("i", range(N))
def test_c_restatement_reproduces_golden(i):
    c = case(i)
    out = run_plan_cref(build_plan(c, "cpu"))[0]
    assert mismatch(out, torch.from_numpy(c["out"])) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(N))
def test_cuda_reproduces_golden(i):
    c = case(i)
    out = run_plan_cuda(build_plan(c, "cuda:0"))[0].cpu()
    assert mismatch(out, torch.from_numpy(c["out"])) == 0
