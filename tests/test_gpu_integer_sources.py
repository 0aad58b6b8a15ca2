"""GPU: int16 / uint8 cached volumes through the TMA-staged paths (config B stores its
volumes as int16, the mask as uint8).  The boxes are staged in the SOURCE element type and converted at the tap;
the items must take the staged / box-copy kinds (not the generic global-memory path) and keep the contract:
identity copies and nearest resamples bit-exact, trilinear <= 1e-4 of the range."""

import ctypes as C

import numpy as np
import pytest
import torch

from adell_mri_b200 import _lib, engine
from adell_mri_b200.plan import BatchPlan
from oracle import monai_restated as M
from tests.helpers import mismatch, rand_affine_matrix, run_plan_cuda

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
KIND_GENERIC, KIND_STAGED, KIND_VCOPY = 0, 1, 2


def _vol(R, shape, dtype):
    if dtype == "i16":
        return torch.from_numpy(R.randint(-300, 4000, size=(1, *shape)).astype(np.int16))
    return torch.from_numpy(R.randint(0, 256, size=(1, *shape)).astype(np.uint8))


def _kinds(plan):
    outs = [torch.empty(tuple(int(x) for x in s), device=DEV) for s in plan.shape]
    dst_ptr = np.array([o.data_ptr() for o in outs], np.uint64)
    dst_stride = np.array([o.stride() for o in outs], np.int64)
    items = plan.build_launches(dst_ptr, dst_stride, lambda n: torch.empty(max(n, 1), device=DEV))[-1]
    buf, n, info = engine.pack_launch(items)
    return buf[: n * engine.ISZ].view(engine.ITEM_DTYPE)["kind"].copy()


@pytest.mark.parametrize("dtype", ["i16", "u8"])
@pytest.mark.parametrize("shape", [(256, 256, 32), (96, 80, 48), (64, 72, 64)])
@pytest.mark.parametrize("padding", ["zeros", "border", "reflection"])
def test_integer_sources_are_staged_and_match_the_oracle(dtype, shape, padding):
    R = np.random.RandomState(hash((dtype, shape, padding)) % 1000)
    img = _vol(R, shape, dtype)
    for i in range(3):
        A = rand_affine_matrix(R, rotate=(np.pi / 8, np.pi / 8, np.pi / 16), translate=(4, 4, 1), scale=(0.1, 0.1, 0.05))
        flips = [a for a in range(3) if R.rand() < 0.5]
        for mode in ("bilinear", "nearest"):
            ref = M.canonical_item(img, affine=A, mode=mode, padding_mode=padding, post_ops=[("flip", flips)] if flips else [])[0]
            mk = lambda: BatchPlan([img[0].to(DEV)]).affine(A.numpy(), mode, padding).flip(np.array([a in flips for a in range(3)]))
            # staged wherever an fp32 copy of the volume is (i.e. wherever the footprint box fits)
            f32 = BatchPlan([img[0].float().to(DEV)]).affine(A.numpy(), mode, padding).flip(np.array([a in flips for a in range(3)]))
            assert (_kinds(mk())[_kinds(f32) == KIND_STAGED] == KIND_STAGED).all(), (dtype, shape, mode)   # (smaller elements may fit where fp32 does not)
            if shape == (256, 256, 32):
                assert (_kinds(mk()) == KIND_STAGED).all(), (dtype, shape, mode)
            out = run_plan_cuda(mk())[0].cpu()
            if mode == "nearest":
                assert mismatch(out, ref) == 0, (dtype, shape, padding, i)
            else:
                assert torch.allclose(out, ref, rtol=1e-4, atol=1e-4 * float(ref.abs().max())), float((out - ref).abs().max())


@pytest.mark.parametrize("dtype", ["i16", "u8"])
def test_integer_identity_items_take_the_box_copy_and_are_bit_exact(dtype):
    """Flip / crop copies of integer volumes: converted on the way out of the staged box; windows that start
    mid-row (16-byte alignment slack of 7 int16 / 15 uint8 elements), every flip combination, an intensity map."""
    R = np.random.RandomState(3)
    shape = (64, 48, 64)
    img = _vol(R, shape, dtype)
    dev = img[0].to(DEV)
    for trial in range(16):
        flips = np.array([(trial >> a) & 1 for a in range(3)], bool)
        start = (int(R.randint(8)), int(R.randint(8)), int(R.randint(20)))
        roi = (48, 32, 32)
        plan = BatchPlan([dev]).crop(start, roi).flip(flips)
        ref = M.crop(img, start, roi)
        fl = [a for a in range(3) if flips[a]]
        ref = (M.flip(ref, fl) if fl else ref).to(torch.float32)
        if trial % 3 == 2:
            plan.intensity(0.5, 0.25)
            ref = ref * torch.tensor(0.5) + torch.tensor(0.25)   # exact in fp32 for these magnitudes
        mk_kind = _kinds(BatchPlan([dev]).crop(start, roi).flip(flips))
        assert (mk_kind == KIND_VCOPY).all(), (dtype, trial, mk_kind)
        out = run_plan_cuda(plan)[0].cpu()
        assert mismatch(out, ref[0]) == 0, (dtype, trial)


@pytest.mark.parametrize("dtype", ["i16", "u8"])
def test_integer_sources_with_crop_windows_pad_bands_and_device_scaling(dtype):
    """The rest of the staged variants on integer boxes: a crop window before the resample that starts mid-row
    (alignment-slack columns zeroed in the source element type), a pad band after it (cold loop), and a
    {scale, offset} pair read from device memory with zeros padding (valid-weight variant)."""
    R = np.random.RandomState(9)
    shape = (72, 64, 48)
    img = _vol(R, shape, dtype)
    dev = img[0].to(DEV)
    for trial in range(4):
        A = rand_affine_matrix(R, rotate=(0.3, 0.3, 0.15), translate=(3, 3, 1), scale=(0.08, 0.08, 0.04))
        start, roi = (int(R.randint(6)), int(R.randint(6)), 1 + int(R.randint(9))), (56, 48, 32)
        for mode in ("bilinear", "nearest"):
            # crop -> affine(border) -> pad
            plan = BatchPlan([dev]).crop(start, roi).affine(A.numpy(), mode, "border").spatial_pad((60, 52, 40))
            ref = M.spatial_pad(M.affine_resample(M.crop(img, start, roi), A, mode, "border"), (60, 52, 40))[0]
            out = run_plan_cuda(plan)[0].cpu()
            if mode == "nearest":
                assert mismatch(out, ref) == 0, (dtype, trial)
            else:
                assert torch.allclose(out, ref, rtol=1e-4, atol=1e-4 * float(ref.abs().max()))
            # device-side {scale, offset}, zeros padding: the offset must not leak into the padding
            pre = torch.tensor([[1.0 / 4000.0, 0.125]], device=DEV)
            plan = BatchPlan([dev]).intensity_from_device(pre).affine(A.numpy(), mode, "zeros")
            scaled = img.to(torch.float32) * torch.tensor(1.0 / 4000.0) + torch.tensor(0.125)
            ref = M.affine_resample(scaled, A, mode, "zeros")[0]
            out = run_plan_cuda(plan)[0].cpu()
            assert torch.allclose(out, ref, rtol=1e-4, atol=1e-4), (dtype, trial, mode, float((out - ref).abs().max()))
