"""GPU parity tests proper: the CUDA path (through the C ABI) vs the oracle on the same
seeded inputs.  Bit-exact for integer/index work, nearest masks and strict trilinear;
<=1e-4 (relative to the reference / its dynamic range) for the default fused trilinear."""

import numpy as np
import pytest
import torch

from adell_mri_b200.plan import BatchPlan
from oracle import monai_restated as M
from tests.cases import CHAIN_CASES, chain_case, noise_then_resample_case, two_resample_case
from tests.helpers import mismatch, rand_affine_matrix, run_plan_cref, run_plan_cuda

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("name,kw", CHAIN_CASES, ids=[c[0] for c in CHAIN_CASES])
def test_chain_bit_exact_vs_torch_oracle(name, kw):
    for seed in range(3):
        plan, ref = chain_case(seed, device=DEV, **kw)
        out = run_plan_cuda(plan)[0].cpu()
        assert out.shape == ref.shape
        assert mismatch(out, ref) == 0, name


@pytest.mark.parametrize("name,kw", CHAIN_CASES, ids=[c[0] for c in CHAIN_CASES])
def test_chain_default_mode(name, kw):
    """Default (fast) mode: tile-local incremental coordinates + nested lerps for trilinear
    (<=1e-4 of the reference / its dynamic range); nearest and integer work stay bit-exact."""
    for seed in (21, 22):
        plan, ref = chain_case(seed, device=DEV, strict=False, **kw)
        out = run_plan_cuda(plan)[0].cpu()
        if kw["mode"] == "nearest" or not kw["affine"]:
            exact_ok = not (kw.get("post_scale") or kw.get("post_offset"))
            if exact_ok:
                assert mismatch(out, ref) == 0, name
                continue
        tol = 1e-4 * float(ref.abs().max())
        assert torch.allclose(out, ref, rtol=1e-4, atol=tol), float((out - ref).abs().max())


@pytest.mark.parametrize("shape", [(64, 64, 32), (96, 80, 48), (40, 36, 20)])
@pytest.mark.parametrize("padding", ["zeros", "border", "reflection"])
def test_default_mode_nearest_is_bit_exact(shape, padding):
    """Masks through the staged fast path (fast coordinates + exact replay near ties)."""
    R = np.random.RandomState(5)
    img = torch.from_numpy((R.rand(1, *shape) * 7).astype(np.int32).astype(np.float32))
    for i in range(6):
        A = rand_affine_matrix(R, rotate=(np.pi / 8, np.pi / 8, np.pi / 16), translate=(4, 4, 1), scale=(0.1, 0.1, 0.05))
        flips = [a for a in range(3) if R.rand() < 0.5]
        ref = M.canonical_item(img, pre_ops=[("flip", flips)] if (flips and i % 2) else [], affine=A, mode="nearest",
                               padding_mode=padding, post_ops=[("flip", flips)] if (flips and not i % 2) else [])[0]
        plan = BatchPlan([img[0].to(DEV)])
        if flips and i % 2:
            plan.flip(np.array([a in flips for a in range(3)]))
        plan.affine(A.numpy(), "nearest", padding)
        if flips and not i % 2:
            plan.flip(np.array([a in flips for a in range(3)]))
        assert mismatch(run_plan_cuda(plan)[0].cpu(), ref) == 0


@pytest.mark.parametrize("shape", [(256, 256, 32), (208, 208, 64), (320, 288, 40)])
@pytest.mark.parametrize("padding", ["zeros", "border", "reflection"])
def test_default_mode_nearest_is_bit_exact_at_benchmark_sizes(shape, padding):
    """Masks at the BASELINE volume sizes: sheared column tiles, wide coordinate ranges (the tie
    window of the fast coordinates scales with the extent), thin reflected axis."""
    R = np.random.RandomState(15)
    img = torch.from_numpy((R.rand(1, *shape) * 5).astype(np.int32).astype(np.float32))
    for i in range(4):
        A = rand_affine_matrix(R, rotate=(np.pi / 8, np.pi / 8, np.pi / 16), translate=(6, 6, 2), scale=(0.1, 0.1, 0.05))
        ref = M.affine_resample(img, A, "nearest", padding)[0]
        out = run_plan_cuda(BatchPlan([img[0].to(DEV)]).affine(A.numpy(), "nearest", padding))[0].cpu()
        assert mismatch(out, ref) == 0, (shape, padding, i)


@pytest.mark.parametrize("mode", ["nearest", "bilinear"])
@pytest.mark.parametrize("shape", [(96, 80, 48), (64, 200, 32)])
def test_pre_offset_does_not_leak_into_zero_padding(shape, mode):
    """Intensity map with an offset BEFORE a zeros-padded resample (percentile / min-max scaling folded
    into the gather): out-of-volume taps are literal zeros of the padded volume, so the offset only
    counts with the weights of the valid taps.  Tiles that straddle the volume edge take the fast
    loops' valid-weight variant; far translations make most tiles straddle or miss the volume."""
    R = np.random.RandomState(17)
    img = torch.from_numpy((R.rand(1, *shape) * 5).astype(np.int32).astype(np.float32))
    for i in range(4):
        A = rand_affine_matrix(R, rotate=(np.pi / 8, np.pi / 8, np.pi / 16), translate=(30, 25, 12), scale=(0.15, 0.15, 0.1))
        pre = (img * np.float32(1.25) + np.float32(0.75)).to(torch.float32)
        ref = M.affine_resample(pre, A, mode, "zeros")[0]
        plan = BatchPlan([img[0].to(DEV)]).intensity(scale=1.25, offset=0.75).affine(A.numpy(), mode, "zeros")
        out = run_plan_cuda(plan)[0].cpu()
        if mode == "nearest":
            assert mismatch(out, ref) == 0, (shape, i)
        else:
            assert float((out - ref).abs().max()) <= 1e-4 * float(ref.abs().max()), (shape, i)


@pytest.mark.parametrize("padding,mode", [("zeros", "nearest"), ("reflection", "nearest"), ("border", "bilinear"), ("zeros", "bilinear")])
def test_config_e_volume_against_oracle(padding, mode):
    """Config E size (512x512x128, 134 MB): 16 column groups of sheared tiles, coordinates up to 512
    (widest tie window); nearest bit-exact, trilinear within 1e-4 of the oracle."""
    R = np.random.RandomState(21)
    shape = (512, 512, 128)
    img = torch.from_numpy((R.rand(1, *shape) * 5).astype(np.int32).astype(np.float32))
    A = rand_affine_matrix(R, rotate=(np.pi / 8, np.pi / 8, np.pi / 16), translate=(8, 8, 3), scale=(0.1, 0.1, 0.05))
    ref = M.affine_resample(img, A, mode, padding)[0]
    out = run_plan_cuda(BatchPlan([img[0].to(DEV)]).affine(A.numpy(), mode, padding))[0].cpu()
    if mode == "nearest":
        assert mismatch(out, ref) == 0
    else:
        assert float((out - ref).abs().max()) <= 1e-4 * float(ref.abs().max())


@pytest.mark.parametrize("padding", ["zeros", "border", "reflection"])
def test_default_mode_trilinear_tolerance_large(padding):
    R = np.random.RandomState(6)
    shape = (128, 128, 32)
    img = torch.from_numpy(R.rand(1, *shape).astype(np.float32))
    worst = 0.0
    for i in range(4):
        A = rand_affine_matrix(R, rotate=(np.pi / 6,) * 3, translate=(10, 10, 3), scale=(0.1, 0.1, 0.1))
        ref = M.affine_resample(img, A, "bilinear", padding)[0]
        out = run_plan_cuda(BatchPlan([img[0].to(DEV)]).affine(A.numpy(), "bilinear", padding))[0].cpu()
        worst = max(worst, float((out - ref).abs().max()))
    assert worst <= 1e-4, worst


def test_multi_pass_chains():
    for fn in (two_resample_case, noise_then_resample_case):
        plan, ref = fn(3, device=DEV)
        assert mismatch(run_plan_cuda(plan)[0].cpu(), ref) == 0


@pytest.mark.parametrize("padding", ["zeros", "border", "reflection"])
@pytest.mark.parametrize("mode", ["nearest", "bilinear"])
def test_benchmark_random_affine_shapes(mode, padding):
    """Config A: the reference's own micro-benchmark shapes and ranges
    (/root/reference/benchmarks/benchmark-random-affine.py:90-103)."""
    R = np.random.RandomState(42)
    for shape in [(128, 128, 16), (64, 64, 8), (128, 128, 32)]:
        img = torch.from_numpy(R.rand(1, *shape).astype(np.float32))
        A = rand_affine_matrix(R, rotate=(np.pi / 6,) * 3, translate=(10, 10, 3), scale=(0.1, 0.1, 0.1))
        ref = M.affine_resample(img, A, mode, padding)[0]
        out = run_plan_cuda(BatchPlan([img[0].to(DEV)], strict=True).affine(A.numpy(), mode, padding))[0].cpu()
        assert mismatch(out, ref) == 0


def test_far_out_of_range_coordinates():
    R = np.random.RandomState(7)
    img = torch.from_numpy(R.rand(1, 12, 10, 8).astype(np.float32))
    for _ in range(6):
        A = rand_affine_matrix(R, rotate=(1.5, 1.5, 1.5), translate=(60, 50, 40), scale=(0.8, 0.8, 0.8))
        for padding in ["zeros", "border", "reflection"]:
            for mode in ["nearest", "bilinear"]:
                ref = M.affine_resample(img, A, mode, padding)[0]
                out = run_plan_cuda(BatchPlan([img[0].to(DEV)], strict=True).affine(A.numpy(), mode, padding))[0].cpu()
                assert mismatch(out, ref) == 0, (padding, mode)


def test_segmentation_batch_collated_in_place():
    """Config B at reduced size: 3 image keys + mask per sample written straight into the
    collated [B,C,H,W,D] batch by ONE launch; mask bit-exact, images bit-exact (strict)."""
    from adell_mri_b200 import engine

    R = np.random.RandomState(1)
    B, shape = 4, (64, 64, 32)
    image = torch.zeros(B, 3, *shape, device=DEV)
    mask = torch.zeros(B, 1, *shape, device=DEV)
    plans, refs, dsts = [], [], []
    for b in range(B):
        fired = R.rand() < 0.6
        A = rand_affine_matrix(R, rotate=(np.pi / 8, np.pi / 8, np.pi / 16), translate=None, scale=None)
        flips = [a for a in range(3) if R.rand() < 0.25]
        vols = [torch.from_numpy(R.rand(1, *shape).astype(np.float32)) for _ in range(3)]
        vols.append(torch.from_numpy((R.rand(1, *shape) > 0.7).astype(np.float32)))
        modes = ["bilinear"] * 3 + ["nearest"]
        for v, m in zip(vols, modes):
            refs.append(M.canonical_item(v, affine=A if fired else None, mode=m, padding_mode="reflection",
                                         post_ops=[("flip", flips)] if flips else [])[0])
        plan = BatchPlan([v[0].to(DEV) for v in vols], strict=True)
        plan.affine(A.numpy(), modes, "reflection", where=fired)
        plan.flip(np.array([a in flips for a in range(3)]))
        plans.append(plan)
        dsts += [image[b, 0], image[b, 1], image[b, 2], mask[b, 0]]
    before = engine.launch_count
    engine.execute(BatchPlan.concat(plans), dsts)
    torch.cuda.synchronize()
    assert engine.launch_count - before == 1
    for d, r in zip(dsts, refs):
        assert mismatch(d.cpu(), r) == 0


def test_full_size_properties():
    """BASELINE full sizes (256x256x32, 192x192x48, 512x512x128): size-independent properties."""
    g = torch.Generator(device=DEV).manual_seed(0)
    for shape in [(256, 256, 32), (192, 192, 48), (512, 512, 128)]:
        x = torch.rand(shape, device=DEV, generator=g)
        # identity affine: nearest is an exact copy; trilinear is an exact copy when every extent is a
        # power of two (2/S exact => weights exactly 1 and 0), else equal up to coordinate rounding
        pow2 = all(s & (s - 1) == 0 for s in shape)
        for mode in ("nearest", "bilinear"):
            out = run_plan_cuda(BatchPlan([x], strict=True).affine(np.eye(4, dtype=np.float32), mode, "zeros"))[0]
            if mode == "nearest" or pow2:
                assert torch.equal(out, x)
            else:
                assert torch.allclose(out, x, rtol=0, atol=1e-4)
        # flip twice == identity; flip == torch.flip; crop == slicing
        out = run_plan_cuda(BatchPlan([x]).flip(np.array([True, True, True])))[0]
        assert torch.equal(out, torch.flip(x, [0, 1, 2]))
        out = run_plan_cuda(BatchPlan([x]).flip(np.array([True, False, True])).flip(np.array([True, False, True])))[0]
        assert torch.equal(out, x)
        crop = tuple(s // 2 for s in shape)
        out = run_plan_cuda(BatchPlan([x]).center_crop(crop))[0]
        sl = tuple(slice(s // 2 - c // 2, s // 2 - c // 2 + c) for s, c in zip(shape, crop))
        assert torch.equal(out, x[sl])
        # linearity of the trilinear resample: T(a*x) == a*T(x) up to rounding
        A = rand_affine_matrix(np.random.RandomState(3)).numpy()
        t1 = run_plan_cuda(BatchPlan([x]).affine(A, "bilinear", "zeros"))[0]
        t2 = run_plan_cuda(BatchPlan([x * 2.0]).affine(A, "bilinear", "zeros"))[0]
        assert torch.allclose(t2, 2.0 * t1, rtol=1e-5, atol=1e-6)
        # 90-degree rotation about axis 2 twice == flip of axes 0 and 1 (square in-plane shapes)
        if shape[0] == shape[1]:
            Rz = M.compose_affine(rotate=[0.0, 0.0, np.pi / 2]).numpy()
            Rz = np.round(Rz)  # exact quarter turn
            r2 = run_plan_cuda(BatchPlan([x], strict=True).affine(Rz, "nearest", "zeros").affine(Rz, "nearest", "zeros"))[0]
            assert torch.equal(r2, torch.flip(x, [0, 1]))


def test_philox_noise_statistics():
    x = torch.zeros(64, 64, 32, device=DEV)
    out = run_plan_cuda(BatchPlan([x]).add_philox_noise(0.5, seed=1234))[0]
    assert abs(float(out.mean())) < 5e-3
    assert abs(float(out.std()) - 0.5) < 5e-3
    out2 = run_plan_cuda(BatchPlan([x]).add_philox_noise(0.5, seed=1234))[0]
    assert torch.equal(out, out2)


def test_philox_noise_depends_on_the_output_position_only():
    """The device noise of a voxel is a function of (seed, offset, output position): the generic path of an identity
    item, the staged path of a resampled item (whatever its tile shape) and a second staged geometry add the very same
    values (zero volumes, zeros padding: the output IS the noise)."""
    for shape in [(48, 40, 32), (33, 22, 16), (64, 64, 20)]:
        x = torch.zeros(shape, device=DEV)
        a = run_plan_cuda(BatchPlan([x]).add_philox_noise(0.5, seed=77, offset=123))[0]
        R = np.random.RandomState(sum(shape))
        for _ in range(2):
            A = rand_affine_matrix(R, rotate=(0.3, 0.2, 0.1), translate=(1, 1, 1), scale=(0.05, 0.05, 0.05)).numpy()
            b = run_plan_cuda(BatchPlan([x]).affine(A, "bilinear", "zeros").add_philox_noise(0.5, seed=77, offset=123))[0]
            assert torch.equal(a, b)
        assert abs(float(a.std()) - 0.5) < 2e-2 and abs(float(a.mean())) < 2e-2


def test_philox_noise_on_resampled_tiles_is_the_plain_tile_plus_the_noise():
    """Tiles whose only special feature is device noise run the plain trilinear loops with a noise epilogue: the result is
    fmaf(std, N(position), plain value) bit for bit, with N read back from an identity item over zeros (std 1: the output
    IS the noise) and the plain value from the same item without noise — every padding mode, ragged and multi-tile shapes."""
    R = np.random.RandomState(5)
    for shape, padding in [((64, 64, 32), "reflection"), ((48, 40, 32), "border"), ((33, 22, 16), "zeros"), ((96, 80, 24), "reflection")]:
        x = torch.from_numpy(R.rand(*shape).astype(np.float32)).to(DEV)
        N = run_plan_cuda(BatchPlan([torch.zeros(shape, device=DEV)]).add_philox_noise(1.0, seed=9, offset=4321))[0]
        A = rand_affine_matrix(R, rotate=(0.3, 0.2, 0.1), translate=(2, 2, 1), scale=(0.1, 0.1, 0.1)).numpy()
        plain = run_plan_cuda(BatchPlan([x]).affine(A, "bilinear", padding).intensity(1.5, -0.25))[0]
        noisy = run_plan_cuda(BatchPlan([x]).affine(A, "bilinear", padding).intensity(1.5, -0.25).add_philox_noise(0.3, seed=9, offset=4321))[0]
        want = (np.float64(np.float32(0.3)) * N.cpu().numpy().astype(np.float64) + plain.cpu().numpy().astype(np.float64)).astype(np.float32)
        assert np.array_equal(noisy.cpu().numpy(), want), (shape, padding)


def test_c_abi_rejects_bad_items():
    import ctypes as C

    from adell_mri_b200 import _lib

    lib = _lib.load()
    from adell_mri_b200.engine import aligned_items
    items = aligned_items(1)
    tiles = np.zeros(2, np.int32)
    info = _lib.LaunchInfo()
    assert lib.adell_aug_prepare(items.ctypes.data, 1, tiles.ctypes.data, C.byref(info)) == -1
    info.total_tiles = 1
    assert lib.adell_aug_gather(8, 8, 1, C.byref(info), None) == -3  # misaligned items pointer
    assert lib.adell_aug_gather(None, None, 0, C.byref(info), None) == 0  # empty batch is a no-op


def test_many_small_items_and_mixed_kinds_in_one_launch():
    """More items than the kernel's shared-memory tile-prefix cache holds (2048), of every kind in
    one launch: TMA-staged resamples, box copies (aligned and unaligned crop windows), and the
    generic path (int16 source, strict), written into one collated tensor."""
    from adell_mri_b200 import engine

    R = np.random.RandomState(11)
    n, shape = 2300, (12, 10, 8)
    base = torch.from_numpy(R.rand(n, *shape).astype(np.float32))
    base_i16 = torch.from_numpy(R.randint(-100, 3000, size=(n, *shape)).astype(np.int16))
    dev_f, dev_i = base.to(DEV), base_i16.to(DEV)
    out = torch.zeros(n, 8, 8, 8, device=DEV)
    plans, refs = [], []
    A = rand_affine_matrix(R, rotate=(0.3, 0.2, 0.1), translate=(1, 1, 1), scale=(0.05, 0.05, 0.05))
    for i in range(n):
        kind = i % 4
        start = (int(R.randint(5)), int(R.randint(3)), int(R.randint(1)) if kind != 2 else 1 - (i % 2) * 0)
        if kind == 0:      # staged resample, then crop
            p = BatchPlan([dev_f[i]]).affine(A.numpy(), "bilinear", "zeros").crop(start, (8, 8, 8))
            r = M.crop(M.affine_resample(base[i][None], A, "bilinear", "zeros"), start, (8, 8, 8))[0]
        elif kind == 1:    # aligned box copy with flips
            p = BatchPlan([dev_f[i]]).flip(np.array([True, False, True])).crop((start[0], start[1], 0), (8, 8, 8))
            r = M.crop(M.flip(base[i][None], [0, 2]), (start[0], start[1], 0), (8, 8, 8))[0]
        elif kind == 2:    # box copy of a window that starts mid-row (unaligned)
            p = BatchPlan([dev_f[i, :, :, 1:]]).crop((start[0], start[1], 0), (8, 8, 4)).spatial_pad((8, 8, 8))
            r = M.spatial_pad(M.crop(base[i, :, :, 1:][None], (start[0], start[1], 0), (8, 8, 4)), (8, 8, 8))[0]
        else:              # generic path: int16 source, nearest
            p = BatchPlan([dev_i[i]], strict=True).affine(A.numpy(), "nearest", "border").center_crop((8, 8, 8))
            r = M.center_spatial_crop(M.affine_resample(base_i16[i][None], A, "nearest", "border"), (8, 8, 8))[0]
        plans.append(p); refs.append(r)
    before = engine.launch_count
    engine.execute(BatchPlan.concat(plans), [out[i] for i in range(n)])
    torch.cuda.synchronize()
    assert engine.launch_count - before == 1
    got = out.cpu()
    for i in range(n):
        if i % 4 == 0:
            assert torch.allclose(got[i], refs[i], rtol=1e-4, atol=1e-4), i
        else:
            assert torch.equal(got[i], refs[i]), (i, i % 4)


@pytest.mark.parametrize("mix", ["copies_only", "resampled_only", "mixed", "single_tile"])
def test_chunk_queues_rearm_on_relaunch(mix):
    """The launch's chunk queues live in the launch buffer and are re-armed by the kernel itself: the
    same packed buffer launched five times in a row (outputs wiped in between) must give the same
    result every time — for launches that feed only the copy queue, only the resampled queue, both
    (items reordered copy-last by adell_aug_prepare), and a launch of one tile (fewer units than
    streams)."""
    from adell_mri_b200 import engine

    R = np.random.RandomState(23)
    shape = (8, 8, 8) if mix == "single_tile" else (40, 36, 32)
    n = 1 if mix == "single_tile" else 10
    vols = [torch.from_numpy(R.rand(*shape).astype(np.float32)) for _ in range(n)]
    A = rand_affine_matrix(R, rotate=(0.3, 0.3, 0.15), translate=(2, 2, 1), scale=(0.05, 0.05, 0.05))
    fired = {"copies_only": [False] * n, "resampled_only": [True] * n, "mixed": [i % 3 == 0 for i in range(n)], "single_tile": [False]}[mix]
    flips = np.array([[bool(i & 1), bool(i & 2), bool(i & 4)] for i in range(n)])
    plan = BatchPlan([v.to(DEV) for v in vols])
    plan.affine(A.numpy(), "bilinear", "zeros", where=np.array(fired))
    plan.flip(flips)
    refs = []
    for i, v in enumerate(vols):
        r = v[None]
        if fired[i]:
            r = M.affine_resample(r, A, "bilinear", "zeros")
        fl = [a for a in range(3) if flips[i, a]]
        refs.append((M.flip(r, fl) if fl else r)[0])
    out = torch.empty(n, *shape, device=DEV)
    dst_ptr = np.array([out[i].data_ptr() for i in range(n)], np.uint64)
    dst_stride = np.array([out[i].stride() for i in range(n)], np.int64)
    items = plan.build_launches(dst_ptr, dst_stride, None)[-1]
    buf, cnt, info = engine.pack_launch(items)
    dev_buf = torch.from_numpy(buf).to(DEV)
    for rep in range(5):
        out.fill_(float("nan"))
        engine.launch_packed(dev_buf, cnt, info)
        torch.cuda.synchronize()
        got = out.cpu()
        for i in range(n):
            if fired[i]:
                assert torch.allclose(got[i], refs[i], rtol=1e-4, atol=1e-4), (mix, rep, i)
            else:
                assert torch.equal(got[i], refs[i]), (mix, rep, i)
    q = dev_buf[cnt * engine.ISZ + 4 * (cnt + 1):].view(torch.int32).cpu()
    assert q.tolist()[:3] == [0, 0, 0]   # both queues and the drained-producer count are back to zero


@pytest.mark.parametrize("shape,roi,start", [
    ((48, 40, 32), (48, 40, 32), (0, 0, 0)),        # whole volumes, full tiles
    ((50, 37, 24), (50, 37, 24), (0, 0, 0)),        # ragged along all axes, rows narrower than the 32-wide box
    ((60, 52, 72), (41, 35, 40), (7, 9, 8)),        # crop window, 16-byte aligned start, two tiles + a ragged one along axis 2
    ((33, 18, 8), (33, 18, 8), (0, 0, 0)),          # one column group only
])
def test_tma_store_copy_tiles_are_bit_exact(shape, roi, start):
    """Plain copy items leave through the destination tensor map (ADELL_KIND_TSTORE): whole boxes, and plane by
    plane to the mirrored plane under an axis-0 flip; stores that reach past the item's extents are clipped by
    the map, never written into the neighbouring channel of the collated tensor.  Items that need the
    consumers (axis-1 / axis-2 flips, an intensity map) share the launch."""
    import ctypes as C

    from adell_mri_b200 import _lib, engine
    from adell_mri_b200.plan import ITEM_DTYPE

    R = np.random.RandomState(sum(shape))
    flips = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0], [0, 0, 1], [1, 0, 1], [0, 0, 0], [1, 0, 0]], bool)
    n = len(flips)
    scale = np.ones(n); scale[6] = 1.5
    vols = [torch.from_numpy(R.rand(*shape).astype(np.float32)) for _ in range(n)]
    plan = BatchPlan([v.to(DEV) for v in vols])
    plan.crop(np.array([start] * n), roi)
    plan.flip(flips)
    plan.intensity(scale=scale, offset=0.0)
    # destination: channel slices of one collated tensor, poisoned so that a stray store shows
    out = torch.full((2, n // 2, *roi), float("nan"), device=DEV)
    dsts = [out[i // (n // 2), i % (n // 2)] for i in range(n)]
    dst_ptr = np.array([d.data_ptr() for d in dsts], np.uint64)
    dst_stride = np.array([d.stride() for d in dsts], np.int64)
    items = plan.build_launches(dst_ptr, dst_stride, None)[-1]
    from adell_mri_b200.engine import aligned_items
    probe = aligned_items(n, items)
    tiles = np.zeros(n + 5, np.int32)
    info = _lib.LaunchInfo()
    _lib.check(_lib.load().adell_aug_plan(probe.ctypes.data, n, tiles.ctypes.data, C.byref(info)), "adell_aug_plan")
    kind = {int(p): int(k) for p, k in zip(probe["dst"], probe["kind"])}
    assert [kind[int(p)] for p in dst_ptr] == [3, 4, 2, 2, 2, 2, 2, 4]
    engine.execute(plan, dsts)
    torch.cuda.synchronize()
    got = out.cpu()
    assert not torch.isnan(got).any()
    for i, v in enumerate(vols):
        r = M.crop(v[None], start, roi)
        fl = [a for a in range(3) if flips[i, a]]
        r = (M.flip(r, fl) if fl else r)[0]
        if i == 6:
            r = r * np.float32(1.5)
        assert torch.equal(got[i // (n // 2), i % (n // 2)], r), i


@pytest.mark.parametrize("n_batches,seed,big", [(140, 2, False), (60, 5, False), (16, 1, True)])
def test_randomised_mixed_launches_against_the_oracle(n_batches, seed, big):
    """tools/fuzz_parity.py as a test: launches of up to 40 random items (random shapes, flips / crops / pads before
    and after, every mode x padding, strict and default plans, fp32 / int16 / uint8 sources, second resamples,
    intensity maps, noise, plain copies) against the torch oracle."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import fuzz_parity

    checked, bad = fuzz_parity.sweep(n_batches, seed, big, verbose=True)
    assert checked > 50 and bad == 0


@pytest.mark.parametrize("padding", ["border", "reflection"])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_reversed_axis_hi_tap_stays_inside_the_staged_box(padding, seed):
    """Found by the randomised sweep: on an axis that runs backwards in memory (a flip before the resample = a
    negative source stride) a coordinate clamped / reflected onto the first cell put the lo tap on the last box
    index and the hi tap one cell past the staged box.  Its weight is 0, but 0 * NaN = NaN: the stages are
    poisoned first by running the same items on an all-NaN source."""
    from adell_mri_b200 import engine

    R = np.random.RandomState(100 + seed)
    shape = (16, 40, 40)
    vols = [torch.from_numpy(R.rand(*shape).astype(np.float32)) for _ in range(4)]
    A = [rand_affine_matrix(R, rotate=(0.3, 0.3, 0.2), translate=(3, 6, 6), scale=(0.1, 0.1, 0.1)) for _ in vols]
    flips = np.array([[1, 1, 1], [1, 0, 0], [0, 1, 1], [1, 1, 0]], bool)

    def run(sources):
        plan = BatchPlan([v.to(DEV) for v in sources])
        plan.flip(flips)
        plan.affine(np.stack([a.numpy() for a in A]), "bilinear", padding)
        plan.spatial_pad((16, 44, 48))
        outs = [torch.full((16, 44, 48), float("nan"), device=DEV) for _ in sources]
        engine.execute(plan, outs)
        torch.cuda.synchronize()
        return [o.cpu() for o in outs]

    run([torch.full(shape, float("nan")) for _ in vols])
    got = run(vols)
    for i, v in enumerate(vols):
        fl = [a for a in range(3) if flips[i, a]]
        ref = M.spatial_pad(M.affine_resample(M.flip(v[None], fl), A[i], "bilinear", padding), (16, 44, 48))[0]
        assert not torch.isnan(got[i]).any(), (i, int(torch.isnan(got[i]).sum()))
        assert torch.allclose(got[i], ref, rtol=1e-4, atol=1e-4), (i, float((got[i] - ref).abs().max()))


@pytest.mark.parametrize("sizes", [[4, 4, 4, 4], [3, 1, 5, 2]])
def test_prepared_steps_equal_step_by_step_calls(sizes):
    """`SegmentationBatchAugmenter.prepare_steps` (what bench.py times: several steps drawn, composed and uploaded at
    once, launched one by one) gives exactly the batches that calling the augmenter step by step gives on the same
    seed — equal and unequal step sizes, launched in order and again in reverse."""
    from adell_mri_b200.pipelines import SegmentationBatchAugmenter

    R = np.random.RandomState(4)
    keys, shape = ["t2", "adc", "dwi"], (48, 40, 16)
    cache = [{k: torch.from_numpy(R.rand(1, *shape).astype(np.float32)).to(DEV) for k in keys + ["mask"]} for _ in range(6)]
    batches, i = [], 0
    for n in sizes:
        batches.append([cache[(i + j) % len(cache)] for j in range(n)]); i += n

    def make():
        aug = SegmentationBatchAugmenter(["affine", "flip"], keys + ["mask"], keys, flip_axis=[0, 1, 2])
        for s in aug.samplers:
            s.prob = 0.5
        return aug.set_random_state(9)

    a = make()
    want = [{k: v.clone() for k, v in a(b).items()} for b in batches]
    b_aug = make()
    outs = [{"image": torch.full((len(b), 3, *shape), float("nan"), device=DEV), "mask": torch.full((len(b), 1, *shape), float("nan"), device=DEV)}
            for b in batches]
    prepared = b_aug.prepare_steps(batches, outs)
    assert len(prepared) == len(sizes)
    for order in (range(len(sizes)), reversed(range(len(sizes)))):
        for o in outs:
            for t in o.values():
                t.fill_(float("nan"))
        for k in order:
            prepared.run(k)
        torch.cuda.synchronize()
        for k in range(len(sizes)):
            assert torch.equal(outs[k]["image"], want[k]["image"]), k
            assert torch.equal(outs[k]["mask"], want[k]["mask"]), k
    fired = sum(int(not torch.equal(want[k]["image"][j, 0], batches[k][j]["t2"][0])) for k in range(len(sizes)) for j in range(sizes[k]))
    assert fired >= 3


@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("seed", [0, 1])
def test_border_tiles_entirely_outside_on_a_reversed_axis(strict, seed):
    """Found by the randomised sweep (an illegal address): border padding, axes flipped BEFORE the resample
    (negative source strides) and tiles that lie entirely below the volume.  The coordinate clamps onto cell 0,
    whose hi tap, cell 1 (weight 0, still read), sat at box index -1 on the reversed axis."""
    from adell_mri_b200 import engine

    R = np.random.RandomState(200 + seed)
    shape = (72, 72, 64)
    vols = [torch.from_numpy(R.rand(*shape).astype(np.float32)) for _ in range(3)]
    A = [rand_affine_matrix(R, rotate=(0.35, 0.35, 0.2), translate=(24, 20, 10), scale=(0.1, 0.1, 0.1)) for _ in vols]
    flips = np.array([[1, 0, 1], [1, 1, 1], [0, 1, 0]], bool)
    plan = BatchPlan([v.to(DEV) for v in vols], strict=strict)
    plan.flip(flips)
    plan.affine(np.stack([a.numpy() for a in A]), "bilinear", "border")
    plan.center_crop((70, 64, 62))
    outs = [torch.full((70, 64, 62), float("nan"), device=DEV) for _ in vols]
    engine.execute(plan, outs)
    torch.cuda.synchronize()
    for i, v in enumerate(vols):
        fl = [a for a in range(3) if flips[i, a]]
        ref = M.center_spatial_crop(M.affine_resample(M.flip(v[None], fl), A[i], "bilinear", "border"), (70, 64, 62))[0]
        got = outs[i].cpu()
        if strict:
            assert torch.equal(got, ref), i
        else:
            assert torch.allclose(got, ref, rtol=1e-4, atol=1e-4), (i, float((got - ref).abs().max()))


@pytest.mark.parametrize("padding", ["zeros", "border", "reflection"])
@pytest.mark.parametrize("shape", [(128, 128, 32), (96, 80, 48)])
def test_strong_zooms_take_the_small_tile_shapes_and_match_the_oracle(shape, padding):
    """The workhorse's scale members draw factors around 2 (reference quirk) and fast mode
    multiplies them: footprints that the regular tile shapes cannot stage within the preferred box get the small
    shapes (8x16x16 ... 4x4x16) instead of a ~100 KB box that would leave the launch with one ring stage per stream."""
    import ctypes as C

    from adell_mri_b200 import _lib, engine

    R = np.random.RandomState(sum(shape))
    img = torch.from_numpy(R.rand(1, *shape).astype(np.float32))
    dev = img[0].to(DEV)
    seen = set()
    for trial, scale in enumerate([(2.0, 1.0, 1.0), (1.0, 2.1, 1.0), (3.1, 1.9, 1.1), (1.0, 1.0, 2.0), (2.0, 2.0, 2.0), (0.4, 0.5, 1.0)]):
        rot = R.uniform(-0.3, 0.3, 3)
        A = M.compose_affine(list(rot), None, list(R.uniform(-4, 4, 3)), list(scale))
        for mode in ("bilinear", "nearest"):
            plan = BatchPlan([dev]).affine(A.numpy(), mode, padding)
            out = torch.empty(shape, device=DEV)
            items = plan.build_launches(np.array([out.data_ptr()], np.uint64), np.array([out.stride()], np.int64), None)[-1]
            buf, n, info = engine.pack_launch(items)
            it = buf[: n * engine.ISZ].view(engine.ITEM_DTYPE)[0]
            seen.add((int(it["kind"]), tuple(int(x) for x in it["tile_dim"])))
            assert info.smem_bytes <= 56 * 1024, (scale, info.smem_bytes)       # never the single-stage regime
            ref = M.affine_resample(img, A, mode, padding)[0]
            got = run_plan_cuda(BatchPlan([dev]).affine(A.numpy(), mode, padding))[0].cpu()
            if mode == "nearest":
                assert mismatch(got, ref) == 0, (scale, padding)
            else:
                assert torch.allclose(got, ref, rtol=1e-4, atol=1e-4), (scale, padding, float((got - ref).abs().max()))
    assert any(k == 1 and t[0] * t[1] * t[2] < 16 * 16 * 16 for k, t in seen), seen     # some item did take a small shape
