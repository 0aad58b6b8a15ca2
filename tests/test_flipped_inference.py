"""FlippedInference (reference utils/inference.py:303-392): the mirror's structure
handling against the reference semantics (torch.flip / np.flip of copies) on the CPU, and the K1 flip path against
torch.flip on the GPU (bit-exact, one launch per flip)."""

import numpy as np
import pytest
import torch

from adell_mri_b200.inference import FlippedInference, flip_volumes


def _ref_call(fn, X, flips):
    """The reference's __call__ restated with torch.flip (tensor input)."""
    out = fn(X)
    for f in flips:
        out = out + torch.flip(fn(torch.flip(X, tuple(f))), tuple(f))
    return out / (len(flips) + 1)


def test_structures_and_flip_keys_follow_the_reference_on_cpu():
    R = np.random.RandomState(0)
    x = torch.from_numpy(R.rand(2, 1, 6, 5, 4).astype(np.float32))
    a = R.rand(2, 1, 6, 5, 4)
    fi = FlippedInference(lambda v: v, flips=[[2], [3, 4]], flip_keys=["image"])
    d = fi.flip({"image": x, "other": a}, [2, 4])
    assert torch.equal(d["image"], torch.flip(x, (2, 4))) and np.array_equal(d["other"], a) and d["other"] is not a
    fi2 = FlippedInference(lambda v: v, flips=[[2]])
    l = fi2.flip([x, a], [3])
    assert torch.equal(l[0], torch.flip(x, (3,))) and np.array_equal(l[1], np.flip(a, 3))
    assert isinstance(fi2.flip((x, a), [3]), tuple)
    with pytest.raises(NotImplementedError):
        fi2.flip("nope", [2])
    fn = lambda v: v * 2 + torch.arange(4.0)
    assert torch.allclose(FlippedInference(fn, [[2], [3], [2, 3, 4]])(x.clone()), _ref_call(fn, x, [[2], [3], [2, 3, 4]]))


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 3, 64, 48, 32), (1, 2, 40, 36, 20), (4, 33, 17, 12)])
def test_k1_flips_equal_torch_flip_bit_for_bit(shape):
    from adell_mri_b200 import engine

    x = torch.rand(shape, device="cuda:0")
    nd = x.dim()
    for axes in ([nd - 3], [nd - 2], [nd - 1], [nd - 3, nd - 1], [nd - 3, nd - 2, nd - 1], [-1, -2]):
        before = engine.launch_count
        got = flip_volumes(x, axes)
        assert engine.launch_count - before == 1
        assert torch.equal(got, torch.flip(x, tuple(axes)))


@pytest.mark.gpu
def test_flipped_inference_on_the_device_equals_the_reference_recipe():
    x = torch.rand(2, 3, 32, 32, 16, device="cuda:0")
    w = torch.rand(3, device="cuda:0").view(1, 3, 1, 1, 1)
    fn = lambda v: torch.sigmoid(v * w + v.roll(1, 2))      # not flip-equivariant: the un-flip matters
    flips = [[2], [3], [4], [2, 3], [2, 3, 4]]
    got = FlippedInference(fn, flips)(x)
    assert torch.allclose(got, _ref_call(fn, x, flips), rtol=0, atol=1e-6)
    # channel flips are not voxel flips: they go through torch.flip like the reference
    assert torch.equal(FlippedInference(fn, [[1]]).flip(x, [1]), torch.flip(x, (1,)))


def _ref_sliding(X, fn, window, stride, n_classes, ibs):
    """The reference's SlidingWindowSegmentation.__call__ restated for one batched tensor [B, C, H, W, D]."""
    sh = X.shape[-3:]
    out = torch.zeros(X.shape[0], n_classes, *sh, device=X.device)
    den = torch.zeros_like(out)
    coords = []
    for i in range(0, sh[0], stride[0]):
        for j in range(0, sh[1], stride[1]):
            for k in range(0, sh[2], stride[2]):
                c = []
                for a, s in zip((i, j, k), range(3)):
                    x1, x2 = a, a + window[s]
                    if x2 > sh[s]:
                        x1, x2 = sh[s] - window[s], sh[s]
                    c.append((x1, x2))
                coords.append(tuple(c))
    for b0 in range(0, len(coords), ibs):
        cc = coords[b0:b0 + ibs]
        batch = torch.cat([X[..., c[0][0]:c[0][1], c[1][0]:c[1][1], c[2][0]:c[2][1]] for c in cc], 0)
        res = torch.split(fn(batch), X.shape[0], 0)
        for r, c in zip(res, cc):
            out[..., c[0][0]:c[0][1], c[1][0]:c[1][1], c[2][0]:c[2][1]] += r
            den[..., c[0][0]:c[0][1], c[1][0]:c[1][1], c[2][0]:c[2][1]] += 1.0
    return out / den


def test_sliding_window_matches_the_reference_recipe_on_cpu():
    from adell_mri_b200.inference import SlidingWindowSegmentation

    x = torch.rand(2, 3, 20, 18, 10)
    fn = lambda b: b[:, :2] * 2 + b[:, 2:3]
    sw = SlidingWindowSegmentation([8, 8, 4], fn, n_classes=2, stride=[6, 5, 3], inference_batch_size=4)
    assert torch.allclose(sw(x), _ref_sliding(x, fn, [8, 8, 4], [6, 5, 3], 2, 4))
    # a dict of inputs, unbatched arrays
    fn_d = lambda b: b["a"][:, :1] + torch.as_tensor(b["b"][:, :1])
    got = SlidingWindowSegmentation([8, 8, 4], fn_d, n_classes=1, inference_batch_size=2)({"a": x[0], "b": x[1]})
    assert got.shape == (1, 20, 18, 10)


@pytest.mark.gpu
def test_sliding_window_gathers_each_inference_batch_with_one_k1_launch():
    from adell_mri_b200 import engine
    from adell_mri_b200.inference import SlidingWindowSegmentation, gather_windows

    x = torch.rand(2, 3, 40, 36, 24, device="cuda:0")
    coords = [((0, 16), (4, 20), (8, 16)), ((24, 40), (20, 36), (16, 24)), ((5, 21), (1, 17), (3, 11))]
    before = engine.launch_count
    got = gather_windows(x, coords, batched=True)
    assert engine.launch_count - before == 1
    want = torch.cat([x[..., c[0][0]:c[0][1], c[1][0]:c[1][1], c[2][0]:c[2][1]] for c in coords], 0)
    assert torch.equal(got, want)
    assert torch.equal(gather_windows(x[0], coords, batched=False),
                       torch.stack([x[0][..., c[0][0]:c[0][1], c[1][0]:c[1][1], c[2][0]:c[2][1]] for c in coords], 0))
    fn = lambda b: torch.sigmoid(b[:, :2] * 3 - b[:, 2:3])
    sw = SlidingWindowSegmentation([16, 16, 8], fn, n_classes=2, stride=[12, 12, 8], inference_batch_size=5)
    assert torch.allclose(sw(x), _ref_sliding(x, fn, [16, 16, 8], [12, 12, 8], 2, 5), rtol=0, atol=1e-6)
