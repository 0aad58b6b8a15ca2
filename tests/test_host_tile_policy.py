"""Host policy of K1 without a GPU or a driver (`adell_aug_plan`): item order, tile shapes, the
column-group shear of the tile grid and the staged boxes.

* partition: with the shifts in `adell_item.shear`, every output voxel belongs to exactly one
  (tile, thread) of the item's tile grid;
* coverage: for every tile, the taps of every voxel of the tile (coordinates from the fp64 map
  `fp_U0 / fp_D` the device uses, zero padding) fall inside a window of the encoded box extents;
* queues: identity items are moved behind the others and `first_copy_tile` splits the tile range.
"""

import ctypes as C

import numpy as np
import pytest
import torch

from adell_mri_b200 import _lib, geometry
from adell_mri_b200.engine import aligned_items
from adell_mri_b200.plan import ITEM_DTYPE, BatchPlan


def plan_items(plan, out_shapes):
    outs = [torch.empty(tuple(int(x) for x in s), dtype=torch.float32) for s in out_shapes]
    dst_ptr = np.array([o.data_ptr() for o in outs], np.uint64)
    dst_stride = np.array([o.stride() for o in outs], np.int64)
    items = plan.build_launches(dst_ptr, dst_stride, None)[-1]
    n = items.shape[0]
    it = aligned_items(n, items)
    tiles = np.zeros(n + 5, np.int32)
    info = _lib.LaunchInfo()
    _lib.check(_lib.load().adell_aug_plan(it.ctypes.data, n, tiles.ctypes.data, C.byref(info)), "adell_aug_plan")
    return it, tiles, info, outs


def rand_affine(R, big=False):
    ang = R.uniform(-1, 1, 3) * (np.array([np.pi / 6] * 3) if big else np.array([np.pi / 8, np.pi / 8, np.pi / 16]))
    tr = R.uniform(-1, 1, 3) * np.array([6, 6, 2])
    sc = 1 + R.uniform(-0.1, 0.1, 3)
    return geometry.compose_affine(rotate=ang[None], translate=tr[None], scale=sc[None])[0]


@pytest.mark.parametrize("shape,padding", [((128, 96, 32), "reflection"), ((40, 56, 48), "zeros"), ((36, 48, 20), "zeros"), ((96, 80, 32), "border"), ((72, 72, 72), "border")])
def test_sheared_tile_grid_partitions_the_output_and_boxes_cover_the_footprints(shape, padding):
    R = np.random.RandomState(sum(shape))
    vols = [torch.zeros(shape) for _ in range(6)]
    plan = BatchPlan(vols)
    plan.affine(np.stack([rand_affine(R, big=(i % 2 == 1)) for i in range(6)]), "bilinear", padding)
    plan.flip(np.array([[i & 1, i & 2, i & 4] for i in range(6)], bool))
    it, tiles, info, _ = plan_items(plan, [shape] * 6)
    assert info.n_staged < 0 and info.first_copy_tile == info.total_tiles      # a plan is not launchable; no copy items
    assert -1 - info.n_staged == 6                                              # every item takes the staged path
    for item in it:
        assert not (int(item["flags"]) & _lib.F_TMAP)
        if int(item["kind"]) != 1:    # ADELL_KIND_STAGED
            continue
        T = [int(x) for x in item["tile_dim"]]
        O = [int(x) for x in item["out_shape"]]
        nt = [int(x) for x in item["n_tiles"]]
        sh = np.asarray(item["shear"]).reshape(2, 16).astype(np.int64)
        assert (sh >= 0).all()
        # ---- partition: voxel (i, j, k) -> tile ((i + s0[G]) // T0, (j + s1[G]) // T1, k // T2), G = k >> 3
        k = np.arange(O[2])
        G = (k >> 3) & 15
        for a in range(2):
            b = (np.arange(O[a])[:, None] + sh[a][G][None, :]) // T[a]
            assert b.min() >= 0 and b.max() < nt[a], (a, T, nt)
        assert (O[2] - 1) // T[2] < nt[2]
        # ---- coverage: per tile, the span of the un-padded source coordinates of its voxels fits the box
        U0, D = np.asarray(item["fp_U0"], np.float64), np.asarray(item["fp_D"], np.float64).reshape(3, 3)
        box = [int(x) for x in item["tmap_box"]]
        S = [int(x) for x in item["src_shape"]]
        ii, jj, kk = np.meshgrid(np.arange(O[0]), np.arange(O[1]), np.arange(O[2]), indexing="ij")
        Gv = (kk >> 3) & 15
        b0, b1, b2 = (ii + sh[0][Gv]) // T[0], (jj + sh[1][Gv]) // T[1], kk // T[2]
        tid = (b0 * nt[1] + b1) * nt[2] + b2
        for a in range(3):
            u = U0[a] + D[a, 0] * ii + D[a, 1] * jj + D[a, 2] * kk
            lo, hi = np.floor(u), np.floor(u) + 1                    # tap indices
            if padding != "zeros" and box[a] >= S[a]:
                continue                                                # axis staged whole (thin, or the footprint spans it)
            # border / reflection fold the coordinates with a 1-Lipschitz map: the folded taps of a tile
            # span no more cells than the unfolded ones checked here
            order = np.argsort(tid.ravel())
            t_sorted = tid.ravel()[order]
            starts = np.concatenate([[0], np.nonzero(np.diff(t_sorted))[0] + 1])
            tmin = np.minimum.reduceat(lo.ravel()[order], starts)
            tmax = np.maximum.reduceat(hi.ravel()[order], starts)
            need = (tmax - tmin + 1).max() + (3 if a == 2 else 0)      # + worst alignment of the box origin
            assert need <= box[a], (a, need, box, T)


def test_copy_items_go_last_and_split_the_tile_range():
    R = np.random.RandomState(1)
    shape = (32, 32, 32)
    vols = [torch.zeros(shape) for _ in range(7)]
    fired = np.array([False, True, False, False, True, False, True])
    plan = BatchPlan(vols)
    plan.affine(np.stack([rand_affine(R) for _ in range(7)]), "bilinear", "zeros", where=fired)
    it, tiles, info, outs = plan_items(plan, [shape] * 7)
    kinds = [int(k) for k in it["kind"]]
    assert kinds == sorted(kinds, key=lambda k: k == 2)            # VCOPY (2) items behind the others, stable
    dst_order = [int(p) for p in it["dst"]]
    want = [outs[i].data_ptr() for i in np.nonzero(fired)[0]] + [outs[i].data_ptr() for i in np.nonzero(~fired)[0]]
    assert dst_order == want
    n_res = int(fired.sum())
    assert info.first_copy_tile == tiles[n_res] and tiles[7] == info.total_tiles
    assert list(tiles[8:12]) == [0, 0, 0, 0]
    assert (np.diff(tiles[:8]) > 0).all()


def test_a_plan_cannot_be_launched():
    plan = BatchPlan([torch.zeros(16, 16, 16)])
    it, tiles, info, _ = plan_items(plan, [(16, 16, 16)])
    assert _lib.load().adell_aug_gather(64, 64, 1, C.byref(info), None) == -1


def test_plain_copy_items_leave_through_the_destination_tensor_map():
    """Identity items without an intensity map and without a flip along the contiguous axis become
    ADELL_KIND_TSTORE + split (0 whole boxes, 1 plane by plane: axis 0 flipped); a flip along axis 1 would
    need row-by-row stores (split 2: measured slower than the consumer copy, off by default) and, like
    everything else, keeps the vectorised consumer copy (ADELL_KIND_VCOPY)."""
    shape = (48, 40, 32)
    flips = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0], [0, 0, 1], [1, 1, 1], [0, 0, 0], [0, 0, 0]], bool)
    vols = [torch.zeros(shape) for _ in range(len(flips))]
    plan = BatchPlan(vols)
    plan.flip(flips)
    scale = np.ones(len(flips)); scale[6] = 1.25
    plan.intensity(scale=scale, offset=0.0)
    outs_shapes = [shape] * len(flips)
    it, tiles, info, outs = plan_items(plan, outs_shapes)
    by_dst = {int(p): int(k) for p, k in zip(it["dst"], it["kind"])}
    kinds = [by_dst[o.data_ptr()] for o in outs]
    assert kinds[:6] == [3, 4, 2, 2, 2, 2]
    assert kinds[6] == 2            # an intensity map needs the consumers
    assert kinds[7] == 3
    assert info.first_copy_tile == 0 and info.total_tiles == int(tiles[len(flips)])   # all of them feed the copy queue


def test_unaligned_crop_windows_keep_the_consumer_copy():
    shape, roi = (40, 40, 40), (32, 32, 32)
    vols = [torch.zeros(shape) for _ in range(3)]
    plan = BatchPlan(vols)
    plan.crop(np.array([[3, 5, 0], [3, 5, 4], [3, 5, 2]]), roi)
    it, tiles, info, outs = plan_items(plan, [roi] * 3)
    by_dst = {int(p): int(k) for p, k in zip(it["dst"], it["kind"])}
    assert [by_dst[o.data_ptr()] for o in outs] == [3, 3, 2]   # a window start that is not 16-byte aligned carries slack columns
