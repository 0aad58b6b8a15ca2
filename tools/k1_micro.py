"""K1 launch-time microbenchmark (device-resident parameters, CUDA events on the launching stream).

  python tools/k1_micro.py [--prof] [workload ...]

Workloads: seg (reference probabilities), seg_all_affine, seg_copy (no affine ever fires: pure
flip copies), ssl (config C, two views, zeros padding), cls (config D, 208x208x64 -> 192x192x48),
a (config A: 512 x 1x128x128x32, the reference micro-benchmark's affine), e (config E: 4 x 512x512x128 with
device percentiles).
`--prof` loads libadell_b200_prof.so (built with -DK1_PROFILE) and prints its cycle counters.
"""
import ctypes
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from adell_mri_b200 import _lib

PROF = "--prof" in sys.argv
if PROF:
    _lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libadell_b200_prof.so")
from adell_mri_b200 import engine, geometry
from adell_mri_b200.plan import BatchPlan

dev = torch.device("cuda:0")
PEAK = 6541.1


def seg_items(R, prob, n_batches=4, B=8, shape=(256, 256, 32), integer=False, exact_k=None, drop=None):
    """integer=True: int16 image volumes + uint8 mask with a device-side {scale, offset} (config B from raw volumes)."""
    g = torch.Generator(device=dev).manual_seed(0)
    out_img = torch.empty((B, 3, *shape), device=dev)
    out_mask = torch.empty((B, 1, *shape), device=dev)
    launches = []
    keep = [out_img, out_mask]
    for _ in range(n_batches):
        vols, mats, fired, flips, dsts = [], [], [], [], []
        fired_b = R.permutation(B) < exact_k if exact_k is not None else None
        for b in range(B):
            f = R.rand() < prob
            if fired_b is not None:
                f = bool(fired_b[b])
            ang = R.uniform(-1, 1, 3) * np.array([np.pi / 8, np.pi / 8, np.pi / 16])
            A = geometry.compose_affine(rotate=ang[None])[0]
            fl = R.rand(3) < 0.25
            tstore = (not f) and not fl[1] and not fl[2]   # identity item that takes the TMA-store path
            if drop == "tstore" and tstore:
                continue
            if drop == "not_tstore" and not tstore:
                continue
            if drop == "vcopy" and not f and not tstore:
                continue
            for k in range(4):
                if not integer:
                    vols.append(torch.rand(shape, device=dev, generator=g))
                elif k < 3:
                    vols.append(torch.randint(0, 4000, shape, device=dev, generator=g, dtype=torch.int16))
                else:
                    vols.append((torch.rand(shape, device=dev, generator=g) > 0.7).to(torch.uint8))
                mats.append(A); fired.append(f); flips.append(fl)
                dsts.append(out_img[b, k] if k < 3 else out_mask[b, 0])
        if not vols:
            continue
        nb = len(vols) // 4
        plan = BatchPlan(vols)
        if integer:
            pre = torch.tensor([[1.0 / 4000.0, 0.0]] * 3 + [[1.0, 0.0]], device=dev).repeat(nb, 1).contiguous()
            plan.intensity_from_device(pre)
            keep.append(pre)
        plan.affine(np.stack(mats), (["bilinear"] * 3 + ["nearest"]) * nb, "reflection", where=np.array(fired))
        plan.flip(np.stack(flips))
        launches.append(pack(plan, dsts))
        keep.append(vols)
    if drop is not None:
        print("   voxel-channels per launch (mean over the batches):", sum(l[1] for l in launches) * int(np.prod(shape)) // len(launches))
    return launches, B * 4 * int(np.prod(shape)), keep


def ssl_items(R, n_batches=2, B=64, src=(160, 160, 40), roi=(128, 128, 32)):
    g = torch.Generator(device=dev).manual_seed(0)
    o1 = torch.empty((B, 1, *roi), device=dev)
    o2 = torch.empty((B, 1, *roi), device=dev)
    launches, keep = [], [o1, o2]
    for _ in range(n_batches):
        vols, mats, starts, dsts = [], [], [], []
        for b in range(B):
            v = torch.rand(src, device=dev, generator=g)
            st = [R.randint(s - r + 1) for s, r in zip(src, roi)]
            if os.environ.get('SSL_ALIGNED'):
                st[2] &= ~3
            for view in range(2):
                rot = np.zeros(3); rot[R.randint(3)] = R.uniform(-1, 1) * (np.pi / 12)
                tr = np.zeros(3); tr[R.randint(3)] = R.uniform(-15, 15)
                A = geometry.compose_affine(rotate=rot[None], translate=tr[None])[0]
                vols.append(v); mats.append(A); starts.append(st)
                dsts.append((o1 if view == 0 else o2)[b, 0])
        plan = BatchPlan(vols)
        plan.crop(np.array(starts), roi)
        plan.affine(np.stack(mats), "bilinear", "zeros")
        plan.intensity(scale=1.1, offset=0.05)
        launches.append(pack(plan, dsts))
        keep.append(vols)
    return launches, B * 2 * int(np.prod(roi)), keep


def cls_items(R, n_batches=2, B=32, src=(208, 208, 64), crop=(192, 192, 48), K=3):
    g = torch.Generator(device=dev).manual_seed(0)
    out = torch.empty((B, K + 1, *crop), device=dev)
    launches, keep = [], [out]
    for _ in range(n_batches):
        vols, mats, flips, dsts, modes = [], [], [], [], []
        for b in range(B):
            rot = np.array([R.uniform(-1, 1) * np.pi / 16])
            tr = R.uniform(-1, 1, 3) * np.array([4, 4, 1])
            sc = 1 + R.uniform(-1, 1, 3) * np.array([0.1, 0.1, 0.05])
            A = geometry.compose_affine(rotate=rot[None], translate=tr[None], scale=sc[None])[0]
            fl = R.rand(3) < 0.3
            for k in range(K + 1):
                vols.append(torch.rand(src, device=dev, generator=g))
                mats.append(A); flips.append(fl); dsts.append(out[b, k])
                modes.append("bilinear" if k < K else "nearest")
        plan = BatchPlan(vols)
        plan.flip(np.stack(flips))
        plan.affine(np.stack(mats), modes, "zeros")
        plan.center_crop(crop)
        launches.append(pack(plan, dsts))
        keep.append(vols)
    return launches, B * (K + 1) * int(np.prod(crop)), keep


def a_items(R, n_batches=2, N=512, shape=(128, 128, 32)):
    """Config A (the reference's benchmarks/benchmark-random-affine.py case, batched): RandAffined prob 1,
    rotate +-pi/6 x3, translate [10,10,3], scale +-0.1, trilinear, reflection."""
    g = torch.Generator(device=dev).manual_seed(0)
    out = torch.empty((N, 1, *shape), device=dev)
    src = [torch.rand(shape, device=dev, generator=g) for _ in range(64)]
    launches, keep = [], [out, src]
    for _ in range(n_batches):
        vols, mats, dsts = [], [], []
        for b in range(N):
            rot = R.uniform(-1, 1, 3) * (np.pi / 6)
            tr = R.uniform(-1, 1, 3) * np.array([10, 10, 3])
            sc = 1 + R.uniform(-0.1, 0.1, 3)
            vols.append(src[b % 64]); dsts.append(out[b, 0])
            mats.append(geometry.compose_affine(rotate=rot[None], translate=tr[None], scale=sc[None])[0])
        plan = BatchPlan(vols)
        plan.affine(np.stack(mats), "bilinear", "reflection")
        launches.append(pack(plan, dsts))
    return launches, N * int(np.prod(shape)), keep


def e_items(R, n_batches=2, M=4, shape=(512, 512, 128)):
    """Config E (large volumes): M volumes of 512x512x128 lognormal fp32, percentile (1, 99) scaling read
    from the device ({scale, offset} through pre_dev) + affine gather.  The statistics are timed apart."""
    from adell_mri_b200 import _lib, stats
    g = torch.Generator(device=dev).manual_seed(0)
    vols = [torch.empty(shape, device=dev).log_normal_(0, 1, generator=g) for _ in range(M)]
    out = torch.empty((M, 1, *shape), device=dev)
    flat = [v.reshape(-1) for v in vols]

    def pre():
        pct = stats.percentiles(flat, [1.0, 99.0])
        return stats.coefs_to_affine(stats.scaler_coefs(pct, _lib.SCALER_RANGE, 0.0, 1.0))

    for _ in range(2):
        pre_dev = pre()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        pre_dev = pre()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    nvox = M * int(np.prod(shape))
    print(f"   config E statistics (K2/K3 exact percentiles 1/99 of {M} x {shape}): {ms:.3f} ms = {4.0 * nvox / ms / 1e6:.0f} GB/s "
          f"(4 B per voxel credited once) = {4.0 * nvox / ms / 1e6 / PEAK:.3f} of measured peak")
    launches, keep = [], [out, vols, pre_dev]
    for _ in range(n_batches):
        mats = []
        for b in range(M):
            rot = R.uniform(-1, 1, 3) * np.array([np.pi / 8, np.pi / 8, np.pi / 16])
            mats.append(geometry.compose_affine(rotate=rot[None])[0])
        plan = BatchPlan(vols)
        plan.intensity_from_device(pre_dev)
        plan.affine(np.stack(mats), "bilinear", "zeros")
        launches.append(pack(plan, [out[b, 0] for b in range(M)]))
    return launches, nvox, keep


def pack(plan, dsts):
    dst_ptr = np.array([d.data_ptr() for d in dsts], np.uint64)
    dst_stride = np.array([d.stride() for d in dsts], np.int64)
    items = plan.build_launches(dst_ptr, dst_stride, None)[-1]
    buf, n, info = engine.pack_launch(items)
    it = buf[: n * engine.ISZ].view(engine.ITEM_DTYPE)
    STATS.append((it["tile_dim"].copy(), it["tmap_box"].copy(), it["n_tiles"].copy(), it["kind"].copy(), it["out_shape"].copy()))
    return torch.from_numpy(buf).to(dev), n, info, plan


STATS = []


def tile_stats():
    """Host policy summary of the packed launches: tile shapes, staged bytes per output voxel."""
    td = np.concatenate([s[0] for s in STATS]); box = np.concatenate([s[1] for s in STATS]).astype(np.int64)
    nt = np.concatenate([s[2] for s in STATS]).astype(np.int64); kind = np.concatenate([s[3] for s in STATS])
    osh = np.concatenate([s[4] for s in STATS]).astype(np.int64)
    out = []
    for k in np.unique(kind):
        for shp in np.unique(td[kind == k], axis=0):
            sel = (kind == k) & (td == shp).all(axis=1)
            tiles = nt[sel].prod(axis=1).sum()
            bb = (box[sel].prod(axis=1) * 4 * nt[sel].prod(axis=1)).sum()
            out.append(f"kind {k} tile {tuple(int(x) for x in shp)}: {int(sel.sum())} items, {int(tiles)} tiles, mean box {bb / max(tiles, 1) / 1024:.1f} KiB, "
                       f"{bb / osh[sel].prod(axis=1).sum():.1f} staged B/voxel")
    STATS.clear()
    return "; ".join(out)


def time_launches(launches, reps=5):
    st = torch.cuda.current_stream()
    for (buf, n, info, _) in launches:
        engine.launch_packed(buf, n, info)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        for (buf, n, info, _) in launches:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st); engine.launch_packed(buf, n, info); b.record(st)
            ts.append((a, b))
    torch.cuda.synchronize()
    return [a.elapsed_time(b) for a, b in ts]


def main():
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or ["seg", "seg_all_affine", "seg_copy", "ssl", "cls"]
    lib = _lib.load()
    if PROF:
        lib.adell_debug_prof.argtypes = [ctypes.c_void_p, ctypes.c_int]
    for name in names:
        R = np.random.RandomState(7)
        if name == "seg":
            L, vox, keep = seg_items(R, 0.2)
        elif name == "seg_all_affine":
            L, vox, keep = seg_items(R, 1.0)
        elif name == "seg_copy":
            L, vox, keep = seg_items(R, 0.0)
        elif name.startswith("seg_k"):   # exactly k of the 8 samples of every batch resampled; suffix _rv / _t: without / only the TMA-store samples
            kk, _, sfx = name[5:].partition("_")
            L, vox, keep = seg_items(R, 0.0, exact_k=int(kk), drop={"": None, "rv": "tstore", "t": "not_tstore", "rt": "vcopy"}[sfx])
        elif name in ("seg_i16", "seg_i16_all", "seg_i16_copy"):
            L, vox, keep = seg_items(R, {"seg_i16": 0.2, "seg_i16_all": 1.0, "seg_i16_copy": 0.0}[name], integer=True)
        elif name == "ssl":
            L, vox, keep = ssl_items(R)
        elif name == "cls":
            L, vox, keep = cls_items(R)
        elif name == "a":
            L, vox, keep = a_items(R)
        elif name == "e":
            L, vox, keep = e_items(R)
        else:
            raise SystemExit(name)
        print("  ", tile_stats())
        if PROF:
            out = (ctypes.c_ulonglong * 16)()
            lib.adell_debug_prof(out, 1)
            lib.adell_debug_prof2.argtypes = [ctypes.c_void_p, ctypes.c_int]
            lib.adell_debug_prof2((ctypes.c_ulonglong * 16)(), 1)
        ts = time_launches(L)
        ms = statistics.mean(ts)
        gbs = 8.0 * vox / (ms * 1e-3) / 1e9
        print(f"{name}: {ms:.4f} ms/launch (min {min(ts):.4f}), {vox/ms/1e6:.1f} Gvox/s, {gbs:.0f} GB/s algorithmic = {gbs/PEAK:.3f} of measured peak; "
              f"staged items {L[0][2].n_staged}/{L[0][1]}, smem {L[0][2].smem_bytes}")
        if PROF:
            lib.adell_debug_prof(out, 1)
            v = list(out)
            print("   cycles summed over warps: prod wait-empty %d, issue %d, prepare %d | cons wait-full %d, compute %d" % tuple(v[:5]))
            nt = max(v[7], 1)
            print("   per tile: producer prepare %.0f, wait-empty %.0f, issue %.0f cycles | per consumer warp: wait-full %.0f, compute %.0f cycles; %d tiles"
                  % (v[2] / nt, v[0] / nt, v[1] / nt, v[3] / nt / 8, v[4] / nt / 8, nt))
            for lbl, k in (("resampled", 8), ("consumer-copy", 10), ("TMA-store", 12), ("other", 14)):
                if v[k + 1]:
                    print("   %s tiles: %d, %.0f cycles per tile (first warp of the group)" % (lbl, v[k + 1], v[k] / v[k + 1]))
            lib.adell_debug_prof2.argtypes = [ctypes.c_void_p, ctypes.c_int]
            o2 = (ctypes.c_ulonglong * 16)()
            lib.adell_debug_prof2(o2, 1)
            w = list(o2)
            print("   producer phases, cycles per tile (lane 0): origin %.0f | column groups %.0f | footprint / padding cases / fit %.0f | votes, box, "
                  "fast origin %.0f | shuffles %.0f | stores %.0f || prefix walk %.0f | item fetch + context copies %.0f | tile coordinates %.0f"
                  % tuple(x / nt for x in (w[0], w[1], w[2], w[3], w[4], w[5], w[8], w[9], w[10])))
            nl = len(ts)
            print("   consumer-group kernel time: mean %.0f cycles per launch, max over groups and launches %d cycles; event time %.0f cycles at 1965 MHz"
                  % (v[5] / (nl * 148 * 2), v[6], ms * 1e-3 * 1965e6))
        del L, keep
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
