"""Randomised parity sweep of K1 on the GPU against the torch oracle (oracle/monai_restated.py): batches of
random items — random shapes, pre / post flips, crops, pads, affines of every mode x padding, intensity maps,
plain copies (TMA-store and consumer paths), nearest masks — all in ONE launch per batch, so that item
switches, queue order and tile kinds are mixed the way no hand-written case mixes them.

    python tools/fuzz_parity.py [n_batches] [seed] [big]

Exact items (integer work, nearest, everything of a strict plan) must be bit-equal; trilinear items within
rtol 1e-4 / atol 1e-4 of the value range (default fast-coordinate mode).  Sources are fp32, int16 or uint8; some
items carry a second resample (one more pass) or an injected noise volume.  Test infrastructure: imports oracle/."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from adell_mri_b200 import engine
from adell_mri_b200.plan import BatchPlan
from oracle import monai_restated as M
from tests.helpers import rand_affine_matrix

DEV = "cuda:0"


BIG = False


def random_item(R):
    if BIG:   # volumes of several tiles per axis: the staged (TMA) paths, sheared tile grids, ragged edges
        shape = tuple(int(R.choice([64, 72, 80, 96, 112, 128, 160])) for _ in range(2)) + (int(R.choice([16, 24, 32, 48, 64])),)
    else:
        shape = tuple(int(R.choice([8, 12, 16, 20, 24, 32, 33, 40, 48, 56])) for _ in range(2)) + (int(R.choice([8, 12, 16, 24, 32, 40])),)
    desc = [f"shape={shape}"]
    dt = R.rand()
    if dt < 0.12:
        img = torch.from_numpy(R.randint(-200, 4096, size=(1, *shape)).astype(np.int16))
    elif dt < 0.18:
        img = torch.from_numpy(R.randint(0, 256, size=(1, *shape)).astype(np.uint8))
    else:
        img = torch.from_numpy(R.rand(1, *shape).astype(np.float32))
    strict = R.rand() < 0.25     # ATen operation order: the trilinear sum is then bit-exact as well
    dev_img = img[0].to(DEV)
    view = R.rand()
    desc.append(f"dtype={img.dtype} strict={strict} view={view:.2f}")
    if view < 0.3:               # a window of a larger allocation: offsets on every axis, rows that start mid-line
        pad = [int(R.randint(0, 4)) for _ in range(3)]
        big = torch.zeros(tuple(s + 2 * p for s, p in zip(shape, pad)), dtype=img.dtype, device=DEV)
        desc.append(f"pad={pad}")
        big[pad[0]:pad[0] + shape[0], pad[1]:pad[1] + shape[1], pad[2]:pad[2] + shape[2]] = dev_img
        dev_img = big[pad[0]:pad[0] + shape[0], pad[1]:pad[1] + shape[1], pad[2]:pad[2] + shape[2]]
    elif view < 0.4:             # every second element along one axis (a non-unit stride on axis 2 takes the generic path)
        ax = int(R.randint(3))
        desc.append(f"step2 axis={ax}")
        big_shape = list(shape); big_shape[ax] *= 2
        big = torch.zeros(tuple(big_shape), dtype=img.dtype, device=DEV)
        sl = [slice(None)] * 3; sl[ax] = slice(None, None, 2)
        big[tuple(sl)] = dev_img
        dev_img = big[tuple(sl)]
    plan = BatchPlan([dev_img], strict=strict)
    ref = img
    exact = True
    scale_of_values = 4096.0 if dt < 0.12 else (255.0 if dt < 0.18 else 1.0)
    cur = list(shape)
    # integer ops before the resample
    if R.rand() < 0.4:
        ax = [a for a in range(3) if R.rand() < 0.4]
        if ax:
            plan.flip(np.array([a in ax for a in range(3)])); ref = M.flip(ref, ax); desc.append(f"preflip={ax}")
    if R.rand() < 0.3:
        roi = [max(4, (c - int(R.randint(0, 9))) // 4 * 4 if a == 2 else c - int(R.randint(0, 9))) for a, c in enumerate(cur)]
        roi = [min(r, c) for r, c in zip(roi, cur)]
        st = [int(R.randint(0, c - r + 1)) for c, r in zip(cur, roi)]
        plan.crop(st, roi); ref = M.crop(ref, st, roi); cur = roi; desc.append(f"crop={st}+{roi}")
    if R.rand() < 0.25:          # an intensity map BEFORE the resample: zero-padded taps must stay 0, not the offset
        s0, o0 = float(np.float32(R.uniform(0.5, 1.5))), float(np.float32(R.uniform(-0.3, 0.3)))
        plan.intensity(scale=s0, offset=o0)
        ref = ref.to(torch.float32) * torch.tensor(s0, dtype=torch.float32) + torch.tensor(o0, dtype=torch.float32)
        exact = False
        desc.append("pre-intensity")
    kind = R.rand()
    if kind < 0.55:
        mode = "nearest" if R.rand() < 0.35 else "bilinear"
        padding = str(R.choice(["zeros", "border", "reflection"]))
        A = rand_affine_matrix(R, rotate=(0.4, 0.4, 0.2), translate=(4, 4, 2), scale=(0.1, 0.1, 0.1))
        plan.affine(A.numpy(), mode, padding); ref = M.affine_resample(ref, A, mode, padding); desc.append(f"affine {mode} {padding}")
        exact = exact and (mode == "nearest" or strict)
        if R.rand() < 0.15:      # a second resample (the reference's shear after the affine): one more pass
            A2 = rand_affine_matrix(R, rotate=(0.1, 0.1, 0.1), translate=(1, 1, 1), scale=(0.05, 0.05, 0.05))
            plan.affine(A2.numpy(), mode, padding); ref = M.affine_resample(ref, A2, mode, padding)
    # integer ops after it
    if R.rand() < 0.4:
        ax = [a for a in range(3) if R.rand() < 0.4]
        if ax:
            plan.flip(np.array([a in ax for a in range(3)])); ref = M.flip(ref, ax)
    if R.rand() < 0.3:
        size = [c + int(R.choice([0, 0, 3, 4, 8])) for c in cur]
        size[2] = (size[2] + 3) // 4 * 4
        plan.spatial_pad(size); ref = M.spatial_pad(ref, size); cur = list(ref.shape[1:])
    if R.rand() < 0.3:
        roi = [max(4, c - int(R.choice([0, 2, 4, 8]))) for c in cur]
        plan.center_crop(roi); ref = M.center_spatial_crop(ref, roi); cur = list(ref.shape[1:])
    if R.rand() < 0.3:
        s, o = float(np.float32(R.uniform(0.5, 1.5))), float(np.float32(R.uniform(-0.2, 0.2)))
        plan.intensity(scale=s, offset=o)
        ref = ref.to(torch.float32) * torch.tensor(s, dtype=torch.float32) + torch.tensor(o, dtype=torch.float32)
        exact = False
    if R.rand() < 0.1:           # injected noise (RandGaussianNoised parity path)
        nz = torch.from_numpy(R.normal(0, 0.1, size=tuple(ref.shape[1:])).astype(np.float32))
        plan.add_noise([nz.to(DEV)])
        ref = ref.to(torch.float32) + nz
    desc.append(f"out={tuple(ref.shape[1:])}")
    return plan, ref[0].to(torch.float32), exact, scale_of_values, " ".join(desc)


_POISON = None


def poison_stages():
    """Fill the kernel's shared-memory stages with NaN before every launch (copies and whole-volume boxes of all-NaN
    sources): a tap that is read outside its staged box — even with weight 0 — then shows up as NaN in the output."""
    global _POISON
    if _POISON is None:
        nan = float("nan")
        vols = [torch.full((64, 64, 64), nan, device=DEV), torch.full((16, 40, 40), nan, device=DEV), torch.full((128, 128, 32), nan, device=DEV)]
        R = np.random.RandomState(0)
        plan = BatchPlan(vols)
        A = np.stack([rand_affine_matrix(R, rotate=(0.3, 0.3, 0.2)).numpy() for _ in vols])
        plan.affine(A, "bilinear", "reflection", where=np.array([True, True, False]))
        _POISON = (plan, [torch.empty(tuple(v.shape), device=DEV) for v in vols])
    engine.execute(_POISON[0], _POISON[1])


def sweep(n_batches, seed, big=False, verbose=True):
    """Returns (items checked, mismatches)."""
    global BIG
    BIG = big
    R = np.random.RandomState(seed)
    bad = checked = 0
    for b in range(n_batches):
        items = [random_item(R) for _ in range(int(R.randint(1, 12 if BIG else 40)))]
        outs = [torch.full(tuple(it[1].shape), float("nan"), device=DEV) for it in items]
        poison_stages()
        if os.environ.get("FUZZ_VERBOSE"):
            print(f"batch {b}:", flush=True)
            for i, it in enumerate(items):
                print(f"  {i}: {it[4]}", flush=True)
        engine.execute(BatchPlan.concat([it[0] for it in items]), outs)
        torch.cuda.synchronize()
        for i, ((_, ref, exact, vs, _d), out) in enumerate(zip(items, outs)):
            got = out.cpu()
            ok = torch.equal(got, ref) if exact else torch.allclose(got, ref, rtol=1e-4, atol=1e-4 * vs)
            checked += 1
            if not ok:
                bad += 1
                d = (got - ref).abs()
                if verbose:
                    print(f"MISMATCH batch {b} item {i}: shape {tuple(ref.shape)} exact={exact} max|d|={float(d.max()):.3g} "
                          f"n={int((got != ref).sum())} nan={int(torch.isnan(got).sum())}")
    return checked, bad


def main():
    n_batches = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    big = len(sys.argv) > 3 and sys.argv[3] == "big"
    checked, bad = sweep(n_batches, seed, big)
    print(f"fuzz: {checked} items in {n_batches} launches, {bad} mismatches (seed {seed})")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
