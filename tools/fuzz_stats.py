"""Randomised sweep of the statistics kernels on the GPU against numpy (exact percentiles of the 'linear' method,
min / max, mean / std): adversarial distributions — constants, two values, heavy ties, signed zeros, denormals,
huge dynamic range, int16 / uint8 sources, one-element volumes, odd lengths and unaligned views, volumes of
different sizes in one call, per-volume and pooled (dataset_wide) histograms.

    python tools/fuzz_stats.py [n_rounds] [seed]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from adell_mri_b200 import stats

DEV = "cuda:0"


def random_volume(R):
    n = int(R.choice([1, 2, 3, 5, 17, 255, 1000, 4097, 65536, 200_003]))
    kind = R.choice(["uniform", "const", "two", "ties", "zeros", "denorm", "range", "normal", "i16", "u8", "lognormal"])
    if kind == "uniform":
        v = R.rand(n).astype(np.float32) * 1000
    elif kind == "const":
        v = np.full(n, np.float32(R.choice([0.0, -0.0, 3.5, -7.25, 1e-30, 1e30])), np.float32)
    elif kind == "two":
        v = R.choice(np.array([-1.5, 2.25], np.float32), size=n)
    elif kind == "ties":
        v = R.randint(0, 7, size=n).astype(np.float32)
    elif kind == "zeros":
        v = np.where(R.rand(n) < 0.7, R.choice(np.array([0.0, -0.0], np.float32), size=n), R.rand(n).astype(np.float32))
    elif kind == "denorm":
        v = (R.rand(n) * 1e-40).astype(np.float32) * R.choice(np.array([-1, 1], np.float32), size=n)
    elif kind == "range":
        v = (10.0 ** R.uniform(-30, 30, size=n)).astype(np.float32) * R.choice(np.array([-1, 1], np.float32), size=n)
    elif kind == "normal":
        v = R.normal(0, 100, size=n).astype(np.float32)
    elif kind == "lognormal":
        v = R.lognormal(0, 1, size=n).astype(np.float32)
    elif kind == "i16":
        v = R.randint(-32768, 32768, size=n).astype(np.int16)
    else:
        v = R.randint(0, 256, size=n).astype(np.uint8)
    return v, kind


def sweep(rounds, seed, verbose=True):
    """Returns (checks, mismatches)."""
    R = np.random.RandomState(seed)
    bad = checked = 0
    for r in range(rounds):
        dtype_kind = R.choice(["f32", "f32", "i16", "u8"])
        vols = []
        while len(vols) < int(R.randint(1, 6)):
            v, kind = random_volume(R)
            want = {"f32": np.float32, "i16": np.int16, "u8": np.uint8}[dtype_kind]
            if v.dtype == want:
                vols.append((v, kind))
        qs = sorted(float(q) for q in R.choice([0.0, 0.5, 1.0, 2.0, 25.0, 50.0, 75.0, 98.0, 99.0, 99.5, 100.0, float(R.uniform(0, 100))],
                                                size=int(R.randint(1, 5)), replace=False))
        devs = []
        for v, _ in vols:
            if R.rand() < 0.3 and v.size > 1:    # a view that is not 16-byte aligned
                buf = torch.empty(v.size + 1, dtype=torch.from_numpy(v).dtype, device=DEV)
                buf[1:].copy_(torch.from_numpy(v))
                devs.append(buf[1:])
            else:
                devs.append(torch.from_numpy(v).to(DEV))
        got = stats.percentiles(devs, qs).cpu().numpy()
        mm = stats.minmax(devs).cpu().numpy()
        for i, (v, kind) in enumerate(vols):
            f = v.astype(np.float32)
            ref = np.percentile(f, np.asarray(qs, np.float64)).astype(np.float32)
            checked += 1
            ok = np.array_equal(got[i], ref) and mm[i, 0] == f.min() and mm[i, 1] == f.max()
            if not ok:
                bad += 1
                if verbose:
                    print(f"MISMATCH round {r} vol {i} kind {kind} n {v.size} qs {qs}: got {got[i]} ref {ref} minmax {mm[i]} vs {f.min()} {f.max()}")
        pooled = stats.percentiles(devs, qs, dataset_wide=True).cpu().numpy()[0]
        ref = np.percentile(np.concatenate([v.astype(np.float32) for v, _ in vols]), np.asarray(qs, np.float64)).astype(np.float32)
        checked += 1
        if not np.array_equal(pooled, ref):
            bad += 1
            if verbose:
                print(f"MISMATCH round {r} pooled kinds {[k for _, k in vols]} qs {qs}: got {pooled} ref {ref}")
    return checked, bad


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    checked, bad = sweep(rounds, seed)
    print(f"fuzz_stats: {checked} checks in {rounds} rounds, {bad} mismatches (seed {seed})")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
