"""CPU ORACLE — numpy emulation of the three statistics entry points (adell_hist_pass,
adell_hist_select, adell_percentile_finalize; include/adell_b200.h).  Test infrastructure only:
lets the multi-rank host protocol of ``adell_mri_b200.stats.percentiles`` run on CPU (gloo)."""

from __future__ import annotations

import numpy as np
import torch


def key_f32(x: np.ndarray) -> np.ndarray:
    u = x.astype(np.float32).view(np.uint32)
    return np.where(u & 0x80000000, ~u, u | 0x80000000).astype(np.uint32)


def unkey_f32(k: np.ndarray) -> np.ndarray:
    k = k.astype(np.uint32)
    u = np.where(k & 0x80000000, k & 0x7FFFFFFF, ~k).astype(np.uint32)
    return u.view(np.float32)


class NumpyKernels:
    def __init__(self, vols):
        self.keys = [key_f32(v.numpy().reshape(-1)) for v in vols]

    def zeros(self, n, dtype):
        return torch.zeros(n, dtype=dtype)

    def upload(self, arr):
        return torch.from_numpy(arr.copy())

    def hist_pass(self, n_sel, shared, prefix, shift, bits, bins):
        first = shift + bits == 32
        b = bins.numpy().view(np.uint64)
        pre = prefix.numpy().view(np.uint32)
        nb = 1 << bits
        for v, k in enumerate(self.keys):
            h = 0 if shared else v
            if first:
                b[h * nb:(h + 1) * nb] += np.bincount(k >> np.uint32(shift), minlength=nb).astype(np.uint64)
            else:
                hi = shift + bits
                for s in range(n_sel):
                    m = (k >> np.uint32(hi)) == (pre[h * n_sel + s] >> np.uint32(hi))
                    sel = (k[m] >> np.uint32(shift)) & np.uint32(nb - 1)
                    o = (h * n_sel + s) * nb
                    b[o:o + nb] += np.bincount(sel, minlength=nb).astype(np.uint64)

    def hist_select(self, bins, n_hist, n_sel, shift, bits, prefix, rank):
        first = shift + bits == 32
        b = bins.numpy().view(np.uint64)
        pre = prefix.numpy().view(np.uint32)
        r = rank.numpy().view(np.uint64)
        nb = 1 << bits
        for h in range(n_hist):
            for s in range(n_sel):
                o = (h if first else h * n_sel + s) * nb
                cum = np.cumsum(b[o:o + nb])
                i = h * n_sel + s
                idx = int(np.searchsorted(cum, r[i], side="right"))
                before = 0 if idx == 0 else int(cum[idx - 1])
                pre[i] = (0 if first else pre[i]) | np.uint32(idx << shift)
                r[i] = r[i] - np.uint64(before)

    def finalize(self, prefix, frac, n_hist, n_q):
        k = prefix.numpy().view(np.uint32).reshape(n_hist, n_q, 2)
        v = unkey_f32(k.reshape(-1)).reshape(n_hist, n_q, 2)
        t = frac.numpy().reshape(n_hist, n_q)
        a, b = v[..., 0], v[..., 1]
        diff = (b - a).astype(np.float32)
        r = a.astype(np.float64) + diff.astype(np.float64) * t
        r = np.where(t >= 0.5, b.astype(np.float64) - diff.astype(np.float64) * (1 - t), r)
        return torch.from_numpy(r.astype(np.float32))
